#!/bin/bash
# A/B on one GPU box: kid_b200/libkidmp_prev.so (a build of an earlier commit) against the in-tree library.
# State hashes must be identical for a bit-exact optimisation; times are comparable because it is the same box.
for L in prev new prev new; do
  if [ $L = prev ]; then export KIDMP_LIB=$PWD/kid_b200/libkidmp_prev.so; else unset KIDMP_LIB; fi
  python tools/state_hash.py --steps 8
done
for L in prev new; do
  if [ $L = prev ]; then export KIDMP_LIB=$PWD/kid_b200/libkidmp_prev.so; else unset KIDMP_LIB; fi
  python tools/state_hash.py --steps 4 --dt 60 --dz 100 --columns 262144
  python tools/state_hash.py --steps 4 --warm --columns 262144
done
