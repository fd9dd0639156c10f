#!/bin/bash
# A/B on one GPU box: kid_b200/libkidmp_prev.so (a build of an earlier commit) against the in-tree library.
# State hashes must be identical for a bit-exact optimisation; times are comparable because it is the same box.
# KIDMP_FUSE: 0 = split kernels only, 1 = adaptive (default), 2 = always fused + redo of the columns with sub-steps.
run() { echo "== $1 fuse=${2:-default}"; if [ $1 = prev ]; then export KIDMP_LIB=$PWD/kid_b200/libkidmp_prev.so; else unset KIDMP_LIB; fi
        if [ -n "$2" ]; then export KIDMP_FUSE=$2; else unset KIDMP_FUSE; fi; shift; shift; python tools/state_hash.py "$@" | cut -c1-8,60-400; }
for L in prev new; do run $L "" --steps 8; done
run new 0 --steps 5
run new 2 --steps 5
for L in prev new; do run $L "" --steps 4 --dt 60 --dz 100 --columns 262144; run $L "" --steps 4 --warm --columns 262144; done
run new 2 --steps 4 --dt 60 --dz 100 --columns 262144
run new 0 --steps 4 --dt 60 --dz 100 --columns 262144
run new 2 --steps 4 --warm --columns 262144
