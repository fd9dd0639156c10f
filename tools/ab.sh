#!/bin/bash
# A/B on one GPU box: the round-1 library (column-walk kernels, validated against the oracle in round 1) against the
# in-tree build (cell-parallel class kernels).  The state hashes must be identical.
OLD=${OLD:-tools/_ab/libkidmp_r1.so}
run() { echo "== $*"; KIDMP_LIB=$OLD python tools/state_hash.py "$@"; python tools/state_hash.py "$@"; }
run --steps 6
run --steps 4 --dt 60 --dz 100 --columns 262144
run --steps 4 --warm --columns 262144
run --steps 4 --columns 14400
run --steps 4 --columns 1
