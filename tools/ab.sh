#!/bin/bash
# A/B on one GPU box of the physics-kernel variants (same bits expected: compare the state hashes).
# KIDMP_UNITS: 1 = unit-parallel kernel, 0 = column-walk kernel, -1 = by domain size (default).
# KIDMP_FUSE (column-walk only): 0 / 1 / 2.
run() { echo "== units=$1"; KIDMP_UNITS=$1 python tools/state_hash.py "${@:2}" | cut -c1-8,60-400; }
for U in 0 1 0 1; do run $U --steps 6; done
for U in 0 1; do run $U --steps 4 --dt 60 --dz 100 --columns 262144; run $U --steps 4 --warm --columns 262144; run $U --steps 4 --columns 14400; run $U --steps 4 --columns 1; done
