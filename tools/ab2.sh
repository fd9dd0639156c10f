#!/bin/bash
# A/B of the in-tree library against other builds on the same box, with the per-kernel breakdown: OLD="tools/_ab/libkidmp_base.so ..."
OLD=${OLD:-tools/_ab/libkidmp_base.so}
for i in 1 2; do
  for l in $OLD; do KIDMP_LIB=$l python tools/state_hash.py --steps 6 --timing | cut -c1-400; done
  python tools/state_hash.py --steps 6 --timing | cut -c1-400
done
