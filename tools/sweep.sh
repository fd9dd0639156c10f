#!/bin/bash
# usage: tools/sweep.sh "8 12 16"  -> kernel ms for the bench domain per KIDMP_MINB value
for m in $1; do
  echo "MINB=$m"; KIDMP_MINB=$m python tools/time_step.py --columns 1048576 --steps 4
done
