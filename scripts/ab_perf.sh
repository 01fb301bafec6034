#!/bin/bash
# A/B timing of stage-kernel variants on one box: each variant is a full library build under build/ab/.
# usage: scripts/ab_perf.sh tag mode...    (modes of scripts/dev_perf.py)
tag=$1; shift
for lib in build/ab/*.so; do
  name=$(basename $lib .so)
  for mode in "$@"; do
    THETIS_B200_LIB=$PWD/$lib python scripts/dev_perf.py --mode $mode --steps 100 2>&1 | tail -2 | sed "s/^/$name: /" >> gpurun_out/${tag}_ab.txt
  done
done
# second pass in reverse order (clock / thermal drift shows up as a difference between the passes)
for lib in $(ls -r build/ab/*.so); do
  name=$(basename $lib .so)
  for mode in "$@"; do
    THETIS_B200_LIB=$PWD/$lib python scripts/dev_perf.py --mode $mode --steps 100 2>&1 | tail -2 | sed "s/^/$name (pass 2): /" >> gpurun_out/${tag}_ab.txt
  done
done
