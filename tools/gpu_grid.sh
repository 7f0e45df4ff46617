#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
NF=${1:-12500}
for g in 148 296 444 592; do
  WHALE_REV_GRID=$g timeout 600 python tools/bench_configs.py --only c3 --c3-families $NF --reps 6 > $O/grid_c3_$g.json 2> $O/grid_c3_$g.err
  echo "grid $g: $(cut -c1-220 $O/grid_c3_$g.json)"
done
tail -n 2 $O/grid_*.err
