#!/bin/bash
# C5 backtracking at full size (numbers) + ncu --set full of k_backtrack and of the ℓ-keeping k_dp (HBM roofline of that mode)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python tools/bench_track.py --families 10000 --samples 100 > $O/track_c5_n1.json 2> $O/track_c5_n1.err
cat $O/track_c5_n1.json; tail -n 3 $O/track_c5_n1.err
WHALE_CALIBRATE=0 timeout 600 ncu --set full --clock-control none --import-source on -k k_backtrack -s 1 -c 1 -f -o $O/prof_bt \
    python tools/bench_track.py --families 10000 --samples 100 > $O/ncu_bt.log 2>&1
WHALE_CALIBRATE=0 timeout 600 ncu --set full --clock-control none --import-source on -k k_dp -s 1 -c 1 -f -o $O/prof_keep \
    python tools/bench_track.py --families 10000 --samples 100 > $O/ncu_keep.log 2>&1
for r in prof_bt prof_keep; do
  if [ -f $O/$r.ncu-rep ]; then
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
    rm -f $O/$r.ncu-rep
    python tools/ncu_summary.py $O/$r.raw.csv | head -40
  fi
done
