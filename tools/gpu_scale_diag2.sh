#!/bin/bash
# N = 8: same data on every rank (is a slow rank the GPU or its shard?), then the normal shards again, with per-rank clocks
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv > $O/diag_smi8.txt
WHALE_BENCH_SAME_DATA=1 NS=8 STEPS=200 bash tools/gpu_scale_diag.sh
mv $O/diag_n8.json $O/diag_n8_samedata.json
NS=8 STEPS=200 bash tools/gpu_scale_diag.sh
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv >> $O/diag_smi8.txt
cat $O/diag_smi8.txt
