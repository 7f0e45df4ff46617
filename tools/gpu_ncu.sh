#!/bin/bash
# one ncu --set full capture of the gradient kernel on the bench workload (pass a kernel regex, default k_dp_rev)
set -u
mkdir -p gpurun_out
K=${1:-k_dp_rev}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -f -o gpurun_out/prof_$K \
    python bench.py --steps 3 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_$K.log 2>&1
tail -n 5 gpurun_out/ncu_$K.log
ls -la gpurun_out/
