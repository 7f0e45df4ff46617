#!/bin/bash
set -u
mkdir -p gpurun_out
export WHALE_GRAD_MODE=rev
timeout 300 python -m pytest tests -m gpu -x -q -k "gradient_modes or one_persistent or known" > gpurun_out/pytest_rev.log 2>&1; tail -3 gpurun_out/pytest_rev.log
bash tools/gpu_prof.sh
timeout 600 python tools/bench_configs.py --only c3 --c3-families 12500 --reps 6 | cut -c1-330
