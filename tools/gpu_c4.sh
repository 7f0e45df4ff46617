#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "c4_shape_16" > $O/pytest_c4.log 2>&1
echo "pytest exit $?" >> $O/pytest_c4.log
tail -3 $O/pytest_c4.log
timeout 1500 python tools/bench_configs.py --c3-families 2000 --c4-families ${1:-2000} --reps 6 > $O/configs_c4_2000.json 2> $O/configs_c4_2000.err
cat $O/configs_c4_2000.json; tail -n 3 $O/configs_c4_2000.err
