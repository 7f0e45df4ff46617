#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 900 python tools/bench_track.py --families 10000 --samples 100 > $O/track_c5_n1.json 2> $O/track_c5_n1.err
cat $O/track_c5_n1.json; tail -n 3 $O/track_c5_n1.err
timeout 900 python tools/check_track_parity.py --families 1000 --samples 100 > $O/track_parity.json 2> $O/track_parity.err
cat $O/track_parity.json; tail -n 3 $O/track_parity.err
