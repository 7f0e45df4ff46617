#!/bin/bash
# compute-sanitizer over a small slice of the GPU parity suite (both gradient modes, backtracking, exchange kernels)
set -u
mkdir -p gpurun_out
O=gpurun_out
export WHALE_CALIBRATE=0
SEL="known_answer_single_family or constant_rates or keep_ell or track_sample or multi_device or one_persistent"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/san_memcheck.log 2>&1
echo "memcheck exit $?" >> $O/san_memcheck.log
WHALE_GRAD_MODE=rev timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "known_answer_single_family or constant_rates" > $O/san_racecheck_rev.log 2>&1
echo "racecheck exit $?" >> $O/san_racecheck_rev.log
WHALE_GRAD_MODE=fwd timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "known_answer_single_family or constant_rates" > $O/san_racecheck_fwd.log 2>&1
echo "racecheck exit $?" >> $O/san_racecheck_fwd.log
timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "peer_sum_between_processes" > $O/san_memcheck_peer.log 2>&1
echo "memcheck (peer exchange, both ranks) exit $?" >> $O/san_memcheck_peer.log
for f in $O/san_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit|Error|hazard" $f | tail -8; done
