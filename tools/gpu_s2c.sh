#!/bin/bash
set -u
mkdir -p gpurun_out
export WHALE_GRAD_MODE=rev
timeout 300 python -m pytest tests -m gpu -x -q -k "gradient_modes or one_persistent or known or c4_shape_16" > gpurun_out/pytest_rev.log 2>&1; tail -2 gpurun_out/pytest_rev.log
for s2 in default 0; do
  if [ $s2 = default ]; then unset WHALE_STAGE2_MAX; else export WHALE_STAGE2_MAX=$s2; fi
  timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --c3-families 0 > gpurun_out/s2_c2_$s2.json 2> gpurun_out/s2_c2_$s2.err
  timeout 600 python tools/bench_configs.py --only c3 --c3-families 12500 --reps 6 > gpurun_out/s2_c3_$s2.json 2> gpurun_out/s2_c3_$s2.err
python - $s2 <<'PY'
import json,sys
nt=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/s2_c2_{nt}.json").read().strip().splitlines()[-1])
    print("C2 s2", nt, round(d['value']), d['kernels_ms']['k_dp'], {k:round(v[0]) for k,v in d['dp_phase_cycles_mean_max'].items()})
except Exception as e: print("C2 s2", nt, "ERR", e)
try:
    d=json.loads(open(f"gpurun_out/s2_c3_{nt}.json").read().strip().splitlines()[-1])
    print("C3 s2", nt, round(d['C3']['evals_per_s']), d['C3']['first_pass_kernels_ms'])
except Exception as e: print("C3 s2", nt, "ERR", e)
PY
done
