#!/bin/bash
set -u
mkdir -p gpurun_out
export WHALE_GRAD_MODE=rev
for s2 in 0 20000 28000 40000; do
  WHALE_STAGE2_MAX=$s2 timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --c3-families 0 > gpurun_out/s2_c2_$s2.json 2> gpurun_out/s2_c2_$s2.err
  WHALE_STAGE2_MAX=$s2 timeout 600 python tools/bench_configs.py --only c3 --c3-families 12500 --reps 6 > gpurun_out/s2_c3_$s2.json 2> gpurun_out/s2_c3_$s2.err
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
tail -n 2 gpurun_out/s2_*.err | grep -v "^$\|==>" | head
