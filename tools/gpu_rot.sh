#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
for r in 1 0; do
  WHALE_REV_ROT=$r timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > $O/rot_c2_$r.json 2> $O/rot_c2_$r.err
  WHALE_REV_ROT=$r timeout 600 python tools/bench_configs.py --only c3 --c3-families 12500 --reps 6 > $O/rot_c3_$r.json 2> $O/rot_c3_$r.err
python - $r <<'PY'
import json,sys
nt=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/rot_c2_{nt}.json").read().strip().splitlines()[-1])
    print("C2 rot", nt, round(d['value']), round(d['e2e']['value']), d['kernels_ms']['k_dp'], d['dp_phase_cycles_mean_max'])
except Exception as e: print("C2 rot", nt, "ERR", e)
try:
    d=json.loads(open(f"gpurun_out/rot_c3_{nt}.json").read().strip().splitlines()[-1])
    print("C3 rot", nt, d)
except Exception as e: print("C3 rot", nt, "ERR", e)
PY
done
tail -n 2 $O/rot_*.err
