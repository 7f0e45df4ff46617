#!/bin/bash
# launch-shape sweep of k_dp_rev: C2 bench (1000 families) and C3 (branch-wise rates, many families)
set -u
mkdir -p gpurun_out
O=gpurun_out
NF=${1:-12500}
for nt in 128 64 32; do
  WHALE_REV_NT=$nt timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > $O/shape_c2_nt$nt.json 2> $O/shape_c2_nt$nt.err
  WHALE_REV_NT=$nt timeout 600 python tools/bench_configs.py --only c3 --c3-families $NF --reps 10 > $O/shape_c3_nt$nt.json 2> $O/shape_c3_nt$nt.err
done
WHALE_GRAD_MODE=fwd timeout 600 python tools/bench_configs.py --only c3 --c3-families $NF --reps 5 > $O/shape_c3_fwd.json 2> $O/shape_c3_fwd.err
for nt in 128 64 32; do
python - $nt <<'PY'
import json,sys
nt=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/shape_c2_nt{nt}.json").read().strip().splitlines()[-1])
    print("C2 nt", nt, round(d['value']), round(d['e2e']['value']), d['kernels_ms']['k_dp'])
except Exception as e: print("C2 nt", nt, "ERR", e)
try:
    d=json.loads(open(f"gpurun_out/shape_c3_nt{nt}.json").read().strip().splitlines()[-1])
    print("C3 nt", nt, d)
except Exception as e: print("C3 nt", nt, "ERR", e)
PY
done
cat $O/shape_c3_fwd.json
tail -n 2 $O/shape_*.err
