#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
NF=${1:-12500}
for mb in 4 5 6 8; do
  WHALE_REV_MINB=$mb timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > $O/minb_c2_$mb.json 2> $O/minb_c2_$mb.err
  WHALE_REV_MINB=$mb timeout 600 python tools/bench_configs.py --only c3 --c3-families $NF --reps 10 > $O/minb_c3_$mb.json 2> $O/minb_c3_$mb.err
python - $mb <<'PY'
import json,sys
nt=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/minb_c2_{nt}.json").read().strip().splitlines()[-1])
    print("C2 minb", nt, round(d['value']), round(d['e2e']['value']), d['kernels_ms']['k_dp'])
except Exception as e: print("C2 minb", nt, "ERR", e)
try:
    d=json.loads(open(f"gpurun_out/minb_c3_{nt}.json").read().strip().splitlines()[-1])
    print("C3 minb", nt, d)
except Exception as e: print("C3 minb", nt, "ERR", e)
PY
done
tail -n 2 $O/minb_*.err
