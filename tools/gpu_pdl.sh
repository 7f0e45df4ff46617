#!/bin/bash
# A/B of programmatic dependent launch (k_dp / k_dp_rev behind the table kernel): WHALE_PDL=0 vs 1, C2 + C3 (12 500 families),
# then the GPU parity suite on the default (PDL on).
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 200 --warmup 10 --no-cpu-baseline --c3-families 12500"
for rep in 1 2; do
for v in 0 1; do
  WHALE_PDL=$v timeout 300 python bench.py $B > $O/pdl${v}_r$rep.json 2> $O/pdl${v}_r$rep.err
  echo "PDL=$v rep $rep: $(python - "$O/pdl${v}_r$rep.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c=d.get('c3_strong') or {}
    print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['kernels_ms'].items() if isinstance(v,float)}, 'c3', round(c.get('value',0)), c.get('ms_per_step'))
except Exception as e:
    print('ERR', e)
PY
)"
done
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
