#!/bin/bash
# A/B of experiment builds against the default library on one box: bench.py (C2 only) per variant, twice, interleaved
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 200 --warmup 10 --no-cpu-baseline --c3-families ${C3:-0}"
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c=d.get('c3_strong') or {}
    print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), round(d['kernels_ms']['k_dp'],4), {k: round(v[0]) for k,v in d['dp_phase_cycles_mean_max'].items()}, 'c3', round(c.get('value',0)))
except Exception as e:
    print('ERR', e)
PY
}
for rep in 1 2; do
  timeout 300 python bench.py $B > $O/ab_default_r$rep.json 2> $O/ab_default_r$rep.err
  echo "default rep $rep: $(show $O/ab_default_r$rep.json)"
  for v in "$@"; do
    n=$(basename $v .so)
    timeout 300 python tools/ab_bench.py $v $B > $O/ab_${n}_r$rep.json 2> $O/ab_${n}_r$rep.err
    echo "$n rep $rep: $(show $O/ab_${n}_r$rep.json)"
  done
done
