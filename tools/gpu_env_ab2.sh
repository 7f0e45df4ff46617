#!/bin/bash
# Runtime-switch experiments on the default build: each line of $1 is "name ENV=.. ENV=.."; C2 + a C3 leg of $C3 families
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 200 --warmup 10 --no-cpu-baseline --c3-families ${C3:-12500}"
while read -r name envs; do
  [ -z "$name" ] && continue
  env $envs timeout 300 python bench.py $B > $O/env_$name.json 2> $O/env_$name.err
  echo "$name: $(python - "$O/env_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c=d.get('c3_strong') or {}
    print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), round(d['kernels_ms']['k_dp'],4), {k: round(v[0]) for k,v in d['dp_phase_cycles_mean_max'].items()}, 'c3', round(c.get('value',0)), c.get('ms_per_step'))
except Exception as e:
    print('ERR', e)
PY
)"
done < "$1"
