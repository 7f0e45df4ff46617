#!/bin/bash
# Runtime-switch experiments on the default build: each line of $1 is "name ENV=.. ENV=.."
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 100 --warmup 5 --no-cpu-baseline"
while read -r name envs; do
  [ -z "$name" ] && continue
  env $envs timeout 200 python bench.py $B > $O/env_$name.json 2> $O/env_$name.err
  echo "$name: $(python - "$O/env_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value']), round(d['e2e']['value']), {k: round(v,4) for k,v in d['kernels_ms'].items()}, d['dp_phase_cycles_mean_max']['total'])
except Exception as e:
    print('ERR', e)
PY
)"
done < "$1"
