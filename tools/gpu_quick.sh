#!/bin/bash
# Short GPU call: A/B the listed experiment builds, then GPU parity tests + bench on the default build.
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 100 --warmup 5 --no-cpu-baseline"
for v in "$@"; do
  n=$(basename $v .so)
  timeout 200 python tools/ab_bench.py $v $B > $O/ab_$n.json 2> $O/ab_$n.err
done
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 200 python tools/prof_nodes.py > $O/node_cycles.json 2> $O/node_cycles.err
tail -3 $O/pytest_gpu.log
for f in $O/ab_*.json $O/bench_n1.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value']), round(d['e2e']['value']), d['kernels_ms'], d['tables_cycles']['total'], d['dp_phase_cycles_mean_max']['total'])
except Exception as e:
    print('ERR', e)
PY
)"; done
