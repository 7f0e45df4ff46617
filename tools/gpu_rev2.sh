#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_auto.json 2> $O/bench_auto.err
timeout 900 python tools/bench_configs.py > $O/configs_auto.json 2> $O/configs_auto.err
tail -5 $O/pytest_gpu.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_auto.json").read().strip().splitlines()[-1])
print("bench", round(d['value']), round(d['e2e']['value']), d['kernels_ms'])
PY
cat $O/configs_auto.json
tail -n 3 $O/*.err
