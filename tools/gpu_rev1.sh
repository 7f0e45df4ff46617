#!/bin/bash
# round-2 GPU call: parity tests + bench in both gradient modes + the other configs
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_rev.json 2> $O/bench_rev.err
WHALE_GRAD_MODE=fwd timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_fwd.json 2> $O/bench_fwd.err
timeout 600 python tools/bench_configs.py > $O/configs_rev.json 2> $O/configs_rev.err
WHALE_GRAD_MODE=fwd timeout 600 python tools/bench_configs.py > $O/configs_fwd.json 2> $O/configs_fwd.err
tail -5 $O/pytest_gpu.log
for f in $O/bench_rev.json $O/bench_fwd.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value']), round(d['e2e']['value']), d['kernels_ms'], d['dp_phase_cycles_mean_max'])
except Exception as e:
    print('ERR', e)
PY
)"; done
cat $O/configs_rev.json $O/configs_fwd.json
tail -3 $O/*.err
