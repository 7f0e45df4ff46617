#!/bin/bash
# 2-GPU call: multi-device C-ABI test, bench with peer exchange vs NCCL
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_device or known_answer" > $O/pytest_gpu2.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu2.log
tail -3 $O/pytest_gpu2.log
for ex in peer nccl; do
WHALE_BENCH_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --c3-families 20000 > $O/bench_n2_$ex.json 2> $O/bench_n2_$ex.err
python - $ex <<'PY'
import json,sys
ex=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_n2_{ex}.json").read().strip().splitlines()[-1])
    print(ex, "C2", round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['config']['sharding'])
    c=d['c3_strong']; print(ex, "C3", round(c['value']), round(c['e2e']['value']), c['ms_per_step'], c['exchange'], c['loglik_last'])
except Exception as e: print(ex, "ERR", e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $O/bench_n2_$ex.err | tail -5
done
