#!/bin/bash
# 8-GPU call: bench with the peer-memory exchange (C2 weak + C3 strong 10^5 families), then NCCL for comparison (C2 only)
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
nvidia-smi -L | head -8 > $O/smi_n$N.txt; nproc >> $O/smi_n$N.txt
WHALE_BENCH_EXCHANGE=peer timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 5 > $O/bench_n${N}_peer.json 2> $O/bench_n${N}_peer.err
WHALE_BENCH_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 100 --warmup 5 --c3-families 0 > $O/bench_n${N}_nccl.json 2> $O/bench_n${N}_nccl.err
for ex in peer nccl; do
python - $ex $N <<'PY'
import json,sys
ex,N=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/bench_n{N}_{ex}.json").read().strip().splitlines()[-1])
    print(ex, "C2", round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['config']['sharding'][:60])
    c=d.get('c3_strong')
    if c: print(ex, "C3", round(c['value']), round(c['e2e']['value']), c['ms_per_step'], c['exchange'], c['loglik_last'], c['gen_s'], c['pack_s'])
except Exception as e: print(ex, "ERR", e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $O/bench_n${N}_$ex.err | tail -4
done
