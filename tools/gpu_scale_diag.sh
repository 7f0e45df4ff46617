#!/bin/bash
# Multi-GPU diagnosis of the C2 weak-scaling step: per-rank step / kernel times and the exchange alone, at N = 8, 4, 2
set -u
mkdir -p gpurun_out
O=gpurun_out
nproc > $O/diag_nproc.txt
for N in ${NS:-8 4 2}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+N)) bench.py --gpus $N --steps ${STEPS:-200} --warmup 10 --c3-families ${C3:-0} > $O/diag_n$N.json 2> $O/diag_n$N.err
  python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/diag_n{N}.json").read().strip().splitlines()[-1])
    print("N", N, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), d['config']['sharding'][-50:])
    r=d.get('ranks') or {}
    for k,v in r.items(): print("  ", k, [round(x,4) for x in v] if isinstance(v,list) else v)
    c=d.get('c3_strong')
    if c: print("   C3", round(c['value']), round(c['e2e']['value']), c['ms_per_step'], c.get('exchange'))
except Exception as e: print(N, "ERR", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $O/diag_n$N.err | tail -3
done
