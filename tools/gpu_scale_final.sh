#!/bin/bash
# The driver's scaling run on one 8-GPU box: bench.py (C2 weak + C3 strong, 10^5 families) at N = 8, 4, 2, 1
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -8 > $O/smi_scale.txt; nproc >> $O/smi_scale.txt
for N in ${NS:-8 4 2 1}; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline > $O/scale_n1.json 2> $O/scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 100 --warmup 5 > $O/scale_n$N.json 2> $O/scale_n$N.err
  fi
  python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/scale_n{N}.json").read().strip().splitlines()[-1])
    print("N", N, "C2", round(d['value']), "e2e", round(d['e2e']['value']), round(d['ms_per_step'],4), d['config']['sharding'][-60:])
    r=d.get('ranks') or {}
    if r: print("   k_dp_ms", [round(x,4) for x in r['k_dp_ms']], r['exchange_alone_us_rank0']['median'])
    c=d.get('c3_strong')
    if c: print("   C3", round(c['value']), "e2e", round(c['e2e']['value']), round(c['ms_per_step'],3), c.get('exchange'), c['loglik_last'], c['gen_s'], c['pack_s'])
except Exception as e: print(N, "ERR", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $O/scale_n$N.err | tail -3
done
