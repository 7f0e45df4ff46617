#!/bin/bash
# C5 (backtracking, 10^4 families x 100 samples in total) sharded over N = 8 and 4 ranks of one box; no exchange on this path
set -u
mkdir -p gpurun_out
O=gpurun_out
for N in ${NS:-8 4}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29560+N)) tools/bench_track.py --families 10000 --samples 100 > $O/track_c5_n$N.json 2> $O/track_c5_n$N.err
  python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/track_c5_n{N}.json").read().strip().splitlines()[-1])
    print("N", N, "walks", round(d['walks_device_rng']['trees_per_s_total']), "track_sample", round(d['track_sample']['trees_per_s_total']), d['walks_device_rng'].get('ms_max_over_ranks'), d['track_sample'].get('ms_max_over_ranks'))
except Exception as e: print(N, "ERR", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $O/track_c5_n$N.err | tail -3
done
