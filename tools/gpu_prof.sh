#!/bin/bash
# quick GPU call: per-phase / per-node cycle profile of the gradient kernel on C2 + a short bench
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/prof_nodes.py > $O/node_cycles.json 2> $O/node_cycles.err
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_rev.json 2> $O/bench_rev.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/node_cycles.json"))
print(d["kernels_ms"]); print(d["phases"])
for n in d["nodes"]: print(n)
try:
    b=json.loads(open("gpurun_out/bench_rev.json").read().strip().splitlines()[-1])
    print("bench", round(b["value"]), round(b["e2e"]["value"]), b["kernels_ms"])
except Exception as e: print("ERR", e)
PY
tail -n 3 $O/node_cycles.err $O/bench_rev.err
