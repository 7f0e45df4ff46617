#!/bin/bash
# One GPU call for the round's artefacts on the default build: GPU parity tests, bench (+ reference arm), per-node
# cycle profile, ncu launch list, ncu full captures of k_dp and k_tables, the other BASELINE configs and the
# backtracking rate.  Everything lands in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 200 python tools/prof_nodes.py > $O/node_cycles.json 2> $O/node_cycles.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_dp -s 4 -c 2 -o $O/prof_dp \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_tables -s 4 -c 1 -o $O/prof_tab \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_tab.log 2>&1
timeout 400 python tools/bench_configs.py > $O/configs_c3_c4.json 2> $O/configs_c3_c4.err
timeout 300 python tools/bench_track.py > $O/backtrack_c5.json 2> $O/backtrack_c5.err
tail -3 $O/pytest_gpu.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json", "gpurun_out/bench_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), d.get("e2e", {}).get("value"), d.get("kernels_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
