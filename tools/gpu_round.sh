#!/bin/bash
# One GPU call for the round's artefacts on the default build: GPU parity tests, smoke, bench (+ reference arm), per-node
# cycle profile, ncu launch list, ncu full captures of k_dp (C2, forward tangents), k_dp_rev (C3 leg, reverse mode) and
# k_tables3.  Everything lands in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
O=gpurun_out
C3=${C3:-25000}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 200 python tools/prof_nodes.py > $O/node_cycles_fwd.json 2> $O/node_cycles.err
WHALE_GRAD_MODE=rev timeout 200 python tools/prof_nodes.py > $O/node_cycles_rev.json 2>> $O/node_cycles.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --c3-families $C3 > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_dp -s 4 -c 1 -f -o $O/prof_dp \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --c3-families 0 > $O/ncu_dp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dp_rev -s 3 -c 1 -f -o $O/prof_rev \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --c3-families $C3 > $O/ncu_rev.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tables3 -s 3 -c 1 -f -o $O/prof_tab3 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --c3-families $C3 > $O/ncu_tab3.log 2>&1
# the reports are large: keep their raw and source pages as CSV, drop the .ncu-rep files (gpurun_out is capped at 64 MiB)
for r in prof_dp prof_rev prof_tab3; do
  if [ -f $O/$r.ncu-rep ]; then
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
    ncu -i $O/$r.ncu-rep --page source --csv > $O/$r.source.csv 2>/dev/null
    rm -f $O/$r.ncu-rep
  fi
done
tail -3 $O/pytest_gpu.log; tail -3 $O/smoke.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json", "gpurun_out/bench_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), d.get("e2e", {}).get("value"), d.get("kernels_ms"), (d.get("c3_strong") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
