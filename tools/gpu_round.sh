#!/bin/bash
# One GPU call: A/B the experiment builds (dev builds: default k_dp variant only), pick the fastest slice-loop
# variant and install its full build as the default library, then on that build: GPU parity tests, bench,
# per-node cycle profile, reference arm, ncu launch list and ncu full captures.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 100 --warmup 5 --no-cpu-baseline"
val() { python - "$1" <<'PY'
import json,sys
try:
    print(json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])['value'])
except Exception:
    print(0)
PY
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
for v in build/ab/v*.so; do
  n=$(basename $v .so)
  timeout 200 python tools/ab_bench.py $v $B > $O/ab_$n.json 2> $O/ab_$n.err
done
best=v3; bv=$(val $O/ab_v3_slice_hoist.json)
v2=$(val $O/ab_v2_slice.json); v1=$(val $O/ab_v1_tables_reduce.json)
if python -c "import sys; sys.exit(0 if $v2 > 1.01*$bv else 1)"; then best=v2; bv=$v2; fi
if python -c "import sys; sys.exit(0 if $v1 > 1.01*$bv else 1)"; then best=v1; bv=$v1; fi
echo "slice variant chosen: $best ($bv) [v1=$v1 v2=$v2 v3=$(val $O/ab_v3_slice_hoist.json)]" | tee $O/choice.txt
if [ $best != v3 ]; then cp build/ab/full_$best.so whale.jl_b200/libwhalecuda.so; fi
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 200 python tools/prof_nodes.py > $O/node_cycles.json 2> $O/node_cycles.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_dp -s 4 -c 2 -o $O/prof_dp \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
# attribution and occupancy variants on the chosen build (runtime switches)
WHALE_TABLES_CHAIN=1 timeout 200 python bench.py $B > $O/ab_default_chain.json 2> $O/ab_default_chain.err
WHALE_FUSED_REDUCE=0 timeout 200 python bench.py $B > $O/ab_default_nofuse.json 2> $O/ab_default_nofuse.err
WHALE_MINB=5 timeout 200 python bench.py $B > $O/ab_default_nt128_mb5.json 2> $O/ab_default_nt128_mb5.err
WHALE_MINB=6 timeout 200 python bench.py $B > $O/ab_default_nt128_mb6.json 2> $O/ab_default_nt128_mb6.err
WHALE_NT=96 WHALE_MINB=6 timeout 200 python bench.py $B > $O/ab_default_nt96_mb6.json 2> $O/ab_default_nt96_mb6.err
WHALE_NT=64 WHALE_MINB=8 timeout 200 python bench.py $B > $O/ab_default_nt64_mb8.json 2> $O/ab_default_nt64_mb8.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_tables -s 4 -c 1 -o $O/prof_tab \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_tab.log 2>&1
tail -3 $O/pytest_gpu.log
for f in $O/ab_*.json $O/bench_n1.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value']), round(d['e2e']['value']), d['kernels_ms'])
except Exception as e:
    print('ERR', e)
PY
)"; done
