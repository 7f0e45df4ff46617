#!/bin/bash
set -u
mkdir -p gpurun_out
export WHALE_GRAD_MODE=rev
for s2 in 0 40000; do
  WHALE_STAGE2_MAX=$s2 timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --c3-families 0 > gpurun_out/s2_c2_$s2.json 2> gpurun_out/s2_c2_$s2.err
  for nf in 300 600; do
  WHALE_STAGE2_MAX=$s2 timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --c3-families 0 --families $nf > gpurun_out/s2_c2_${s2}_$nf.json 2> gpurun_out/s2_c2_${s2}_$nf.err
  done
python - $s2 <<'PY'
import json,sys
nt=sys.argv[1]
for suf in ("", "_300", "_600"):
    try:
        d=json.loads(open(f"gpurun_out/s2_c2_{nt}{suf}.json").read().strip().splitlines()[-1])
        print("C2 s2", nt, suf, round(d['value']), d['kernels_ms']['k_dp'], {k:round(v[0]) for k,v in d['dp_phase_cycles_mean_max'].items()})
    except Exception as e: print("C2 s2", nt, suf, "ERR", e)
PY
done
tail -n 2 gpurun_out/s2_c2*.err | grep -v "^$\|==>" | head
