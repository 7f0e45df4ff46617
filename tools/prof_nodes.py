#!/usr/bin/env python
"""Per-node SM-cycle breakdown of k_dp on the C2 bench workload (profiling aid; prints JSON).

    python tools/prof_nodes.py [--families 1000] [--reps 20]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--families", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import whale_jl_b200 as W
    from whale_jl_b200 import synth, lib as wlib
    from whale_jl_b200.core import _data_handle
    d = synth.cache_dir(f"c2_seed2_rank0_n{args.families}")
    synth.generate(d, args.families, seed=2)
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), synth.c1_species_tree(), 0.05)
    ccd = W.read_ale(d, w)
    L = wlib.get()
    mh, dh = _data_handle(w, ccd)
    x = w.x()
    ms = []
    for _ in range(args.reps):
        L.logpdf_grad(mh, dh, x, w.p_leaf(), 1, want_grad=True, profile=True)
        ms.append(L.last_kernel_ms(dh))
    ms = np.array(ms)[3:]
    sl, r1 = L.last_node_cycles(dh)
    inner = [int(e) for e in w.order if w.kind[e] != 0]
    rev = L.L.whale_data_grad_mode(dh) == 1  # reverse mode: per node [forward total, transposed total] (whale_rev.cuh)
    out = {"grad_mode": "reverse" if rev else "forward", "kernels_ms": dict(zip(["k_tables", "k_dp", "k_reduce"], ms.mean(0).round(4).tolist())),
           "phases": L.last_phase_cycles(dh),
           "nodes": [{"node": e, "kind": int(w.kind[e]), "n_slices": int(w.n_slices[e]), "slices_cycles": round(a),
                      "per_slice": round(a / max(1, int(w.n_slices[e]))), "stage_row1_cycles": round(b)}
                     for e, a, b in zip(inner, sl, r1)]}
    fc = L.last_family_cycles(dh)
    order = np.argsort(-fc[:, 6])
    out["family_total_percentiles"] = {str(q): float(np.percentile(fc[:, 6], q)) for q in (50, 90, 99, 100)}
    out["slowest"] = [{"fam": int(f), "G": len(ccd[int(f)].nleaf), "leaves": len(ccd[int(f)].leaves),
                       "phases": [round(v) for v in fc[f, :7]]} for f in order[:12]]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
