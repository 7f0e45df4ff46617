#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one GPU (they are parity-test cases, not bench.py
lines; this tool records their rates for DESIGN.md):

    C3  9-taxon DLWGD, branch-wise rates (P = 37), families per GPU as given      (configs[2], per-GPU shard)
    C4  30-taxon tree + 5 WGDs, CCDs of ~2,000 clades, constant (P = 8) and branch-wise (P = 122) rates (configs[3])

    python tools/bench_configs.py [--c3-families 2000] [--c4-families 64] [--reps 10]

Each evaluation is one whale_logpdf_grad call with host pointers (new θ per call); the rate is families × reps /
wall time after 3 warm-up calls."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rate(L, mh, dh, w, xs, reps, F, cond=1):
    for i in range(3):
        L.logpdf_grad(mh, dh, xs[i], w.p_leaf(), cond, want_grad=True)
    t0 = time.perf_counter()
    for i in range(reps):
        ll = L.logpdf_grad(mh, dh, xs[3 + i], w.p_leaf(), cond, want_grad=True)[0]
    dt = time.perf_counter() - t0
    L.logpdf_grad(mh, dh, xs[0], w.p_leaf(), cond, want_grad=True, profile=True)
    km = L.last_kernel_ms(dh)
    return {"families": F, "evals_per_s": F * reps / dt, "ms_per_eval": 1e3 * dt / reps, "loglik": ll,
            "first_pass_kernels_ms": dict(zip(["k_tables", "k_dp", "k_reduce"], [round(v, 4) for v in km]))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-families", type=int, default=2000)
    ap.add_argument("--c4-families", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="", help="comma list of c3,c4 (default: both)")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    import whale_jl_b200 as W
    from whale_jl_b200 import synth, newick, lib as wlib
    from whale_jl_b200.core import _data_handle
    L = wlib.get()
    rng = np.random.default_rng(7)
    out = {}
    # ---- C3: branch-wise rates on the 9-taxon tree ----
    d = synth.generate(synth.cache_dir(f"c3_seed3_n{args.c3_families}"), args.c3_families, seed=3)
    r = W.DLWGD(lam=list(rng.normal(np.log(0.15), 0.3, 17)), mu=list(rng.normal(np.log(0.15), 0.3, 17)), q=[0.2, 0.1], eta=0.67)
    w = W.WhaleModel(r, synth.c1_species_tree(), 0.05)
    ccd = W.read_ale_native(d, w)
    mh, dh = _data_handle(w, ccd)
    x0 = w.x()
    xs = x0[None, :] + 0.02 * rng.standard_normal((args.reps + 3, len(x0)))
    xs[:, -3:] = np.clip(xs[:, -3:], 1e-3, 1 - 1e-3)
    out["C3"] = dict(rate(L, mh, dh, w, xs, args.reps, len(ccd)), P=int(w.n_params),
                     grad_mode="reverse" if L.L.whale_data_grad_mode(dh) == 1 else "forward",
                     gradient_passes=int(L.L.whale_data_grad_passes(dh)))
    ccd.close()
    if only and "c4" not in only:
        print(json.dumps(out))
        return
    # ---- C4: 30 taxa, 5 WGDs, ~2,000 clades ----
    nws = newick.nwstr(synth.c4_species_tree(), True) + ";"
    d = synth.generate(synth.cache_dir(f"c4_seed4_n{args.c4_families}"), args.c4_families, seed=4,
                       tree=synth.c4_species_tree(), **synth.C4_FAMILY)
    q = [0.2, 0.1, 0.2, 0.1, 0.2]
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=q, eta=0.67), newick.readnw(nws), 0.05)
    t0 = time.perf_counter()
    ccd = W.read_ale_native(d, w)
    mh, dh = _data_handle(w, ccd)
    out["C4_ingest_pack_s"] = round(time.perf_counter() - t0, 2)
    x0 = w.x()
    xs = x0[None, :] * np.exp(0.03 * rng.standard_normal((args.reps + 3, len(x0))))
    xs[:, 2:] = np.clip(xs[:, 2:], 1e-3, 1 - 1e-3)
    out["C4_constant"] = dict(rate(L, mh, dh, w, xs, args.reps, len(ccd)), P=int(w.n_params),
                              clades_median=int(np.median(ccd.n_clades)),
                              grad_mode="reverse" if L.L.whale_data_grad_mode(dh) == 1 else "forward",
                              gradient_passes=int(L.L.whale_data_grad_passes(dh)), arena_bytes=int(L.L.whale_data_arena_bytes(dh)))
    ccd.close()
    rb = W.DLWGD(lam=list(rng.normal(np.log(0.15), 0.3, 59)), mu=list(rng.normal(np.log(0.15), 0.3, 59)), q=q, eta=0.67)
    wb = W.WhaleModel(rb, newick.readnw(nws), 0.05)
    ccdb = W.read_ale_native(d, wb)
    mh, dh = _data_handle(wb, ccdb)
    x0 = wb.x()
    xs = x0[None, :] + 0.02 * rng.standard_normal((max(args.reps // 3, 2) + 3, len(x0)))
    xs[:, -6:] = np.clip(xs[:, -6:], 1e-3, 1 - 1e-3)
    out["C4_branchwise"] = dict(rate(L, mh, dh, wb, xs, max(args.reps // 3, 2), len(ccdb)), P=int(wb.n_params),
                                grad_mode="reverse" if L.L.whale_data_grad_mode(dh) == 1 else "forward",
                                gradient_passes=int(L.L.whale_data_grad_passes(dh)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
