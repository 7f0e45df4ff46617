#!/usr/bin/env python
"""Backtracking throughput (BASELINE config 5 shape): reconciled trees per second on one GPU.

    python tools/bench_track.py [--families 1000] [--samples 100]

logpdf! (keep ℓ) once, then `n_samples` backtracked trees per family from host-supplied uniforms; reports the
device-side rate (CUDA events around whale_backtrack's kernels are not exposed, so this is the wall time of the
C-ABI call, which includes uploading the uniforms and downloading the node arrays) next to the oracle's rate on
a sample of the same families."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--families", type=int, default=1000)
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--max-nodes", type=int, default=384)
    args = ap.parse_args()
    import whale_jl_b200 as W
    from whale_jl_b200 import synth, lib as wlib
    from whale_jl_b200.core import _data_handle
    d = synth.cache_dir(f"c5_seed5_n{args.families}")
    synth.generate(d, args.families, seed=5)
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), synth.c1_species_tree(), 0.05)
    ccd = W.read_ale(d, w)
    L = wlib.get()
    mh, dh = _data_handle(w, ccd)
    t0 = time.perf_counter()
    W.logpdf_(w, ccd)
    t_keep = time.perf_counter() - t0
    F, S, MN = len(ccd), args.samples, args.max_nodes
    U = np.random.default_rng(0).random((F, S, 4 * MN))
    L.backtrack(mh, dh, 2, U[:, :2], MN)  # warm-up
    t0 = time.perf_counter()
    cnt, st, nodes = L.backtrack(mh, dh, S, U, MN)
    dt = time.perf_counter() - t0
    assert np.all(st == 0), np.unique(st)
    kms = L.last_backtrack_ms(dh)
    ell_bytes = 8 * sum(L.L.whale_ell_size(dh, f) for f in range(F))
    out = {"metric": "backtracked reconciled trees/s", "value": F * S / dt, "kernel_ms": kms,
           "kernel_trees_per_s": F * S / (kms * 1e-3), "ell_bytes_resident": int(ell_bytes), "families": F, "samples": S,
           "mean_nodes_per_tree": float(cnt.mean()), "call_s": dt, "logpdf_keep_ell_s": t_keep,
           "h2d_bytes": int(U.nbytes), "d2h_bytes": int(nodes.nbytes + cnt.nbytes + st.nbytes)}
    # fused track_and_sum loop (src/track.jl:47-63): a new θ per sample, logpdf! + one walk per family, on the device
    T = min(S, 50)
    rng = np.random.default_rng(1)
    X = w.x()[None, :] * np.exp(0.03 * rng.standard_normal((T, w.n_params)))
    X[:, 2:] = np.clip(X[:, 2:], 1e-3, 1 - 1e-3)
    L.track(mh, dh, X[:2], w.p_leaf(), 1, U[:, :2], MN)  # warm-up
    t0 = time.perf_counter()
    c2, s2, n2, ll2 = L.track(mh, dh, X, w.p_leaf(), 1, U[:, :T], MN)
    dtt = time.perf_counter() - t0
    assert np.all(s2 == 0)
    out["fused_track"] = {"thetas": T, "trees_per_s_call": F * T / dtt, "device_ms": L.last_backtrack_ms(dh),
                          "trees_per_s_device": F * T / (L.last_backtrack_ms(dh) * 1e-3),
                          "note": "every tree from its own posterior draw: logpdf! (keep ℓ) + walk per draw, enqueued back to back"}
    # oracle on a sample
    from oracle import whale_oracle as wo, flat
    ow = wo.WhaleModel(wo.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), wo.c1_tree(), 0.05)
    files = sorted(os.listdir(d))[:8]
    spmap = {n.name: n.id for n in ow.order if n.isleaf()}
    occd = [wo.CCD(wo.parse_aleobserve(os.path.join(d, f)), ow, spmap) for f in files]
    fm, ff = flat.FlatModel(ow), flat.FlatFams(occd, len(ow))
    t0 = time.perf_counter()
    n = 0
    for f in range(len(occd)):
        for s in range(10):
            k, arr, used = flat.backtrack(fm, ff, f, U[f, s], max_nodes=MN)
            assert k == cnt[f, s] and np.array_equal(arr, nodes[f, s, :k]), (f, s)
            n += 1
    out["oracle_trees_per_s_single_thread_incl_dp"] = n / (time.perf_counter() - t0)
    out["parity_checked_trees"] = n
    print(json.dumps(out))


if __name__ == "__main__":
    main()
