#!/usr/bin/env python
"""Backtracking parity at scale: the GPU's reconciled trees against the CPU oracle's for the same uniforms, on >= 10^5
trees (the kept ℓ differs from the oracle's in the last bits — FMA contraction, reduction order — so a decision on a
knife edge could flip: this counts how often it does).

    python tools/check_track_parity.py [--families 1000] [--samples 100]

Prints one JSON line: trees compared, trees that differ, and for those the first differing node."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--families", type=int, default=1000)
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--stride", type=int, default=1024)
    ap.add_argument("--max-nodes", type=int, default=384)
    args = ap.parse_args()
    import whale_jl_b200 as W
    from whale_jl_b200 import synth, lib as wlib
    from whale_jl_b200.core import _data_handle
    from oracle import flat, whale_oracle as wo  # the checker
    d = synth.cache_dir(f"c5_seed5_n{args.families}_shard0of1")
    synth.generate(d, args.families, seed=5)
    pars = dict(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67)
    w = W.WhaleModel(W.ConstantDLWGD(**pars), synth.c1_species_tree(), 0.05)
    ccd = W.read_ale_native(d, w)
    L = wlib.get()
    mh, dh = _data_handle(w, ccd)
    F, S, MN = len(ccd), args.samples, args.max_nodes
    U = np.random.default_rng(123).random((F, S, args.stride))
    L.logpdf_grad(mh, dh, w.x(), w.p_leaf(), 1, keep_ell=True)
    tot = L.backtrack_device(mh, dh, S, U, max_nodes=MN)
    cnt, st = L.trees_counts(dh, F * S)
    off, nodes = L.trees_get(dh, F * S, tot)
    ow = wo.WhaleModel(wo.ConstantDLWGD(**pars), wo.c1_tree(), 0.05)
    spmap = {n.name: n.id for n in ow.order if n.isleaf()}
    files = sorted(f for f in os.listdir(d) if f.endswith(".ale"))
    occd = [wo.CCD(wo.parse_aleobserve(os.path.join(d, f)), ow, spmap) for f in files]
    fm, ff = flat.FlatModel(ow), flat.FlatFams(occd, len(ow))
    x = np.ascontiguousarray(fm.x)

    def one(f):
        bad = []
        for s in range(S):
            n, arr, used = flat.backtrack(fm, ff, f, U[f, s], x=x, max_nodes=MN)
            mine = nodes[off[f * S + s]:off[f * S + s + 1]]
            if st[f * S + s] != 0 or n != len(mine) or not np.array_equal(arr, mine):
                k = 0
                while n > 0 and k < min(n, len(mine)) and np.array_equal(arr[k], mine[k]):
                    k += 1
                bad.append((f, s, int(st[f * S + s]), int(n), int(len(mine)), k))
        return bad

    t0 = time.time()
    with ThreadPoolExecutor(max_workers=len(os.sched_getaffinity(0))) as ex:
        bad = [b for r in ex.map(one, range(F)) for b in r]
    print(json.dumps({"trees_compared": F * S, "families": F, "samples": S, "nodes_total": int(tot), "trees_differing": len(bad),
                      "first_differences": bad[:10], "oracle_s": round(time.time() - t0, 1),
                      "config": "C5 shape: 9-taxon tree + 2 WGD, ConstantDLWGD, ~200 clades, host uniforms (seed 123)"}))


if __name__ == "__main__":
    main()
