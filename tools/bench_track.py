#!/usr/bin/env python
"""Backtracking throughput (BASELINE.json configs[4], "C5": posterior reconciled trees for 10^4 families at 1/2/4/8 GPUs).

    python tools/bench_track.py [--families 10000] [--samples 100] [--thetas 100]
    python -m torch.distributed.run --nproc-per-node N tools/bench_track.py ...     (families sharded over ranks, no collective)

Legs (all through the C ABI, wall clock around the calls = end to end, incl. every host<->device copy):
  walks_device_rng   whale_backtrack_device from ONE kept ℓ, uniforms drawn on the device, trees fetched in compact form
  walks_host_stream  the same with a host-supplied uniform stream (stride 512 doubles per walk: PCIe-bound)
  walks_padded_r1    round 1's whale_backtrack (padded outputs, stride 4*max_nodes) on a subset, for comparison
  track_sample       track_and_sum's loop (src/track.jl:47-63): every (family, sample) draws its own posterior row out of
                     `--thetas`; logpdf! of the families that drew a row + their walks per row
  summary            device-side tree identity hash + per-family dedup (sumtrees)
Prints one JSON line (rank 0)."""
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
    ap.add_argument("--families", type=int, default=10000)
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--thetas", type=int, default=100)
    ap.add_argument("--max-nodes", type=int, default=384)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from whale_jl_b200 import synth
    lo, hi = args.families * rank // world, args.families * (rank + 1) // world
    d = synth.cache_dir(f"c5_seed5_n{args.families}_shard{rank}of{world}")
    t0 = time.time()
    synth.generate(d, hi - lo, seed=5, first=lo, workers=max(1, min(64, len(os.sched_getaffinity(0)) // world)))
    gen_s = time.time() - t0
    import whale_jl_b200 as W
    from whale_jl_b200 import lib as wlib
    from whale_jl_b200.core import _data_handle
    L = wlib.get()
    L.check(L.L.whale_set_device(local))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), synth.c1_species_tree(), 0.05)
    ccd = W.read_ale_native(d, w)
    mh, dh = _data_handle(w, ccd)
    F, S, MN = len(ccd), args.samples, args.max_nodes
    Wn = F * S
    x0, pl = w.x(), w.p_leaf()

    def sync_all():
        if dist is not None:
            dist.barrier()

    out = {"families_total": args.families, "families_rank0": F, "samples": S, "n_gpus": world, "gen_s": round(gen_s, 1)}
    # ---- logpdf! (keep ℓ) ----
    L.logpdf_grad(mh, dh, x0, pl, 1, keep_ell=True)
    t0 = time.perf_counter()
    L.logpdf_grad(mh, dh, x0, pl, 1, keep_ell=True)
    out["logpdf_keep_ell_ms"] = 1e3 * (time.perf_counter() - t0)
    # HBM roofline of the ℓ-keeping mode: the evaluation streams every ℓ row it forms to HBM once and reads the
    # packed arena once; kernel time from the library's CUDA events (WHALE_PROFILE)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ell_bytes = 8 * int(sum(L.L.whale_ell_size(dh, f) for f in range(F)))
    arena = int(L.L.whale_data_arena_bytes(dh))
    kd = []
    for _ in range(3):
        L.logpdf_grad(mh, dh, x0, pl, 1, keep_ell=True, profile=True)
        kd.append(L.last_kernel_ms(dh)[1])
    kdp = float(np.mean(kd))
    out["keep_ell_roofline"] = {"bound": "hbm", "kernel": "k_dp (value plan, KEEP_ELL)", "ell_bytes_written": ell_bytes,
                                "arena_bytes_read": arena, "k_dp_ms": kdp, "achieved": (ell_bytes + arena) / (kdp * 1e-3) / 1e9,
                                "peak": hbm_peak, "unit": "GB/s", "frac": (ell_bytes + arena) / (kdp * 1e-3) / 1e9 / hbm_peak,
                                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}
    # ---- walks from one kept ℓ, device stream, compact fetch ----
    L.backtrack_device(mh, dh, S, None, seed=1, max_nodes=MN)
    L.trees_view(dh, Wn)
    sync_all()
    reps = 3
    t0 = time.perf_counter()
    for r in range(reps):
        tot = L.backtrack_device(mh, dh, S, None, seed=2 + r, max_nodes=MN)
        off, nodes = L.trees_view(dh, Wn)
    dt = (time.perf_counter() - t0) / reps
    cnt, st = L.trees_counts(dh, Wn)
    out["walks_device_rng"] = {"trees_per_s_rank": Wn / dt, "ms": 1e3 * dt, "kernel_ms": L.last_backtrack_ms(dh),
                               "nodes_per_tree": tot / Wn, "d2h_bytes": int(16 * tot + 8 * (Wn + 1) + 8 * Wn),
                               "failed": int((st != 0).sum())}
    # k_backtrack against the HBM roofline: a walk reads ℓ cells scattered over its family's kept matrix (32-byte sectors,
    # mostly L2 hits: S walks share one family's ℓ) and writes 16 bytes per tree node; the algorithmic floor is one read of
    # the kept ℓ + the node records.  DRAM bytes measured by ncu: profiles/r2_ncu_k_backtrack_summary.txt
    kb_ms = float(L.last_backtrack_ms(dh))
    out["backtrack_roofline"] = {"bound": "hbm", "kernel": "k_backtrack", "kernel_ms": kb_ms,
                                 "algorithmic_bytes": int(ell_bytes + 16 * tot), "achieved": (ell_bytes + 16 * tot) / (kb_ms * 1e-3) / 1e9,
                                 "peak": hbm_peak, "unit": "GB/s", "frac": (ell_bytes + 16 * tot) / (kb_ms * 1e-3) / 1e9 / hbm_peak}
    # ---- device summary ----
    L.trees_summary(dh, F, S)  # (first call allocates)
    t0 = time.perf_counter()
    nd, h, c, f1, _ = L.trees_summary(dh, F, S)
    out["summary"] = {"ms": 1e3 * (time.perf_counter() - t0), "distinct_trees_per_family_mean": float(nd.mean())}
    # ---- host-supplied stream ----
    stride = 512
    rng = np.random.default_rng(3 + rank)
    nf_h = min(F, 2000)
    # (a handle over the first nf_h families would need a second pack: time the full batch with a stream only when it fits)
    if Wn * stride * 8 <= 6 << 30:
        U = rng.random((F, S, stride))
        L.backtrack_device(mh, dh, S, U, max_nodes=MN)
        t0 = time.perf_counter()
        tot = L.backtrack_device(mh, dh, S, U, max_nodes=MN)
        L.trees_view(dh, Wn)
        dt = time.perf_counter() - t0
        cnt, st = L.trees_counts(dh, Wn)
        out["walks_host_stream"] = {"trees_per_s_rank": Wn / dt, "ms": 1e3 * dt, "h2d_bytes": int(U.nbytes), "stride": stride,
                                    "exhausted_or_failed": int((st != 0).sum())}
        del U
    # ---- track_and_sum's loop: a posterior row per (family, sample) ----
    nt = args.thetas
    X = x0[None, :] * np.exp(0.05 * rng.standard_normal((nt, len(x0))))
    X[:, 2:] = np.clip(X[:, 2:], 1e-3, 1 - 1e-3)
    ti = rng.integers(nt, size=(F, S)).astype(np.int32)
    L.track_sample(mh, dh, X, pl, S, ti, None, seed=9, max_nodes=MN)
    sync_all()
    t0 = time.perf_counter()
    tot = L.track_sample(mh, dh, X, pl, S, ti, None, seed=10, max_nodes=MN)
    off, nodes = L.trees_view(dh, Wn)
    dt = time.perf_counter() - t0
    cnt, st = L.trees_counts(dh, Wn)
    out["track_sample"] = {"trees_per_s_rank": Wn / dt, "ms": 1e3 * dt, "device_ms": L.last_backtrack_ms(dh), "thetas": nt,
                           "family_evaluations": int(sum(len(np.unique(np.nonzero(ti == j)[0])) for j in range(nt))),
                           "failed": int((st != 0).sum())}
    if dist is not None:
        import torch
        keys = [("walks_device_rng", "ms"), ("track_sample", "ms")]
        t = torch.tensor([out[a][b] for a, b in keys], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for (a, b), v in zip(keys, t.tolist()):
            out[a]["ms_max_over_ranks"] = v
            out[a]["trees_per_s_total"] = args.families * S / (v * 1e-3)
        dist.destroy_process_group()
    else:
        for a in ("walks_device_rng", "track_sample"):
            out[a]["trees_per_s_total"] = out[a]["trees_per_s_rank"]
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
