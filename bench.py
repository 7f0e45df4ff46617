#!/usr/bin/env python
"""bench.py — family (log-likelihood + gradient) evaluations per second on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--c3-families F]

Headline workload (BASELINE.json configs[1], "C2"): per GPU 1000 synthetic families x ~200 clades on the reference's
9-taxon tree with 2 WGD nodes (Whale.extree + wgd_1/wgd_2, test/runtests.jl:8-11), ConstantDLWGD rates
(P = 5: λ, μ, q1, q2, η), Δt = 0.05, RootCondition.  One *step* = one logpdf+∇ evaluation of every family
for a NEW parameter vector (NUTS/MLE style: slice tables are recomputed each step).  Families shard across
ranks (weak scaling: 1000 families per GPU); the only collective is the sum of 1+P doubles.

Second leg in the same run (BASELINE.json configs[2], the north star's scaling configuration, "C3"): `--c3-families`
(default 100000) synthetic families in TOTAL, branch-wise rates (DLWGD, P = 37), STRONG scaling: rank r holds the
families [r·F/N, (r+1)·F/N), one MLE gradient step = every rank's reverse-mode evaluation + the same all-reduce.
Reported under the key "c3_strong" (device-timed and end-to-end); the headline `value` stays the C2 metric.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with the arena resident in HBM (CUDA events
per step, L2 flushed between steps, max over ranks); `e2e` = the same metric through the C-ABI call with HOST
buffers (θ H2D and loglik+grad D2H inside the timed region); `roofline` = the DP kernel against the measured
fp64 FMA peak (the DP never leaves shared memory), `roofline_hbm` the same launch against
MEASURED_PEAKS.json's HBM bandwidth; `cpu_baseline` = the C++ oracle port on this box's host cores (all of them,
set explicitly: torchrun exports OMP_NUM_THREADS=1).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "family loglik+grad evals/s"
UNIT = "family-evals/s"
BASE_X = np.array([0.2, 0.3, 0.2, 0.1, 0.67])  # λ, μ, q1, q2, η  
DT = 0.05
FAMILIES_PER_GPU = 1000
SEED = 2


def thetas(n, seed=0):
    """A new θ per step (like successive NUTS/MLE iterates): 3 % log-normal jitter around the C2 point."""
    rng = np.random.default_rng(1000 + seed)
    x = BASE_X[None, :] * np.exp(0.03 * rng.standard_normal((n, len(BASE_X))))
    x[:, 2:] = np.clip(x[:, 2:], 1e-3, 1 - 1e-3)
    return np.ascontiguousarray(x)


def host_threads():
    """Host cores this process may use (the CPU arms set their OpenMP thread count to this explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def dataset(rank, n_fam):
    from whale_jl_b200 import synth
    d = synth.cache_dir(f"c2_seed{SEED}_rank{rank}_n{n_fam}")
    t0 = time.time()
    synth.generate(d, n_fam, seed=SEED + 7919 * rank, workers=min(host_threads(), 16))
    return d, time.time() - t0


C3_SEED = 3


def c3_dataset(rank, world, total):
    """This rank's contiguous shard of the C3 families (strong scaling: `total` families over `world` ranks)."""
    from whale_jl_b200 import synth
    lo, hi = total * rank // world, total * (rank + 1) // world
    d = synth.cache_dir(f"c3_seed{C3_SEED}_n{total}_shard{rank}of{world}")
    t0 = time.time()
    synth.generate(d, hi - lo, seed=C3_SEED, first=lo, workers=max(1, min(64, host_threads() // max(1, world))))
    return d, hi - lo, time.time() - t0


def c3_model():
    """9-taxon tree + 2 WGDs, branch-wise rates on the log scale (DLWGD, P = 2·17 + 2 + 1 = 37; SURVEY §8d: log-rates
    ~ N(log 0.15, 0.3²))."""
    import whale_jl_b200 as W
    rng = np.random.default_rng(7)
    r = W.DLWGD(lam=list(rng.normal(np.log(0.15), 0.3, 17)), mu=list(rng.normal(np.log(0.15), 0.3, 17)), q=[0.2, 0.1], eta=0.67)
    return W.WhaleModel(r, W.synth.c1_species_tree(), DT)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _oracle_c2(ale_dir, limit=None):
    from oracle import flat, whale_oracle as wo
    ow = wo.WhaleModel(wo.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), wo.c1_tree(), DT)
    files = sorted(f for f in os.listdir(ale_dir) if f.endswith(".ale"))[:limit]
    spmap = {n.name: n.id for n in ow.order if n.isleaf()}
    ccds = [wo.CCD(wo.parse_aleobserve(os.path.join(ale_dir, f)), ow, spmap) for f in files]
    return flat, flat.FlatModel(ow), flat.FlatFams(ccds, len(ow)), len(ccds)


def cpu_baseline(ale_dir, budget_s=12.0):
    """The oracle port (C++ restatement of the reference loops, OpenMP over families like Threads.@threads,
    src/core.jl:58-64) timed on ALL families of the step, logpdf + ForwardDiff-style gradient, on every host core
    (thread count set explicitly, not taken from OMP_NUM_THREADS)."""
    flat, fm, ff, n_fam = _oracle_c2(ale_dir)
    nt = host_threads()
    xs = thetas(64, seed=99)
    flat.logpdf(fm, ff, x=xs[0], grad=True, nthreads=nt)  # warm-up
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s:
        flat.logpdf(fm, ff, x=xs[n % len(xs)], grad=True, nthreads=nt)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n_fam * n / dt, "unit": UNIT, "cores": nt, "kind": "port",
            "sample": f"all {n_fam} C2 families x {n} evaluations (logpdf + Dual<5> gradient), {nt} OpenMP threads "
                      f"(set explicitly; OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')}); "
                      f"JULIA_NUM_THREADS n/a (no julia in this image)"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Julia original cannot run here) on the
    same config — every family of the step, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, _ = dataset(0, FAMILIES_PER_GPU)
    flat, fm, ff, S = _oracle_c2(d)
    nt = host_threads()
    xs = thetas(args.steps + args.warmup)
    for i in range(args.warmup):
        flat.logpdf(fm, ff, x=xs[i], grad=True, nthreads=nt)
    t0 = time.perf_counter()
    for i in range(args.steps):
        flat.logpdf(fm, ff, x=xs[args.warmup + i], grad=True, nthreads=nt)
    dt = time.perf_counter() - t0
    val = S * args.steps / dt
    sample = (f"all {S} C2 families per step, logpdf + Dual<5> gradient, {nt} OpenMP threads (set explicitly; "
              f"OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')})")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 1000 synthetic families x ~200 clades, 9-taxon tree + 2 WGD, ConstantDLWGD P=5, dt=0.05"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nt, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def setup_exchange(L, dh, rank, world):
    """Sum over ranks: the library's one-shot peer-memory exchange (whale_peer_*: every rank stores its 1+P doubles into
    every peer's buffer over NVLink and adds them in rank order) when CUDA IPC works between the ranks, else an NCCL
    all-reduce (WHALE_BENCH_EXCHANGE=nccl forces that).  Returns True for the peer exchange."""
    if world == 1:
        return False
    import torch
    import torch.distributed as dist
    ok, h = 1, bytes(64)
    if os.environ.get("WHALE_BENCH_EXCHANGE", "peer") != "peer":
        ok = 0
    else:
        try:
            h = L.peer_export(dh, rank, world)
        except Exception:
            ok = 0
    t = torch.tensor(list(h), dtype=torch.uint8, device="cuda")
    allh = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allh, t)
    if ok:
        try:
            for q in range(world):
                if q != rank:
                    L.peer_import(dh, q, bytes(allh[q].cpu().tolist()))
        except Exception as exc:
            print(f"[bench] rank {rank}: peer import failed ({exc}); NCCL all-reduce instead", file=sys.stderr)
            ok = 0
    f = torch.tensor([ok], device="cuda", dtype=torch.int32)
    dist.all_reduce(f, op=dist.ReduceOp.MIN)
    return bool(int(f.item()))


def c3_leg(args, L, wlib, rank, world, d3, n3, gen3_s, flush, stream):
    """BASELINE.json configs[2] (the north star's scaling configuration): `--c3-families` families in total, branch-wise
    rates (P = 37), strong scaling over the ranks, one MLE gradient step = reverse-mode evaluation of the rank's shard +
    all-reduce of 1+P doubles.  Device-timed (CUDA events, L2 flushed, max over ranks) and end to end (host θ in,
    loglik+∇ out)."""
    import torch
    import torch.distributed as dist
    import whale_jl_b200 as W
    from whale_jl_b200.core import _data_handle
    model = c3_model()
    t0 = time.time()
    ccds = W.read_ale_native(d3, model)
    mh, dh = _data_handle(model, ccds)
    pack_s = time.time() - t0
    P = model.n_params
    K, Wm = max(3, min(args.steps, 12)), 3
    rng = np.random.default_rng(11)
    x0 = model.x()
    xs = x0[None, :] + 0.02 * rng.standard_normal((K + Wm, P))
    xs[:, -3:] = np.clip(xs[:, -3:], 1e-3, 1 - 1e-3)
    X = torch.tensor(xs, device="cuda", dtype=torch.float64)
    OUT = torch.zeros(1 + P, device="cuda", dtype=torch.float64)
    cond = 1
    peer = setup_exchange(L, dh, rank, world)
    PS = wlib.PEER_SUM if peer else 0

    def step(i, flags):
        L.logpdf_grad_async(mh, dh, X[i].data_ptr(), cond, flags | PS, OUT.data_ptr(), stream)
        if world > 1 and not peer:
            dist.all_reduce(OUT)

    for i in range(Wm):
        step(i, wlib.WANT_GRAD)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        step(Wm + i, wlib.WANT_GRAD)
        ev[i][1].record()
    torch.cuda.synchronize()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    last = OUT.cpu().numpy().copy()
    kms = np.zeros((3, 3))
    for i in range(3):
        flush.zero_()
        step(Wm + i, wlib.WANT_GRAD | wlib.PROFILE)
        torch.cuda.synchronize()
        kms[i] = L.last_kernel_ms(dh)
    pin_x = torch.empty(P, dtype=torch.float64).pin_memory()
    pin_o = torch.empty(1 + P, dtype=torch.float64).pin_memory()
    Xd = torch.empty(P, device="cuda", dtype=torch.float64)

    def e2e_step(i):
        if world == 1 or peer:
            return L.logpdf_grad(mh, dh, xs[i], model.p_leaf(), cond, want_grad=True, peer_sum=peer)[0]
        pin_x.copy_(torch.from_numpy(xs[i]))
        Xd.copy_(pin_x, non_blocking=True)
        L.logpdf_grad_async(mh, dh, Xd.data_ptr(), cond, wlib.WANT_GRAD | PS, OUT.data_ptr(), stream)
        if not peer:
            dist.all_reduce(OUT)
        pin_o.copy_(OUT, non_blocking=True)
        torch.cuda.synchronize()
        return float(pin_o[0])

    for i in range(Wm):  # (the third call with the same flags is the first CUDA-graph replay: keep the capture out of the timing)
        e2e_step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        ll = e2e_step(Wm + i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    flops, abytes = L.work_estimate(mh, dh, True)
    arena = int(L.L.whale_data_arena_bytes(dh))
    mode = "reverse" if L.L.whale_data_grad_mode(dh) == 1 else "forward"
    passes = int(L.L.whale_data_grad_passes(dh))
    tot = torch.tensor([total_ms, e2e_s, float(n3)], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_s, fam_total = float(mx[0]), float(mx[1]), int(round(float(sm[2])))
    else:
        fam_total = n3
    ccds.close()  # releases the shard's device arena (whale_data_destroy)
    assert np.isfinite(ll) and abs(ll - last[0]) <= 1e-9 * abs(last[0]), (ll, last[0])
    dp_ms = float(kms[:, 1].mean())
    return {"workload": f"C3: {fam_total} synthetic families in total (strong scaling, {n3} on rank 0), 9-taxon tree + 2 WGD, "
                        f"DLWGD branch-wise rates P={P}, dt={DT}, RootCondition, new theta each step",
            "families_total": fam_total, "P": P, "grad_mode": mode, "gradient_passes": passes, "steps": K, "warmup": Wm,
            "exchange": ("peer-memory packet exchange fused into the tail of k_dp_rev" if peer else "NCCL all-reduce") if world > 1 else "none",
            "value": fam_total * K / (total_ms * 1e-3), "unit": UNIT, "ms_per_step": total_ms / K, "scaling": "strong",
            "e2e": {"value": fam_total * K / e2e_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / K,
                    "h2d_bytes_per_step": 8 * (P + model.nn), "d2h_bytes_per_step": 8 * (1 + P)},
            "kernels_ms_rank0": {"k_tables": float(kms[:, 0].mean()), "k_dp": dp_ms, "k_reduce": float(kms[:, 2].mean())},
            "algorithmic_tflops_rank0": flops / (dp_ms * 1e-3) / 1e12,
            "arena_bytes_rank0": arena, "gen_s": round(gen3_s, 1), "pack_s": round(pack_s, 1), "loglik_last": float(last[0])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--families", type=int, default=FAMILIES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c3-families", type=int, default=100000,
                    help="total families of the strong-scaling C3 leg (BASELINE configs[2]); 0 disables the leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # synthetic inputs first (worker processes are forked here, before this process touches CUDA)
    d, gen_s = dataset(0 if os.environ.get("WHALE_BENCH_SAME_DATA") else rank, args.families)  # (diagnosis: every rank rank 0's shard)
    d3 = n3 = gen3_s = None
    if args.c3_families > 0:
        d3, n3, gen3_s = c3_dataset(rank, world, args.c3_families)

    import torch
    import torch.distributed as dist
    import whale_jl_b200 as W
    from whale_jl_b200 import lib as wlib

    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    if world > 1:
        import datetime
        # a short collective timeout: a rank that falls out of step fails the run in two minutes instead of ten
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    L = wlib.get()
    L.check(L.L.whale_set_device(local))

    # ---- data: this rank's shard (generated above), read_ale, pack once ----
    model = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), W.synth.c1_species_tree(), DT)
    t0 = time.time()
    ccds = W.read_ale_native(d, model)  # whale_read_ale: parse + CCD construction + packing in the library
    from whale_jl_b200.core import _data_handle
    mh, dh = _data_handle(model, ccds)
    pack_s = time.time() - t0
    F, P = len(ccds), model.n_params
    cond = 1  # RootCondition
    flops, abytes = L.work_estimate(mh, dh, True)
    arena_bytes = L.L.whale_data_arena_bytes(dh)

    K, Wm = args.steps, args.warmup
    xs = thetas(K + Wm, seed=rank * 0)  # same θ on every rank (one model, sharded data)
    X = torch.tensor(xs, device="cuda", dtype=torch.float64)
    OUT = torch.zeros(1 + P, device="cuda", dtype=torch.float64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream
    peer = setup_exchange(L, dh, rank, world)
    PS = wlib.PEER_SUM if peer else 0
    FLAGS = wlib.WANT_GRAD | wlib.PROFILE | PS

    def step(i, flags=FLAGS):
        L.logpdf_grad_async(mh, dh, X[i].data_ptr(), cond, flags, OUT.data_ptr(), stream)
        if world > 1 and not peer:
            dist.all_reduce(OUT)

    for i in range(Wm):
        step(i)
    torch.cuda.synchronize()
    launches0 = L.L.whale_launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # The K timed steps are enqueued back to back (one CUDA event pair per step around the evaluation + exchange, the L2
    # flush between steps outside the pairs) and synchronised ONCE at the end: a host round trip per step would let the
    # ranks drift apart and show up as waiting time inside the exchange.  Per-kernel times come from extra profiled steps.
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()  # evict the arena from L2 (outside the timed event pair)
        ev[i][0].record()
        step(Wm + i, wlib.WANT_GRAD | PS)
        ev[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = L.L.whale_launch_count() - launches0
    last = OUT.cpu().numpy().copy()  # result of the last timed step (compared with the end-to-end leg's below)
    NK = min(K, 20)
    kms = np.zeros((NK, 3))
    for i in range(NK):
        flush.zero_()
        step(Wm + i)  # WHALE_PROFILE: events around each kernel
        torch.cuda.synchronize()
        kms[i] = L.last_kernel_ms(dh)
    phase_cycles = L.last_phase_cycles(dh)
    tables_cycles = L.last_tables_cycles(mh, True)
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(step_ms.sum())
    # keep the GPU under the same load a little longer if the region was too short for nvidia-smi to sample
    # (rank-local work only: the number of rounds differs between ranks, so no collective may be issued here)
    t_probe = time.perf_counter()
    while len(sampler.rows) < 8 and time.perf_counter() - t_probe < 3.0:
        for i in range(Wm):  # (no exchange here: the number of rounds differs between ranks)
            L.logpdf_grad_async(mh, dh, X[i].data_ptr(), cond, wlib.WANT_GRAD, OUT.data_ptr(), stream)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = F * world * K / (total_ms * 1e-3)
    # per-rank view of the same steps (who waits for whom) and the exchange timed alone, back to back
    ranks = None
    if world > 1:
        mine = torch.tensor([float(step_ms.mean()), *kms.mean(axis=0).tolist(), float(clocks["sm_mhz"] or 0),
                             float(len(clocks["reasons"]))], device="cuda", dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        Z = torch.zeros(1 + P, device="cuda", dtype=torch.float64)
        NX = 200
        evx = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(NX)]
        for rep in range(2):  # first round: warm-up
            dist.barrier()
            torch.cuda.synchronize()
            for i in range(NX):
                evx[i][0].record()
                if peer:
                    L.check(L.L.whale_peer_sum_async(dh, Z.data_ptr(), stream))
                else:
                    dist.all_reduce(Z)
                evx[i][1].record()
            torch.cuda.synchronize()
        xus = np.array([a.elapsed_time(b) for a, b in evx]) * 1e3
        r = torch.stack(allr).cpu().numpy()
        ranks = {"step_events_ms": r[:, 0].tolist(), "k_tables_ms": r[:, 1].tolist(), "k_dp_ms": r[:, 2].tolist(),
                 "sm_mhz": r[:, 4].tolist(), "throttle_reasons": r[:, 5].tolist(),
                 "exchange_alone_us_rank0": {"median": float(np.median(xus)), "p90": float(np.percentile(xus, 90)),
                                             "note": "the exchange kernel (or the NCCL all-reduce) enqueued 200 times back to "
                                                     "back with nothing else in the stream, CUDA event pair around each"}}

    # ---- e2e: the reference-facing call with HOST buffers ----
    pin_x = torch.empty(P, dtype=torch.float64).pin_memory()
    pin_o = torch.empty(1 + P, dtype=torch.float64).pin_memory()
    Xd = torch.empty(P, device="cuda", dtype=torch.float64)

    def e2e_step(i):
        if world == 1 or peer:  # exactly the call the Julia glue makes: host pointers in, host pointers out (N > 1: + WHALE_PEER_SUM)
            return L.logpdf_grad(mh, dh, xs[i], model.p_leaf(), cond, want_grad=True, peer_sum=peer)[0]
        pin_x.copy_(torch.from_numpy(xs[i]))
        Xd.copy_(pin_x, non_blocking=True)
        L.logpdf_grad_async(mh, dh, Xd.data_ptr(), cond, wlib.WANT_GRAD | PS, OUT.data_ptr(), stream)
        if not peer:
            dist.all_reduce(OUT)
        pin_o.copy_(OUT, non_blocking=True)
        torch.cuda.synchronize()
        return float(pin_o[0])

    for i in range(Wm):
        e2e_step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        ll_e2e = e2e_step(Wm + i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = F * world * K / e2e_s
    assert np.isfinite(ll_e2e) and abs(ll_e2e - last[0]) <= 1e-9 * abs(last[0]), (ll_e2e, last[0])

    # ---- second leg: the north star's scaling configuration (C3, strong scaling) ----
    c3 = None
    if d3 is not None:
        try:
            c3 = c3_leg(args, L, wlib, rank, world, d3, n3, gen3_s, flush, stream)
        except Exception as exc:  # the headline line must survive a failure of the second leg
            c3 = {"error": f"{type(exc).__name__}: {exc}"}
            if world > 1:
                raise

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_dp) ----
    dp_ms = float(kms[:, 1].mean())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    fp64_peak = L.fp64_peak()  # measured DFMA microbenchmark on this GPU (MEASURED_PEAKS.json has no fp64 entry)
    # DRAM traffic of the k_dp launches of one step, from the committed `ncu --set full` capture of this command
    # (profiles/r2_ncu_k_dp_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum over the step's k_dp launches)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_k_dp_traffic.json")))
        if int(tj.get("families", 0)) == F:
            traffic, traffic_src = float(tj["dram_bytes_per_step"]), tj.get("source")
    except (OSError, ValueError, KeyError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ach_tf = flops / (dp_ms * 1e-3) / 1e12
    ach_gb = abytes / (dp_ms * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2: {F} synthetic families/GPU x ~200 clades (median), 9-taxon tree + 2 WGD "
                               f"(19 nodes, {int(model.n_slices.sum())} slices), ConstantDLWGD P={P}, dt={DT}, "
                               f"RootCondition, new theta each step",
                   "families_total": F * world,
                   "sharding": f"families/{world} ranks, sum of {1 + P} f64 per step: " +
                               ("peer-memory packet exchange in the library (CUDA IPC / NVLink), fused into the tail of the DP kernel" if peer
                                else "NCCL all-reduce" if world > 1 else "single rank"),
                   "l2": "flushed between steps (256 MiB memset outside the timed event pairs)",
                   "timing": "K steps enqueued back to back, one CUDA event pair per step on the launching stream, one "
                             "synchronize at the end, max over ranks of the summed pairs",
                   "arena_bytes_per_gpu": int(arena_bytes), "gen_s": round(gen_s, 1), "pack_s": round(pack_s, 2)},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 8 * (P + model.nn),
                "d2h_bytes_per_step": 8 * (1 + P), "ms_per_step": 1e3 * e2e_s / K,
                "note": "whale_logpdf_grad with host pointers; the packed CCD arena stays resident in HBM like "
                        "the reference's CCD objects stay resident in RAM across logpdf calls"},
        "gpu_launches": int(launches),
        "kernels_ms": {"k_tables": float(kms[:, 0].mean()), "k_dp": dp_ms, "k_reduce": float(kms[:, 2].mean()),
                       "k_reduce_note": "the reduction runs in the tail of k_dp; this is the gap to the step's end event",
                       "step_events": float(step_ms.mean()), "wall_per_step_incl_flush": 1e3 * wall / K},
        "roofline": {"bound": "fp64",
                     "bound_note": "BASELINE.json's metric asks for the fp64 (CUDA-core FMA) roofline: the DP is an fp64 "
                                   "gather-multiply-accumulate that neither streams HBM (see roofline_hbm) nor maps to "
                                   "tensor cores",
                     "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": ach_tf / fp64_peak, "traffic": traffic, "traffic_source": traffic_src, "kernel": "k_dp<128,4>",
                     "flops_per_launch": flops, "peak_source": "whale_fp64_peak DFMA microbenchmark, this GPU, this run"},
        "roofline_hbm": {"bound": "hbm", "achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s",
                         "frac": ach_gb / hbm_peak, "bytes_per_launch": abytes,
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        "dp_phase_cycles_mean_max": phase_cycles, "tables_cycles": tables_cycles,
        "loglik_last": float(last[0]),
        "grad_mode": "reverse" if L.L.whale_data_grad_mode(dh) == 1 else "forward",
    }
    if ranks is not None:
        out["ranks"] = ranks
    if c3 is not None:
        out["c3_strong"] = c3
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(d)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
