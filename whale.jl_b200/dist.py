"""Family sharding across ranks (one process per GPU) — the only parallel axis of the path.

Families are i.i.d. terms of a sum (src/core.jl:54,63), so every rank packs and evaluates its own shard and the
ranks exchange nothing but the sum of `1 + P` doubles (log-likelihood and gradient) per evaluation.  Each shard
subtracts its own `n_fam·condition`, so the all-reduced vector is the full-batch result.
"""
from __future__ import annotations

import numpy as np


def shard(work, rank: int, world: int) -> np.ndarray:
    """Indices of the families owned by `rank`: longest-processing-time assignment on the predicted work
    (clades per family span 15..1025 in the reference's own fixtures), deterministic on every rank."""
    work = np.asarray(work, dtype=np.float64)
    order = np.argsort(-work, kind="stable")
    load = np.zeros(world)
    owner = np.empty(len(work), np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += work[i]
    return np.flatnonzero(owner == rank)


def predicted_work(ccd, n_slices) -> float:
    """Σ_e n_e·(T_e + C_e): triple evaluations + cell updates of one evaluation."""
    nsp = np.diff(ccd.split_off)
    return float(sum(int(n_slices[e]) * (int(nsp[comp].sum()) + len(comp)) for e, comp in enumerate(ccd.compat)))


def allreduce_sum(t):
    """Sum the (1+P)-vector over ranks in place (NCCL on GPU tensors, gloo on CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t
