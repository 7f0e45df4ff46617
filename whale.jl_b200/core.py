"""`logpdf` / `logpdf!` / fused gradient — the reference-facing operator surface (src/core.jl:17-81).

Every function here is a thin driver of the C ABI: pack once (`whale_model_create`, `whale_data_create`),
then one `whale_logpdf_grad` call per evaluation.  No arithmetic of the hot path happens in Python.
"""
from __future__ import annotations

import numpy as np

from . import lib as _lib
from .ccd import CCD, CCDVector, NativeCCDVector
from .model import CONDITIONS, WhaleModel


def _model_handle(wm: WhaleModel):
    L = _lib.get()
    if wm._handle is None or wm._handle[0] is not L:
        wm._handle = (L, L.model_create(wm))
    return wm._handle[1]


def _data_handle(wm: WhaleModel, xs: CCDVector):
    L = _lib.get()
    mh = _model_handle(wm)
    key = (id(L), mh)
    if key not in xs._data:
        if isinstance(xs, NativeCCDVector):
            raise ValueError("a natively read batch belongs to the model it was read with (same structure handle)")
        xs._data[key] = L.data_create(mh, xs.flatten(wm.nn))
    return mh, xs._data[key]


_LIST_BATCHES: "dict[tuple, CCDVector]" = {}  # plain lists of CCDs seen recently -> their batch (keeps the device arena)


def _as_vector(x) -> tuple[CCDVector, bool]:
    if isinstance(x, NativeCCDVector):
        return x, False
    if isinstance(x, CCD):
        if x._batch is None:
            x._batch = CCDVector([x])
        return x._batch, True
    if not isinstance(x, CCDVector):
        # a plain list / tuple of CCDs (e.g. built by a comprehension): identified by its members, so that an optimisation
        # loop calling logpdf(wm, [c1, c2, ...]) packs the device arena once, not on every call (a small LRU: the evicted
        # batch releases its handles)
        key = tuple(id(c) for c in x)
        hit = _LIST_BATCHES.pop(key, None)
        if hit is None or len(hit) != len(key) or any(a is not b for a, b in zip(hit, x)):
            hit = CCDVector(x)
        _LIST_BATCHES[key] = hit
        while len(_LIST_BATCHES) > 8:
            _LIST_BATCHES.pop(next(iter(_LIST_BATCHES))).close()
        x = hit
    return x, False


def _eval(wm: WhaleModel, x, grad=False, keep=False, per_family=False, per_family_grad=False):
    xs, single = _as_vector(x)
    if len(xs) == 0:  # an empty Vector{CCD}: Σ over nothing − 0·condition (src/core.jl:54,63); nothing to pack or launch
        P = wm.n_params
        return (0.0, np.zeros(P) if grad else None, np.zeros(0) if per_family else None,
                np.zeros((0, P)) if per_family_grad else None)
    mh, dh = _data_handle(wm, xs)
    # a single CCD is the *unconditioned* likelihood (src/core.jl:29-43); vectors subtract N·condition (:46-64)
    cond = 0 if single else CONDITIONS[wm.condition]
    return _lib.get().logpdf_grad(mh, dh, wm.x(), wm.p_leaf(), cond, want_grad=grad, keep_ell=keep,
                                  per_family=per_family, per_family_grad=per_family_grad)


def logpdf(wm: WhaleModel, x) -> float:
    """`logpdf(wm, ccd)` / `logpdf(wm, ccds)` (src/core.jl:39-43,58-64)."""
    return _eval(wm, x)[0]


def logpdf_(wm: WhaleModel, x) -> float:
    """`logpdf!(wm, ccd(s))` (src/core.jl:29,46-56): same value, and the full ℓ stays on the device for
    `backtrack` / `ell`."""
    return _eval(wm, x, keep=True)[0]


loglikelihood = logpdf  # src/core.jl:17,81


def logpdf_and_gradient(wm: WhaleModel, x):
    """One fused pass: (ℓ, ∂ℓ/∂raw) with raw = [λ…, μ…, q…, η] on the rates struct's scale — what
    `ForwardDiff.gradient(x -> logpdf(wm(rates(x)), ccd), raw)` returns (test/runtests.jl:36-38)."""
    l, g, _, _ = _eval(wm, x, grad=True)
    return l, g


def logpdf_per_family(wm: WhaleModel, xs, grad=False):
    """Unconditioned per-family log-likelihoods (and gradients) — the building block of the mixture /
    model-array variants (src/core.jl:66-79)."""
    l, g, lf, gf = _eval(wm, xs, grad=grad, per_family=True, per_family_grad=grad)
    return lf, gf


def ell(wm: WhaleModel, xs, fam: int = 0):
    """ℓ matrices of one family after `logpdf_` (ccd.ℓ, src/ccd.jl:59-75): list over nodes (index order)
    of (n_e+1) × C_e arrays."""
    xs, _ = _as_vector(xs)
    mh, dh = _data_handle(wm, xs)
    flat = _lib.get().ell_get(dh, fam)
    out, o = [], 0
    for e in range(wm.nn):
        r, c = int(wm.n_slices[e]) + 1, len(xs[fam].compat[e])
        out.append(flat[o:o + r * c].reshape(r, c))
        o += r * c
    return out


def slices(wm: WhaleModel):
    """Per-node slice tables [ϵ ϕ ψ] computed on the device (src/model.jl:182-191): list over nodes of
    (n_e+1) × 3 arrays."""
    mh = _model_handle(wm)
    eps, phi, psi = _lib.get().slices(mh, wm.x(), wm.p_leaf(), int(wm.row_off[-1]))
    return [np.stack([eps[a:b], phi[a:b], psi[a:b]], axis=1) for a, b in zip(wm.row_off[:-1], wm.row_off[1:])]


class BacktrackFailed(RuntimeError):
    """`error("Backtracking failed, ...")` (src/track.jl:150)."""


def backtrack(wm: WhaleModel, x, n_samples: int = 1, uniforms=None, seed=None, max_nodes: int = 512):
    """`backtrack(wm, ccd)` / `backtrack(wm, ccds)` (src/track.jl:188-194): sample reconciled trees from the ℓ
    left on the device by `logpdf_(wm, x)`.  The reference draws from the global RNG; here the uniform stream
    is explicit (`uniforms[F, n_samples, stride]`, or generated from `seed`), consumed in the reference's order.

    Returns, per family, a list of `n_samples` int arrays of shape (n_nodes, 4): columns (γ, e, t, parent) in
    creation (DFS) order — γ = clade index (−1 for loss nodes), e = species-tree node index, t = 1-based slice
    row where the lineage enters the state (0 for loss nodes), parent = row of the parent node (−1: root).
    Post-processing into timetrees / summaries (src/track.jl:428-488, src/rectree.jl) is host-side and out of
    the hot path."""
    xs, single = _as_vector(x)
    mh, dh = _data_handle(wm, xs)
    F = len(xs)
    if uniforms is None:
        uniforms = np.random.default_rng(seed).random((F, n_samples, 4 * max_nodes))
    cnt, st, nodes = _lib.get().backtrack(mh, dh, n_samples, uniforms, max_nodes)
    if np.any(st == 1):
        f, s = np.argwhere(st == 1)[0]
        raise BacktrackFailed(f"Backtracking failed for family {f}, sample {s}")
    if np.any(st > 1):
        raise _lib.WhaleCudaError(3, "backtrack: node buffer or uniform stream too small (raise max_nodes / stride)")
    out = [[nodes[f, s, :cnt[f, s]].copy() for s in range(n_samples)] for f in range(F)]
    return out[0] if single else out


def _mixture(components, weights, xs, grad):
    wm0 = components[0]
    for wm in components[1:]:
        if wm.nn != wm0.nn or wm.n_params != wm0.n_params or not (np.array_equal(wm.kind, wm0.kind) and
                                                                   np.array_equal(wm.n_slices, wm0.n_slices)):
            raise ValueError("mixture components must share the species tree, slicing and rate parameterisation")
    xs, _ = _as_vector(xs)
    mh, dh = _data_handle(wm0, xs)
    X = np.stack([wm.x() for wm in components])
    lw = np.log(np.asarray(weights, dtype=np.float64))
    return _lib.get().mixture_logpdf_grad(mh, dh, X, lw, wm0.p_leaf(), CONDITIONS[wm0.condition], want_grad=grad)


def logpdf_mixture(components, weights, xs) -> float:
    """`logpdf(mm::MixtureModel{…,<:WhaleModel}, xs)` (src/core.jl:66-76): Σ_i logsumexp_j (ℓ_ij + log p_j −
    condition(component j)), evaluated on the device (`whale_mixture_logpdf_grad`): the F × K matrix never leaves
    HBM.  Components share the species tree and differ in their rates."""
    return _mixture(components, weights, xs, False)[0]


def logpdf_mixture_and_gradient(components, weights, xs):
    """The same with its gradient: (ℓ, ∂ℓ/∂raw parameters of every component [K × P], ∂ℓ/∂log p_j [K]) — what
    ForwardDiff would propagate through the mixture's logsumexp."""
    return _mixture(components, weights, xs, True)


def logpdf_modelarray(models, xs) -> float:
    """`logpdf(m::ModelArray, xs)` (src/core.jl:78-79): family i under its own model i (unconditioned single-CCD
    likelihoods, summed)."""
    return float(sum(logpdf(m, x) for m, x in zip(models, xs)))


def condition(wm: WhaleModel) -> float:
    """`condition(wm)` (src/condition.jl:11-29) evaluated on the device's slice tables: the difference between
    the unconditioned and the conditioned batch likelihood of any one family."""
    if CONDITIONS[wm.condition] == 0:
        return 0.0
    probe = getattr(wm, "_probe", None)
    if probe is None:
        raise ValueError("condition(wm) needs a probe family: call set_probe(wm, ccd) once (any CCD of the model)")
    xs, _ = _as_vector(probe)
    mh, dh = _data_handle(wm, xs)
    L = _lib.get()
    a = L.logpdf_grad(mh, dh, wm.x(), wm.p_leaf(), 0)[0]
    b = L.logpdf_grad(mh, dh, wm.x(), wm.p_leaf(), CONDITIONS[wm.condition])[0]
    return a - b


def set_probe(wm: WhaleModel, ccd):
    """Remember one CCD of this model so `condition(wm)` can be read off the device."""
    wm._probe = ccd
    return wm


def _trees_from(cnt, st, nodes, n_samples, F):
    if np.any(st == 1):
        raise BacktrackFailed("Backtracking failed (no event selected; numerically inconsistent ℓ)")
    if np.any(st == 2):
        raise RuntimeError("backtrack: max_nodes too small for a sampled tree")
    if np.any(st == 3):
        raise RuntimeError("backtrack: uniform stream exhausted (increase the stride)")
    return [[nodes[f, s, :cnt[f, s]].copy() for s in range(n_samples)] for f in range(F)]


def track(wm: WhaleModel, xs, posterior, n: int, fun=None, seed=None, max_nodes: int = 512, return_loglik=False,
          shared_draws=False, device_rng=False, summary=False):
    """`track(TreeTracker(model, data, df, fun), N)` (src/track.jl:30-63): for every family and each of its `n` samples
    draw a posterior row `i = rand(1:length(df))` (:50-53 — independently per family AND sample, as the reference does),
    re-parameterise the model (`fun(model, row)`, default `model(**row)`), `logpdf!` and backtrack one tree — on the
    device (`whale_track_sample`: for every row in use, logpdf! of exactly the families that drew it and their walks,
    enqueued back to back).  Returns trees[family][sample] as (γ, e, t, parent) arrays.

    shared_draws=True is round 1's variant (one draw per sample index shared by all families, `whale_track`; the only
    form that can also return the batch log-likelihood of every draw, `return_loglik`).  device_rng=True draws the walks'
    uniforms on the device (seeded counter-based generator) instead of shipping a host stream.  summary=True also
    returns the device-side `sumtrees` table (see `sumtrees_device`)."""
    rng = np.random.default_rng(seed)
    xs, _ = _as_vector(xs)
    fun = fun or (lambda m, row: m(**row))
    mh, dh = _data_handle(wm, xs)
    L, F = _lib.get(), len(xs)
    if shared_draws or return_loglik:
        rows = [posterior[int(rng.integers(len(posterior)))] for _ in range(n)]
        X = np.stack([fun(wm, row).x() for row in rows])
        U = rng.random((F, n, 4 * max_nodes))
        cnt, st, nodes, ll = L.track(mh, dh, X, wm.p_leaf(), CONDITIONS[wm.condition], U, max_nodes)
        trees = _trees_from(cnt, st, nodes, n, F)
        return (trees, ll) if return_loglik else trees
    X = np.stack([fun(wm, row).x() for row in posterior])
    ti = rng.integers(len(posterior), size=(F, n)).astype(np.int32)
    U = None if device_rng else rng.random((F, n, 4 * max_nodes))
    tot = L.track_sample(mh, dh, X, wm.p_leaf(), n, ti, U, seed=int(rng.integers(1 << 62)), max_nodes=max_nodes)
    trees = _trees_compact(L, dh, F, n, tot)
    return (trees, sumtrees_device(wm, xs, n)) if summary else trees


def _trees_compact(L, dh, F, n, total):
    cnt, st = L.trees_counts(dh, F * n)
    if np.any(st == 1):
        raise BacktrackFailed("Backtracking failed (no event selected; numerically inconsistent ℓ)")
    if np.any(st == 2):
        raise RuntimeError("backtrack: max_nodes too small for a sampled tree")
    if np.any(st == 3):
        raise RuntimeError("backtrack: uniform stream exhausted (increase the stride)")
    off, nodes = L.trees_get(dh, F * n, total)
    return [[nodes[off[f * n + s]:off[f * n + s + 1]] for s in range(n)] for f in range(F)]


def sumtrees_device(wm: WhaleModel, xs, n: int):
    """`sumtrees` (src/rectree.jl:113-133) for the trees left on the device by the last `track` / `backtrack_device`:
    per family the distinct reconciled trees (identity as in src/track.jl:95-113, hashed on the device), most frequent
    first (ties: first sampled first): list over families of lists of dicts {count, freq, first (sample index), hash}."""
    xs, _ = _as_vector(xs)
    mh, dh = _data_handle(wm, xs)
    F = len(xs)
    nd, h, c, f1, _ = _lib.get().trees_summary(dh, F, n)
    out = []
    for f in range(F):
        k = int(nd[f])
        order = sorted(range(k), key=lambda i: (-int(c[f, i]), int(f1[f, i])))
        out.append([{"count": int(c[f, i]), "freq": int(c[f, i]) / n, "first": int(f1[f, i]), "hash": int(h[f, i])} for i in order])
    return out


def treekey(nodes) -> frozenset:
    """Identity of a backtracked reconciled tree as the reference defines it (`nodehash`/`cladehash`,
    src/track.jl:95-113): the SET over nodes of (γ, e, {(γ, e) of the children}) — a loss node is (loss, e, γ of its
    sister) — so two samples that differ only in the slice `t` at which events happen are the same tree.  Here the
    key is the set itself (exact), not a 64-bit hash of it.  `nodes`: one (n_nodes, 4) array from `backtrack`."""
    nodes = np.asarray(nodes)
    n = len(nodes)
    ch = [[] for _ in range(n)]
    for i in range(n):
        p = int(nodes[i, 3])
        if p >= 0:
            ch[p].append(i)
    keys = set()
    for i in range(n):
        g, e, p = int(nodes[i, 0]), int(nodes[i, 1]), int(nodes[i, 3])
        if g < 0:  # loss node
            sib = [j for j in ch[p] if j != i] if p >= 0 else []
            keys.add((-1, e, int(nodes[sib[0], 0]) if sib else -2))
        else:
            keys.add((g, e, frozenset((int(nodes[j, 0]), int(nodes[j, 1])) for j in ch[i])))
    return frozenset(keys)


def sumtrees(trees):
    """`sumtrees(trees, ccd, wm)` (src/rectree.jl:113-133) without the labelled event tables: the distinct
    reconciled trees of ONE family among its N backtracked samples with their posterior frequencies, most frequent
    first (ties: first sampled first), and the `cladecounts` table (src/rectree.jl:138: how many samples contain
    each reconciled clade).  Host-side post-processing of `backtrack`/`track` output; for all families pass
    `[sumtrees(t) for t in trees]` like the reference's matrix method (:110-111).

    Returns (summary, clades): summary = list of dicts {freq, count, tree (the first sample with that identity),
    key}; clades = dict reconciled-clade key -> number of samples containing it."""
    N = len(trees)
    if N == 0:
        return [], {}
    keys = [treekey(t) for t in trees]
    first, counts, clades = {}, {}, {}
    for i, k in enumerate(keys):
        first.setdefault(k, i)
        counts[k] = counts.get(k, 0) + 1
        for c in k:
            clades[c] = clades.get(c, 0) + 1
    order = sorted(counts, key=lambda k: (-counts[k], first[k]))
    summary = [{"freq": counts[k] / N, "count": counts[k], "tree": trees[first[k]], "key": k} for k in order]
    return summary, clades
