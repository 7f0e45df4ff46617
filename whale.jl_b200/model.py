"""WhaleModel — host-side mirror of the reference's model object (src/model.jl).

Only the *structure* is built here (once per species tree): node order and ids (src/model.jl:96-145),
wgd ids (preorder, :102-109), slice counts and lengths (`Slices`, :16-22), species clades incl. the
MUL-tree id sharing (:32-42,133-135), and the mapping from nodes to entries of the raw parameter vector
(`getθ`, src/rmodels.jl:31-33,55-64).  The per-θ arithmetic of `setmodel!` (src/model.jl:162-191) is NOT
done on the host: it runs on the GPU (k_tables) inside every `logpdf` call.

Everything is stored as flat 0-based numpy arrays (node index = Julia id − 1), which is exactly what
crosses the C ABI (`whale_model_desc`, include/whalecuda.h).
"""
from __future__ import annotations

import math

import numpy as np

from . import newick
from .rates import ConstantDLWGD, DLWGD

LEAF, INTERNAL, WGD, ROOT = 0, 1, 2, 3
CONDITIONS = {"none": 0, "root": 1, "nonextinct": 2, "nowhere": 3,
              "NoCondition": 0, "RootCondition": 1, "NonExtinctCondition": 2, "NowhereExtinctCondition": 3}


def iswgd(name: str) -> bool:
    """src/model.jl:54"""
    return name.startswith("wgd")


class WhaleModel:
    """`WhaleModel(rates, tree, Δt; minn=5, maxn=10000, condition=RootCondition())` (src/model.jl:96-97)."""

    def __init__(self, rates, tree: newick.Node, dt: float, minn: int = 5, maxn: int = 10000,
                 condition: str = "root"):
        if condition not in CONDITIONS:
            raise ValueError(f"unsupported condition {condition!r} (have {sorted(CONDITIONS)})")
        self.rates = rates
        self.condition = condition
        self.dt, self.minn, self.maxn = dt, minn, maxn
        self._handle = None  # device model handle, created lazily (lib.py); shared by the wm(θ) copies of this model

        post = newick.postwalk(tree)
        leaves = newick.getleaves(tree)
        names = [l.name for l in leaves]
        mul = {nm: names.count(nm) for nm in set(names) if names.count(nm) > 1}
        # wgdid: preorder counter over WGD nodes
        wgdid = {}
        k = 0
        todo = [tree]
        while todo:
            x = todo.pop()
            if iswgd(x.name):
                k += 1
                wgdid[id(x)] = k
            todo.extend(reversed(x.children))
        order = leaves + [x for x in post if not x.isleaf]
        nn = len(order)
        nonwgd = sum(1 for x in order if not iswgd(x.name))
        ids = {}
        nxt_plain, nxt_wgd = 1, nonwgd + 1
        mulid = {}
        for x in order:
            if iswgd(x.name):
                ids[id(x)] = nxt_wgd
                nxt_wgd += 1
            else:
                ids[id(x)] = nxt_plain
                if x.name in mul:
                    mulid[x.name] = nxt_plain  # last copy wins
                nxt_plain += 1
        idx = lambda x: ids[id(x)] - 1
        self.nn = nn
        self.nwgd = k
        self.order = np.array([idx(x) for x in order], np.int32)
        self.names = [""] * nn
        self.subtree = [""] * nn
        self.child0 = np.full(nn, -1, np.int32)
        self.child1 = np.full(nn, -1, np.int32)
        self.parent = np.full(nn, -1, np.int32)
        self.kind = np.zeros(nn, np.int32)
        self.wgdid = np.zeros(nn, np.int32)
        self.distance = np.full(nn, math.nan)
        self.n_slices = np.zeros(nn, np.int32)
        self.slice_dt = np.zeros(nn)
        self.leafP = np.zeros(nn)
        self.clade: list[frozenset] = [frozenset()] * nn
        for x in order:
            e = idx(x)
            self.names[e] = x.name
            self.subtree[e] = newick.nwstr(x)
            if len(x.children) > 2 or (len(x.children) == 1 and not iswgd(x.name)):
                raise ValueError("species tree must be binary (single-child nodes only for WGDs)")
            if iswgd(x.name) and len(x.children) != 1:
                raise ValueError("a WGD node must have exactly one child")
            if x.children:
                self.child0[e] = idx(x.children[0])
                if len(x.children) == 2:
                    self.child1[e] = idx(x.children[1])
            if x.parent is not None:
                self.parent[e] = idx(x.parent)
            self.kind[e] = WGD if iswgd(x.name) else ROOT if x.parent is None else LEAF if x.isleaf else INTERNAL
            self.wgdid[e] = wgdid.get(id(x), 0)
            t = x.distance
            self.distance[e] = t
            n = 0 if math.isnan(t) else min(maxn, max(minn, math.ceil(t / dt)))
            self.n_slices[e] = n
            self.slice_dt[e] = 0.0 if n == 0 else t / n
            self.leafP[e] = 1.0 / mul[x.name] if x.name in mul else (1.0 if x.isleaf else 0.0)
        for e in self.order:  # children first
            if self.kind[e] == LEAF:
                nm = self.names[e]
                self.clade[e] = frozenset([mulid[nm] if nm in mulid else e + 1])
            else:
                c = self.clade[self.child0[e]]
                if self.child1[e] >= 0:
                    c = c | self.clade[self.child1[e]]
                self.clade[e] = c
        self.root = int(self.order[-1])
        # species name -> id used for gene->species mapping (read_ale, src/ccd.jl:128)
        self.spmap = {}
        for e in self.order:
            if self.kind[e] == LEAF:
                self.spmap[self.names[e]] = int(e) + 1
        self.row_off = np.concatenate([[0], np.cumsum(self.n_slices + 1)]).astype(np.int64)
        self._set_slots()

    # ---- parameter layout ----
    def _set_slots(self):
        r = self.rates
        nn = self.nn
        nr, nq = r.nrates, len(r.q)
        if nq < self.nwgd:
            raise ValueError(f"model has {self.nwgd} WGD nodes but rates.q has {nq} entries")
        self.lam_slot = np.full(nn, -1, np.int32)
        self.mu_slot = np.full(nn, -1, np.int32)
        self.q_slot = np.full(nn, -1, np.int32)
        for e in range(nn):
            c = e
            while self.kind[c] == WGD:  # nonwgdchild src/model.jl:200-203
                c = self.child0[c]
            if isinstance(r, ConstantDLWGD):
                self.lam_slot[e], self.mu_slot[e] = 0, 1
            elif c + 1 <= nr:
                self.lam_slot[e], self.mu_slot[e] = c, nr + c
            if self.kind[e] == WGD:
                self.q_slot[e] = 2 * nr + self.wgdid[e] - 1
        self.eta_slot = 2 * nr + nq
        self.n_params = 2 * nr + nq + 1
        self.log_scale = 1 if r.log_scale else 0

    def layout_key(self):
        return (type(self.rates).__name__, self.rates.nrates, len(self.rates.q))

    # ---- the Julia API ----
    def __len__(self):
        return self.nn

    def __call__(self, rates=None, **theta) -> "WhaleModel":
        """`model(θ)` / `model(rates)` (src/model.jl:147-160): same structure, new parameters.  The device
        model handle is shared as long as the parameter layout is unchanged."""
        import copy
        new = copy.copy(self)
        new.rates = rates if rates is not None else self.rates(**theta)
        if new.layout_key() != self.layout_key():
            new._handle = None
            new._set_slots()
        return new

    def x(self) -> np.ndarray:
        """The raw parameter vector [λ…, μ…, q…, η] on the rates struct's own scale."""
        v = self.rates.vector()
        if len(v) != self.n_params:
            raise ValueError("rates do not match the model's parameter layout")
        return v

    def p_leaf(self) -> np.ndarray:
        """getp (src/rmodels.jl:14): sampling-failure probability per node (leaves only)."""
        out = np.zeros(self.nn)
        p = self.rates.p
        if len(p) > 0:
            for e in range(self.nn):
                if self.kind[e] == LEAF:
                    out[e] = p[e]
        return out

    def setsamplingp(self, d: dict):
        """`setsamplingp!(model, dict)` src/model.jl:248-255."""
        nl = int(np.sum(self.kind == LEAF))
        if len(self.rates.p) == 0:
            self.rates.p = [0.0] * nl
        for e in range(self.nn):
            if self.kind[e] == LEAF:
                self.rates.p[e] = d.get(self.names[e], 0.0)

    def show(self) -> str:
        """Structure table of `show(io, m)` (src/model.jl:214-228)."""
        lines = [f"{self.nn} nodes ({int(np.sum(self.kind == LEAF))} leaves, {self.nwgd} WGD nodes)",
                 "node_id,wgd_id,distance,Δt,n,subtree"]
        for e in self.order:
            d = self.distance[e]
            lines.append(f"{e + 1},{self.wgdid[e]},{d if math.isnan(d) else round(d, 4)},"
                         f"{round(self.slice_dt[e], 4)},{self.n_slices[e]},\"{self.subtree[e]};\"")
        return "\n".join(lines)

    def __repr__(self):
        return f"WhaleModel({type(self.rates).__name__}, {self.nn} nodes, {int(self.n_slices.sum())} slices)"
