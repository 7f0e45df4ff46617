#!/usr/bin/env python
"""Reverse-mode (adjoint) gradient of the ALE/DLWGD log-likelihood — a CPU PROTOTYPE for the round-2 kernel.

DEVELOPMENT AID, not product code and not on any product path: it pins the mathematics of DESIGN.md §6 "next (1)"
against the oracle's forward-mode (dual number) gradient before any CUDA is written (tests/test_adjoint_prototype.py).

Decomposition (what the device kernel will do per family):
  forward   value-only DP keeping every row of every branch                       (oracle.logpdf_single(keep=True))
  backward  ℓ̄ rows from the root down (src/core.jl:83-199 transposed), accumulating the adjoints of everything the
            DP reads from the slice tables: ϕ̄_{e,i}, ψ̄_{e,i}, ϵ̄_e (the last ϵ of a child as used by Πloss / Πwgdloss /
            the root coefficients), q̄_e, η̄
  contract  ∂ log L/∂θ_k = Σ_{e,i} ϕ̄·∂ϕ/∂θ_k + ψ̄·∂ψ/∂θ_k + Σ_e ϵ̄_e·∂ϵ_e/∂θ_k + q̄·∂q/∂θ_k + η̄·∂η/∂θ_k with the table
            TANGENTS that k_tables already produces (here: the oracle's setmodel on Dual rates)
so the DP's cost no longer depends on the number of parameters P.  Leaf branches need nothing special here (their
rows are reached by the same backward recursion); on the device they keep their K = 3 forward tangents.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import whale_oracle as wo  # noqa: E402


def _idx(x, g, e):
    """0-based column of clade g in ℓ[e], or −1 (getl == 0, src/ccd.jl:41-49)."""
    return x.index[g][e] - 1


def family_adjoint(wm, x):
    """log L of one family and the adjoints of the table entries the DP reads.
    Returns (logL, phib, psib, epsb, qb, etab): phib/psib[node id] = per-row lists, epsb[node id] = adjoint of that
    node's LAST ϵ as read by its parent's formulas (for the root: of its own ϵ), qb[node id], etab."""
    logL = wo.logpdf_single(wm, x, keep=True)
    ell = x.ell
    nn = len(wm)
    ellb = [None] * (nn + 1)
    for n in wm.order:
        ellb[n.id] = [[0.0] * len(x.compat[n.id]) for _ in range(len(n))]
    phib = {n.id: [0.0] * len(n) for n in wm.order}
    psib = {n.id: [0.0] * len(n) for n in wm.order}
    epsb = {n.id: 0.0 for n in wm.order}
    qb = {n.id: 0.0 for n in wm.order}
    etab = 0.0
    root = wm.root
    L = ell[root.id][0][-1]
    if not (L > 0.0):
        return -math.inf, phib, psib, epsb, qb, etab
    ellb[root.id][0][-1] = 1.0 / L  # d log L / dL

    def spec_loss_backward(n, c, coef):
        """Transposed Πspeciation + Πloss (src/core.jl:160-176) of clade c at node n: `coef` = adjoint of their sum."""
        f, g = n.children
        lf, lg = ell[f.id][-1], ell[g.id][-1]
        bf, bg = ellb[f.id][-1], ellb[g.id][-1]
        if not c.isleaf():
            for (g1, g2, pr) in c.splits:
                f1, f2, h1, h2 = _idx(x, g1, f.id), _idx(x, g2, f.id), _idx(x, g1, g.id), _idx(x, g2, g.id)
                if f1 >= 0 and h2 >= 0:  # ℓ_f[γ1]·ℓ_g[γ2]
                    bf[f1] += coef * pr * lg[h2]
                    bg[h2] += coef * pr * lf[f1]
                if h1 >= 0 and f2 >= 0:  # ℓ_g[γ1]·ℓ_f[γ2]
                    bg[h1] += coef * pr * lf[f2]
                    bf[f2] += coef * pr * lg[h1]
        cf, cg = _idx(x, c.id, f.id), _idx(x, c.id, g.id)
        if cf >= 0:  # ℓ_f[γ]·ϵ_g
            bf[cf] += coef * g.eps[-1]
            epsb[g.id] += coef * lf[cf]
        if cg >= 0:  # ℓ_g[γ]·ϵ_f
            bg[cg] += coef * f.eps[-1]
            epsb[f.id] += coef * lg[cg]

    for n in reversed(wm.order):
        e = n.id
        if n.isroot():  # whaleroot! src/core.jl:130-158, clades in DESCENDING size
            eta = wo.gettheta(wm.rates, n)["eta"]
            eps = n.eps[-1]
            xi = 1.0 - (1.0 - eta) * eps
            A = (1.0 - eta) * xi / eta
            B = eta * (1.0 - eps) / xi ** 2
            Ab = Bb = 0.0
            row, rowb = ell[e][0], ellb[e][0]
            for c in reversed(x.clades[1:]):
                cb = rowb[c.id - 1]
                if cb == 0.0:
                    continue
                a = b = 0.0
                if not c.isleaf():
                    for (g1, g2, pr) in c.splits:  # Πroot :151-158 (same row)
                        a += pr * row[g1 - 1] * row[g2 - 1]
                        rowb[g1 - 1] += cb * A * pr * row[g2 - 1]
                        rowb[g2 - 1] += cb * A * pr * row[g1 - 1]
                    b = wo._Pspeciation(x, c, ell, n)
                cc = wo._Ploss(x, c, ell, n)
                Ab += cb * a
                Bb += cb * (b + cc)
                spec_loss_backward(n, c, cb * B)
            # A = (1−η)ξ/η, B = η(1−ϵ)/ξ², ξ = 1−(1−η)ϵ
            dxi_deta, dxi_deps = eps, -(1.0 - eta)
            dA_deta = -xi / eta + (1.0 - eta) * dxi_deta / eta - (1.0 - eta) * xi / eta ** 2
            dA_deps = (1.0 - eta) * dxi_deps / eta
            dB_deta = (1.0 - eps) / xi ** 2 - 2.0 * eta * (1.0 - eps) * dxi_deta / xi ** 3
            dB_deps = -eta / xi ** 2 - 2.0 * eta * (1.0 - eps) * dxi_deps / xi ** 3
            etab += Ab * dA_deta + Bb * dB_deta
            epsb[e] += Ab * dA_deps + Bb * dB_deps
            continue
        # ---- slices n+1 .. 2 (within_branch! src/core.jl:121-128,178-185), row by row from the top ----
        for i in range(len(n), 1, -1):
            cur, prev = ellb[e][i - 1], ell[e][i - 2]
            prevb = ellb[e][i - 2]
            phi, psi = n.phi[i - 1], n.psi[i - 1]
            for c_id in x.compat[e]:
                j = x.index[c_id][e] - 1
                lb = cur[j]
                if lb == 0.0:
                    continue
                phib[e][i - 1] += lb * prev[j]
                prevb[j] += lb * phi
                c = x.clades[c_id]
                if not c.isleaf():
                    p = 0.0
                    for (g1, g2, pr) in c.splits:
                        i1, i2 = _idx(x, g1, e), _idx(x, g2, e)
                        if i1 < 0 or i2 < 0:
                            continue
                        p += pr * prev[i1] * prev[i2]
                        prevb[i1] += lb * psi * pr * prev[i2]
                        prevb[i2] += lb * psi * pr * prev[i1]
                    psib[e][i - 1] += lb * p
        # ---- row 1 ----
        if n.iswgd():  # whalewgd! src/core.jl:103-119,187-199
            q = wo.gettheta(wm.rates, n)["q"]
            f = n.children[0]
            lf, bf = ell[f.id][-1], ellb[f.id][-1]
            ef = f.eps[-1]
            for c_id in x.compat[e]:
                lb = ellb[e][0][x.index[c_id][e] - 1]
                if lb == 0.0:
                    continue
                c = x.clades[c_id]
                s = 0.0
                if not c.isleaf():
                    for (g1, g2, pr) in c.splits:
                        i1, i2 = _idx(x, g1, f.id), _idx(x, g2, f.id)
                        if i1 < 0 or i2 < 0:
                            continue
                        s += pr * lf[i1] * lf[i2]
                        bf[i1] += lb * q * pr * lf[i2]
                        bf[i2] += lb * q * pr * lf[i1]
                cf = _idx(x, c_id, f.id)
                lc = lf[cf] if cf >= 0 else 0.0
                qb[e] += lb * (s - lc + 2.0 * ef * lc)
                if cf >= 0:
                    bf[cf] += lb * ((1.0 - q) + 2.0 * q * ef)
                    epsb[f.id] += lb * 2.0 * q * lc
        elif not n.isleaf():  # whale! src/core.jl:83-101: Πspeciation + Πloss
            for c_id in x.compat[e]:
                lb = ellb[e][0][x.index[c_id][e] - 1]
                if lb != 0.0:
                    spec_loss_backward(n, x.clades[c_id], lb)
        # leaf branch: row 1 is leafℙ or 0 — constants
    return logL, phib, psib, epsb, qb, etab


def logpdf_and_gradient_adjoint(wm, xs, x0=None):
    """Σ_f log L_f − N·condition and its gradient w.r.t. the raw parameter vector, the DP differentiated in REVERSE
    mode; only the slice tables (and the condition) see dual numbers."""
    x0 = [float(v) for v in (x0 if x0 is not None else wo.vector_from_rates(wm.rates))]
    P = len(x0)
    wv = wm.with_rates(wo.rates_from_vector(wm.rates, x0))
    wo.setmodel(wv)
    duals = [wo.Dual(v, np.eye(P)[i]) for i, v in enumerate(x0)]
    wd = wm.with_rates(wo.rates_from_vector(wm.rates, duals))
    wo.setmodel(wd)  # table tangents (what k_tables computes)
    dual_of = {n.id: n for n in wd.order}

    def tang(v):
        return v.d if isinstance(v, wo.Dual) else np.zeros(P)

    tot, grad = 0.0, np.zeros(P)
    for x in xs:
        ll, phib, psib, epsb, qb, etab = family_adjoint(wv, x)
        tot += ll
        if not math.isfinite(ll):
            continue
        for n in wv.order:
            nd = dual_of[n.id]
            for i in range(1, len(n)):
                grad += phib[n.id][i] * tang(nd.phi[i]) + psib[n.id][i] * tang(nd.psi[i])
            grad += epsb[n.id] * tang(nd.eps[-1])
            if n.iswgd():
                grad += qb[n.id] * tang(wo.gettheta(wd.rates, nd)["q"])
        grad += etab * tang(wo.gettheta(wd.rates, wd.root)["eta"])
    c = wo.condition(wd)
    cv, cd = (c.v, c.d) if isinstance(c, wo.Dual) else (c, np.zeros(P))
    return tot - len(xs) * cv, grad - len(xs) * cd


if __name__ == "__main__":
    import tempfile
    import whale_jl_b200  # noqa: F401  (the synthetic generator lives with the package)
    from whale_jl_b200 import synth
    d = synth.generate(os.path.join(tempfile.mkdtemp(), "adj"), 3, seed=4)
    w = wo.WhaleModel(wo.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), wo.c1_tree(), 0.05)
    xs = wo.read_ale(d, w)
    a, ga = logpdf_and_gradient_adjoint(w, xs)
    b, gb = wo.logpdf_and_gradient(w, xs)
    print("adjoint ", a, ga)
    print("forward ", b, gb)
    print("max rel diff", np.max(np.abs(ga - gb) / np.abs(gb)))
