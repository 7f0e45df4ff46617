"""Rate-parameter containers of the reference (src/rmodels.jl).

`ConstantDLWGD` (src/rmodels.jl:23-33): one λ, μ for the whole tree, natural scale.
`DLWGD` (src/rmodels.jl:47-64): branch-wise λ, μ stored on the LOG scale and indexed by node id;
WGD nodes borrow the rates of their first non-WGD descendant; ids beyond the vector give NaN rates.
`q` is indexed by wgdid, `p` (sampling-failure probabilities) by leaf id, `eta` is the geometric prior
at the root.  The *raw vector* x = [λ…, μ…, q…, η] in the struct's own scale is what the gradient
returned by the device is taken with respect to (the AD contract).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np


@dataclass
class ConstantDLWGD:
    lam: float
    mu: float
    q: list = field(default_factory=list)
    p: list = field(default_factory=list)
    eta: float = 0.66

    log_scale = False

    @property
    def nrates(self) -> int:
        return 1

    def vector(self) -> np.ndarray:
        return np.array([self.lam, self.mu, *self.q, self.eta], dtype=np.float64)

    def from_vector(self, x) -> "ConstantDLWGD":
        nq = len(self.q)
        return ConstantDLWGD(lam=float(x[0]), mu=float(x[1]), q=[float(v) for v in x[2:2 + nq]], p=self.p,
                             eta=float(x[2 + nq]))

    def __call__(self, **theta) -> "ConstantDLWGD":
        """`m(θ)` (src/rmodels.jl:8-11): a copy with the given fields replaced."""
        return replace(self, **theta)


@dataclass
class DLWGD:
    lam: list
    mu: list
    q: list = field(default_factory=list)
    p: list = field(default_factory=list)
    eta: float = 0.66

    log_scale = True

    @property
    def nrates(self) -> int:
        return len(self.lam)

    def vector(self) -> np.ndarray:
        return np.array([*self.lam, *self.mu, *self.q, self.eta], dtype=np.float64)

    def from_vector(self, x) -> "DLWGD":
        n, nq = len(self.lam), len(self.q)
        return DLWGD(lam=[float(v) for v in x[:n]], mu=[float(v) for v in x[n:2 * n]],
                     q=[float(v) for v in x[2 * n:2 * n + nq]], p=self.p, eta=float(x[2 * n + nq]))

    def __call__(self, **theta) -> "DLWGD":
        return replace(self, **theta)
