"""The reverse-mode decomposition planned for the round-2 kernel (tools/adjoint_prototype.py; DESIGN.md §6 "next")
against the oracle's forward-mode gradient: DP adjoints of the table entries contracted with the table tangents."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("kind", ["constant", "branchwise", "critical"])
def test_adjoint_gradient_matches_forward_mode(tmp_path, kind):
    import whale_jl_b200  # noqa: F401
    from whale_jl_b200 import synth
    from oracle import whale_oracle as wo
    import adjoint_prototype as ap
    d = synth.generate(str(tmp_path / "adj"), 2, seed=8)
    if kind == "constant":
        r = wo.ConstantDLWGD(lam=0.25, mu=0.35, q=[0.3, 0.15], eta=0.7)
    elif kind == "critical":  # λ = μ: the critical branch of getα (src/bdputil.jl:6-7)
        r = wo.ConstantDLWGD(lam=0.3, mu=0.3, q=[0.2, 0.1], eta=0.67)
    else:  # DLWGD: log-scale branch rates, 37 raw parameters like the reference's test model
        rng = np.random.default_rng(1)
        r = wo.DLWGD(lam=list(rng.normal(np.log(0.2), 0.3, 17)), mu=list(rng.normal(np.log(0.25), 0.3, 17)),
                     q=[0.2, 0.1], eta=0.8)
    for cond in ("root", "none"):
        w = wo.WhaleModel(r, wo.c1_tree(), 0.05, condition=cond)
        xs = wo.read_ale(d, w)
        a, ga = ap.logpdf_and_gradient_adjoint(w, xs)
        b, gb = wo.logpdf_and_gradient(w, xs)
        assert a == pytest.approx(b, rel=1e-13)
        np.testing.assert_allclose(ga, gb, rtol=1e-10, atol=1e-12 * np.abs(gb).max())
