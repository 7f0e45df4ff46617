import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
HAVE_REF = os.path.isdir(os.path.join(REF, "example"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def golden_model(g):
    """A duck-typed model carrying the whale_model_desc arrays of a golden fixture."""
    return types.SimpleNamespace(
        nn=len(g["m_order"]), order=g["m_order"], child0=g["m_child0"], child1=g["m_child1"], kind=g["m_kind"],
        n_slices=g["m_nslices"], slice_dt=g["m_dt"], leafP=g["m_leafP"], n_params=int(g["m_P"]),
        lam_slot=g["m_lam_slot"], mu_slot=g["m_mu_slot"], q_slot=g["m_q_slot"], eta_slot=int(g["m_eta_slot"]),
        log_scale=int(g["m_log_scale"]))


def golden_fams(g, sel=None):
    """whale_ccd_desc arrays of a golden fixture (optionally a subset of families)."""
    nn = len(g["m_order"])
    F = len(g["f_clade_off"]) - 1
    sel = list(range(F)) if sel is None else list(sel)
    co, so, cpo = g["f_clade_off"], g["f_split_off"], g["f_compat_off"]
    clade_off, nleaf, split_off, g1, g2, p, compat_off, compat = [0], [], [0], [], [], [], [0], []
    for f in sel:
        c0, c1 = co[f], co[f + 1]
        nleaf.append(g["f_nleaf"][c0:c1])
        s0, s1 = so[c0], so[c1]
        split_off.extend((so[c0 + 1:c1 + 1] - s0 + len(np.concatenate(g1)) if g1 else so[c0 + 1:c1 + 1] - s0).tolist())
        g1.append(g["f_g1"][s0:s1]); g2.append(g["f_g2"][s0:s1]); p.append(g["f_p"][s0:s1])
        clade_off.append(clade_off[-1] + int(c1 - c0))
        base = len(np.concatenate(compat)) if compat else 0
        k0, k1 = cpo[f * nn], cpo[(f + 1) * nn]
        compat_off.extend((cpo[f * nn + 1:(f + 1) * nn + 1] - k0 + base).tolist())
        compat.append(g["f_compat"][k0:k1])
    return dict(n_fam=len(sel), clade_off=np.array(clade_off, np.int64), clade_nleaf=np.concatenate(nleaf).astype(np.int32),
                split_off=np.array(split_off, np.int64), g1=np.concatenate(g1).astype(np.int32),
                g2=np.concatenate(g2).astype(np.int32), p=np.concatenate(p).astype(np.float64),
                compat_off=np.array(compat_off, np.int64), compat=np.concatenate(compat).astype(np.int32))


COND = {"none": 0, "root": 1, "nonextinct": 2}


def run_parity(L, name, rtol=1e-9, sel=None, conds=None, check_family=True):
    """Drive the C ABI (library instance L) with a golden fixture and compare with the oracle's outputs.
    Tolerance: 1e-9 relative (the north-star bar for fp64 results)."""
    g = load_golden(name)
    mh = L.model_create(golden_model(g))
    F = len(g["f_clade_off"]) - 1
    fl = golden_fams(g, sel)
    dh = L.data_create(mh, fl)
    idx = list(range(F)) if sel is None else list(sel)
    try:
        for kind in conds or [k[4:] for k in g if k.startswith("tot_")]:
            for xi, x in enumerate(g["xs"]):
                ll, grad, lf, gf = L.logpdf_grad(mh, dh, x, g["m_pleaf"], COND[kind], want_grad=True,
                                                 per_family=True, per_family_grad=True)
                if sel is None:
                    want = g[f"tot_{kind}"][xi]
                    assert ll == pytest.approx(want, rel=rtol), (name, kind, xi)
                    wg = g[f"grad_{kind}"][xi]
                    np.testing.assert_allclose(grad, wg, rtol=rtol, atol=1e-9 * np.abs(wg).max())
                if check_family:
                    np.testing.assert_allclose(lf, g["ll_fam"][xi][idx], rtol=rtol)
                    wgf = g["grad_fam"][xi][idx]
                    np.testing.assert_allclose(gf, wgf, rtol=rtol, atol=1e-9 * np.abs(wgf).max())
                ll2, _, _, _ = L.logpdf_grad(mh, dh, x, g["m_pleaf"], COND[kind], want_grad=False)
                assert ll2 == pytest.approx(ll, rel=1e-12)
    finally:
        L.L.whale_data_destroy(dh)
        L.L.whale_model_destroy(mh)
    return g
