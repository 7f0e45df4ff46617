"""Host logic of the package: model structure, .ale ingest and bit-exact flattening, the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import whale_jl_b200 as W
from whale_jl_b200 import lib as wlib, synth
from conftest import HAVE_REF, REF, ROOT, load_golden

needs_ref = pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted (GPU box)")


def c1_model(**kw):
    t = synth.c1_species_tree()
    r = W.DLWGD(lam=[1.0] * 17, mu=[1.0] * 17, q=[0.2, 0.1], eta=0.9)
    return W.WhaleModel(r, t, 0.05, **kw)


def test_model_desc_matches_golden():
    """The package's WhaleModel flattens to exactly the arrays the oracle derived from src/model.jl."""
    g = load_golden("c1_example1")
    w = c1_model()
    for a, b in [(w.order, "m_order"), (w.child0, "m_child0"), (w.child1, "m_child1"), (w.kind, "m_kind"),
                 (w.n_slices, "m_nslices"), (w.lam_slot, "m_lam_slot"), (w.mu_slot, "m_mu_slot"),
                 (w.q_slot, "m_q_slot")]:
        assert np.array_equal(a, g[b]), b
    assert np.array_equal(w.slice_dt, g["m_dt"]) and np.array_equal(w.leafP, g["m_leafP"])
    assert (w.eta_slot, w.n_params, w.log_scale) == (int(g["m_eta_slot"]), int(g["m_P"]), int(g["m_log_scale"]))
    assert np.array_equal(w.x(), g["xs"][0])


def test_model_desc_constant_and_mul():
    g = load_golden("const_wgdturing")
    t = W.extree()
    W.insertnode(W.getlca(t, "PPAT", "PPAT"), name="wgd_1")
    W.insertnode(W.getlca(t, "ATHA", "ATRI"), name="wgd_2")
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.1, mu=0.2, q=[0.2, 0.1], eta=0.9), t, 0.1, minn=10, maxn=20)
    assert np.array_equal(w.order, g["m_order"]) and np.array_equal(w.n_slices, g["m_nslices"])
    assert np.array_equal(w.q_slot, g["m_q_slot"]) and w.n_params == 5 and w.log_scale == 0
    g = load_golden("mul_tree")
    mul = W.readnw("((MPOL:4.752,PPAT:4.752):0.292,((SMOE:4.0,PPAT:4.0):0.457,(((OSAT:1.555,(ATHA:0.55"
                   "48,CPAP:0.5548):1.0002):0.738,ATRI:2.293):1.225,(GBIL:3.178,PABI:3.178):0.34):0.93"
                   "9):0.587);")
    n = len(W.postwalk(mul))
    w = W.WhaleModel(W.DLWGD(lam=[-1.0] * n, mu=[-1.0] * n, eta=0.9), mul, 0.05)
    assert np.array_equal(w.leafP, g["m_leafP"]) and np.array_equal(w.order, g["m_order"])
    assert sorted(w.leafP[w.kind == 0].tolist())[:2] == [0.5, 0.5]  # the duplicated PPAT leaves


@needs_ref
@pytest.mark.parametrize("fixture,path,sel", [("c1_example1", "example/example-1/ale", None)])
def test_read_ale_flattening_is_bit_exact(fixture, path, sel):
    """Integer packing parity: clade order, triple order, compat lists and probabilities are identical to the
    oracle's restatement of src/ccd.jl."""
    g = load_golden(fixture)
    w = c1_model()
    flat = W.read_ale(f"{REF}/{path}", w).flatten(w.nn)
    for k, gk in [("clade_off", "f_clade_off"), ("clade_nleaf", "f_nleaf"), ("split_off", "f_split_off"),
                  ("g1", "f_g1"), ("g2", "f_g2"), ("compat_off", "f_compat_off"), ("compat", "f_compat")]:
        assert np.array_equal(flat[k], g[gk]), k
    assert flat["p"].tobytes() == g["f_p"].tobytes()


@needs_ref
def test_read_ale_landplant_and_mul_flattening():
    g = load_golden("landplant100")
    tl = W.readnw(open(f"{REF}/docs/data/landplant/speciestree.nw").readline())
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.1, mu=0.2, eta=1 / 1.5), tl, 0.05)
    flat = W.read_ale(f"{REF}/docs/data/landplant/100fams", w).flatten(w.nn)
    assert np.array_equal(flat["g1"], g["f_g1"]) and np.array_equal(flat["compat"], g["f_compat"])
    assert flat["p"].tobytes() == g["f_p"].tobytes()


def test_synthetic_ale_roundtrip(tmp_path):
    """Generator -> .ale text -> read_ale: counts survive, probabilities per clade sum to 1, the oracle's
    independent parser produces the identical flattening."""
    from oracle import whale_oracle as wo, flat as oflat
    d = synth.generate(str(tmp_path / "fams"), 4, seed=7)
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), synth.c1_species_tree(), 0.05)
    ccds = W.read_ale(d, w)
    assert len(ccds) == 4
    for c in ccds:
        for k in range(len(c)):
            s = c.p[c.split_off[k]:c.split_off[k + 1]].sum()
            assert c.nleaf[k] == 1 or abs(s - 1.0) < 1e-12
        assert np.all(c.g1 < np.repeat(np.arange(len(c)), np.diff(c.split_off)))  # sub-clades have smaller ids
    ow = wo.WhaleModel(wo.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), wo.c1_tree(), 0.05)
    ff = oflat.FlatFams(wo.read_ale(d, ow), len(ow))
    fl = ccds.flatten(w.nn)
    assert np.array_equal(fl["g1"], ff.g1) and np.array_equal(fl["compat"], ff.compat)
    assert fl["p"].tobytes() == ff.p.tobytes()


def test_reparameterise_keeps_structure():
    w = c1_model()
    w2 = w(eta=0.5, q=[0.3, 0.4])
    assert w2.rates.eta == 0.5 and w.rates.eta == 0.9 and w2.n_params == w.n_params
    assert np.array_equal(w2.x()[-3:], [0.3, 0.4, 0.5])


def test_library_exports_every_header_symbol():
    """The built C-ABI library loads (no GPU needed) and exports every function include/whalecuda.h declares."""
    hdr = open(os.path.join(ROOT, "include", "whalecuda.h")).read()
    declared = set(re.findall(r"\b(whale_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(wlib.SYMBOLS), declared ^ set(wlib.SYMBOLS)
    assert os.path.exists(wlib.LIB_PATH), "run __graft_entry__.build() first"
    L = ctypes.CDLL(wlib.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    assert L.whale_version() >= 100


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path fails loudly (there is no CPU fallback)."""
    L = wlib.Lib()
    if L.L.whale_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(wlib.WhaleCudaError) as ei:
        L.model_create(c1_model())
    assert "no CUDA device" in str(ei.value)


def test_model_validation_errors():
    L = wlib.Lib(os.path.join(ROOT, "tests", "emu", "libwhalecuda_emu.so")) if os.path.exists(
        os.path.join(ROOT, "tests", "emu", "libwhalecuda_emu.so")) else None
    if L is None:
        pytest.skip("emulation build missing")
    w = c1_model()
    bad = w()
    bad.order = w.order[::-1].copy()
    with pytest.raises(wlib.WhaleCudaError):
        L.model_create(bad)


def test_sumtrees_identity_ignores_slice_times():
    """`nodehash` (src/track.jl:105-113): reconciled trees that differ only in the slice at which events happen are
    the same tree; a different branch for one event is a different tree; frequencies follow `sumtrees`
    (src/rectree.jl:113-133)."""
    import whale_jl_b200 as W
    # (γ, e, t, parent): root duplication-free toy tree with one loss node (γ = −1)
    a = np.array([[4, 6, 1, -1], [2, 5, 3, 0], [3, 4, 7, 0], [0, 1, 1, 1], [-1, 2, 0, 1], [1, 3, 1, 2]])
    b = a.copy(); b[1, 2] = 5; b[2, 2] = 2            # other slice times only
    c = a.copy(); c[2, 1] = 5                          # clade 3 reconciled on another branch
    assert W.treekey(a) == W.treekey(b) != W.treekey(c)
    summary, clades = W.sumtrees([a, c, b, a])
    assert [s["count"] for s in summary] == [3, 1] and summary[0]["freq"] == 0.75
    assert summary[0]["tree"] is a and summary[1]["tree"] is c
    assert clades[(4, 6, frozenset({(2, 5), (3, 4)}))] == 3 and clades[(-1, 2, 0)] == 4
    assert W.sumtrees([]) == ([], {})


def test_empty_batch_is_zero():
    """`logpdf(wm, CCD[])` is the empty sum minus 0·condition (src/core.jl:54,63): 0 and a zero gradient, without touching
    the device (the C ABI itself rejects n_fam = 0 with WHALE_ERR_ARG)."""
    import whale_jl_b200 as W
    from whale_jl_b200 import synth
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), synth.c1_species_tree(), 0.05)
    assert W.logpdf(w, []) == 0.0
    ll, g = W.logpdf_and_gradient(w, [])
    assert ll == 0.0 and g.shape == (w.n_params,) and not g.any()
    lf, gf = W.logpdf_per_family(w, [], grad=True)
    assert lf.shape == (0,) and gf.shape == (0, w.n_params)
