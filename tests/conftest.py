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


def repeated_calls_parity(L, name, rounds=3):
    """The same (flags, condition) evaluated many times on ONE handle with changing θ — the pattern of an optimiser; from the
    third call on a host-pointer evaluation may be a CUDA-graph replay (WHALE_GRAPHS) — gradient, value-only and ℓ-keeping
    calls interleaved, each checked against the golden values."""
    g = load_golden(name)
    mh = L.model_create(golden_model(g))
    dh = L.data_create(mh, golden_fams(g))
    kinds = [k[4:] for k in g if k.startswith("tot_")]
    try:
        for r in range(rounds):
            for kind in kinds:
                for xi, x in enumerate(g["xs"]):
                    want, wg = g[f"tot_{kind}"][xi], g[f"grad_{kind}"][xi]
                    ll, grad, _, _ = L.logpdf_grad(mh, dh, x, g["m_pleaf"], COND[kind], want_grad=True)
                    assert ll == pytest.approx(want, rel=1e-9), (name, r, kind, xi)
                    np.testing.assert_allclose(grad, wg, rtol=1e-9, atol=1e-9 * np.abs(wg).max())
                    ll0, _, _, _ = L.logpdf_grad(mh, dh, x, g["m_pleaf"], COND[kind])
                    assert ll0 == pytest.approx(want, rel=1e-9)
                    llk, gk, _, _ = L.logpdf_grad(mh, dh, x, g["m_pleaf"], COND[kind], want_grad=True, keep_ell=True)
                    assert llk == pytest.approx(want, rel=1e-9)
                    np.testing.assert_allclose(gk, wg, rtol=1e-9, atol=1e-9 * np.abs(wg).max())
    finally:
        L.L.whale_data_destroy(dh)
        L.L.whale_model_destroy(mh)


def mixture_vs_oracle(tmp_path, n_fam=6, seed=11):
    """Mixture of two ConstantDLWGD components on synthetic families: value and gradient (w.r.t. both components'
    raw parameters and the log mixture weights) composed from the oracle's per-family outputs."""
    import whale_jl_b200 as W
    from whale_jl_b200 import synth
    from oracle import whale_oracle as wo, flat
    d = synth.generate(str(tmp_path / "mix"), n_fam, seed=seed)
    pars = ((0.1, 0.2, [0.2, 0.1], 0.67), (0.4, 0.3, [0.3, 0.05], 0.6))
    wts = np.array([0.3, 0.7])
    comps = [W.WhaleModel(W.ConstantDLWGD(lam=l, mu=m, q=q, eta=e), synth.c1_species_tree(), 0.05) for l, m, q, e in pars]
    ccd = W.read_ale(d, comps[0])
    got, gx, gw = W.logpdf_mixture_and_gradient(comps, wts, ccd)
    M, G, dC = [], [], []
    for (l, m, q, e), pj in zip(pars, wts):
        own = wo.WhaleModel(wo.ConstantDLWGD(lam=l, mu=m, q=q, eta=e), wo.c1_tree(), 0.05, condition="none")
        owr = wo.WhaleModel(wo.ConstantDLWGD(lam=l, mu=m, q=q, eta=e), wo.c1_tree(), 0.05, condition="root")
        ff = flat.FlatFams(wo.read_ale(d, own), len(own))
        tot_n, lf, g_n, gf = flat.logpdf(flat.FlatModel(own), ff, grad=True, per_family=True)
        tot_r, _, g_r, _ = flat.logpdf(flat.FlatModel(owr), ff, grad=True)
        cond = (tot_n - tot_r) / n_fam
        dC.append((g_n - g_r) / n_fam)
        M.append(lf + np.log(pj) - cond)
        G.append(gf)
    M = np.stack(M, 1)
    mx = M.max(1, keepdims=True)
    lse = mx[:, 0] + np.log(np.exp(M - mx).sum(1))
    R = np.exp(M - lse[:, None])
    want = lse.sum()
    wgx = np.stack([(R[:, j, None] * (G[j] - dC[j][None, :])).sum(0) for j in range(2)])
    assert got == pytest.approx(want, rel=1e-9)
    np.testing.assert_allclose(gx, wgx, rtol=1e-8, atol=1e-9 * np.abs(wgx).max())
    np.testing.assert_allclose(gw, R.sum(0), rtol=1e-9)
    assert W.logpdf_mixture(comps, wts, ccd) == pytest.approx(want, rel=1e-9)


def fused_track_equals_stepwise(L, n_theta=3):
    """whale_track (n_theta fused logpdf! + walk rounds) gives exactly the trees of n_theta separate
    whale_logpdf_grad(KEEP_ELL) + whale_backtrack calls with the same uniforms, and the same log-likelihoods."""
    g = load_golden("c1_example1")
    mh = L.model_create(golden_model(g))
    fl = golden_fams(g, [0, 4, 7])
    dh = L.data_create(mh, fl)
    F, MN = 3, 256
    X = np.stack([g["xs"][i % len(g["xs"])] for i in range(n_theta)])
    X[1:, -1] = np.clip(X[1:, -1] * np.linspace(0.9, 0.8, n_theta - 1), 0.05, 0.95)  # distinct draws
    U = np.random.default_rng(9).random((F, n_theta, 4 * MN))
    cnt, st, nodes, ll = L.track(mh, dh, X, g["m_pleaf"], 1, U, MN)
    assert np.all(st == 0)
    for s in range(n_theta):
        l1 = L.logpdf_grad(mh, dh, X[s], g["m_pleaf"], 1, keep_ell=True)[0]
        c1, s1, n1 = L.backtrack(mh, dh, 1, U[:, s:s + 1], MN)
        assert ll[s] == l1
        assert np.array_equal(c1[:, 0], cnt[:, s])
        for f in range(F):
            assert np.array_equal(n1[f, 0, :c1[f, 0]], nodes[f, s, :cnt[f, s]])
    L.L.whale_data_destroy(dh)
    L.L.whale_model_destroy(mh)


def nowhere_condition_vs_oracle(L):
    """NowhereExtinctCondition (src/condition.jl:31-36) on the device against the oracle's restatement of
    treepgf_allbinary (src/bdputil.jl:109-133): batch log-likelihood and gradient for the C1 model at three
    parameter points (incl. the critical λ = μ branch of the BDP pgf), DLWGD with 37 raw parameters."""
    from oracle import whale_oracle as wo
    g = load_golden("c1_example1")
    mh = L.model_create(golden_model(g))
    sel = [0, 5]
    dh = L.data_create(mh, golden_fams(g, sel))
    try:
        rng = np.random.default_rng(3)
        # moderate rates (at the C1 test point λ = μ = e the probability is ~1e-8 and the alternating sum is all
        # cancellation): the critical λ = μ branch of the BDP pgf, and two generic points
        pts = [np.concatenate([np.full(34, np.log(0.3)), [0.2, 0.1, 0.9]]),
               np.concatenate([rng.normal(np.log(0.25), 0.3, 34), [0.35, 0.6, 0.66]]),
               np.concatenate([rng.normal(np.log(0.15), 0.5, 34), [0.05, 0.9, 0.8]])]
        for xi, x in enumerate(pts):
            ll0, g0, _, _ = L.logpdf_grad(mh, dh, x, g["m_pleaf"], 0, want_grad=True)
            ll3, g3, _, _ = L.logpdf_grad(mh, dh, x, g["m_pleaf"], 3, want_grad=True)
            P = len(x)
            duals = [wo.Dual(v, np.eye(P)[i]) for i, v in enumerate(x)]
            base = wo.c1_model(condition="nowhere")
            m = base.with_rates(wo.rates_from_vector(base.rates, duals))
            wo.setmodel(m)
            c = wo.condition(m)
            assert (ll0 - ll3) / len(sel) == pytest.approx(c.v, rel=1e-9), xi
            np.testing.assert_allclose((g0 - g3) / len(sel), c.d, rtol=1e-8, atol=1e-10 * np.abs(c.d).max())
    finally:
        L.L.whale_data_destroy(dh)
        L.L.whale_model_destroy(mh)


def near_critical_vs_oracle(tmp_path, n_fam=3, seed=5):
    """Slice tables around the critical case λ ≈ μ (src/bdputil.jl:6-7): exactly critical, inside the reference's
    1e-6 `isapprox` window, just outside it (the reference's own exp(Δt(λ−μ)) − 1 cancellation regime, where k_tables
    keeps the per-slice recurrence) and far enough for the closed-form rows; log-likelihood and gradient against
    the oracle at the north-star tolerance."""
    import whale_jl_b200 as W
    from whale_jl_b200 import synth
    from oracle import whale_oracle as wo, flat
    d = synth.generate(str(tmp_path / "crit"), n_fam, seed=seed)
    ccd = None
    for lam, mu in ((0.3, 0.3), (0.3, 0.3 + 5e-7), (0.3, 0.3001), (0.3, 0.301), (0.3, 0.31), (0.25, 0.4)):
        w = W.WhaleModel(W.ConstantDLWGD(lam=lam, mu=mu, q=[0.2, 0.1], eta=0.67), synth.c1_species_tree(), 0.05)
        ccd = W.read_ale(d, w) if ccd is None else ccd
        got, gg = W.logpdf_and_gradient(w, ccd)
        ow = wo.WhaleModel(wo.ConstantDLWGD(lam=lam, mu=mu, q=[0.2, 0.1], eta=0.67), wo.c1_tree(), 0.05, condition="root")
        ff = flat.FlatFams(wo.read_ale(d, ow), len(ow))
        want, wg = flat.logpdf(flat.FlatModel(ow), ff, grad=True)[0], flat.logpdf(flat.FlatModel(ow), ff, grad=True)[2]
        assert got == pytest.approx(want, rel=1e-9), (lam, mu)
        # just outside the isapprox window the reference's own gradient is cancellation noise (∂α/∂λ loses
        # ~|λ−μ|⁻¹·1e-11 relative to one ulp of exp), so two correct implementations agree only loosely there
        d_ = abs(lam - mu)
        np.testing.assert_allclose(gg, wg, rtol=1e-5 if 1e-6 < d_ < 1e-3 else 1e-9, atol=1e-9 * np.abs(wg).max(),
                                   err_msg=f"lam {lam} mu {mu}")


def synthetic_c2_shape_vs_oracle(tmp_path, n_fam=64):
    """BASELINE config 1/2 shapes against the oracle through the package API (library chosen with lib.use)."""
    import whale_jl_b200 as W
    from whale_jl_b200 import synth
    from oracle import whale_oracle as wo, flat
    d = synth.generate(str(tmp_path / "c2"), n_fam, seed=2)
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), synth.c1_species_tree(), 0.05)
    ccd = W.read_ale(d, w)
    ll, grad = W.logpdf_and_gradient(w, ccd)
    ow = wo.WhaleModel(wo.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), wo.c1_tree(), 0.05)
    fm, ff = flat.FlatModel(ow), flat.FlatFams(wo.read_ale(d, ow), len(ow))
    tot, _, og, _ = flat.logpdf(fm, ff, grad=True)
    assert ll == pytest.approx(tot, rel=1e-9)
    np.testing.assert_allclose(grad, og, rtol=1e-9)
    # branch-wise rates (BASELINE config 2 parameterisation, P = 37)
    rng = np.random.default_rng(3)
    r = W.DLWGD(lam=list(rng.normal(np.log(0.15), 0.3, 17)), mu=list(rng.normal(np.log(0.15), 0.3, 17)),
                q=[0.2, 0.1], eta=0.67)
    wb = W.WhaleModel(r, synth.c1_species_tree(), 0.05)
    ccdb = W.read_ale(d, wb)
    ll, grad = W.logpdf_and_gradient(wb, ccdb)
    owb = wo.WhaleModel(wo.DLWGD(lam=r.lam, mu=r.mu, q=[0.2, 0.1], eta=0.67), wo.c1_tree(), 0.05)
    fm, ff = flat.FlatModel(owb), flat.FlatFams(wo.read_ale(d, owb), len(owb))
    tot, _, og, _ = flat.logpdf(fm, ff, grad=True)
    assert ll == pytest.approx(tot, rel=1e-9)
    np.testing.assert_allclose(grad, og, rtol=1e-9, atol=1e-9 * np.abs(og).max())


def c4_shape_vs_oracle(tmp_path, n_fam=3, branch_rates=True):
    """BASELINE config 3 shape (30 taxa + 5 WGD, ~2,000 clades) against the oracle through the package API."""
    import whale_jl_b200 as W
    from whale_jl_b200 import synth, newick
    from oracle import whale_oracle as wo, flat
    tree = synth.c4_species_tree()
    nws = newick.nwstr(tree, True) + ";"
    d = synth.generate(str(tmp_path / "c4"), n_fam, seed=4, tree=synth.c4_species_tree(), **synth.C4_FAMILY)
    q = [0.2, 0.1, 0.2, 0.1, 0.2]
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=q, eta=0.67), newick.readnw(nws), 0.05)
    ccd = W.read_ale(d, w)
    assert min(len(x.nleaf) for x in ccd) > 1500
    ll, grad = W.logpdf_and_gradient(w, ccd)
    ow = wo.WhaleModel(wo.ConstantDLWGD(lam=0.2, mu=0.3, q=q, eta=0.67), wo.readnw(nws), 0.05)
    fm, ff = flat.FlatModel(ow), flat.FlatFams(wo.read_ale(d, ow), len(ow))
    tot, _, og, _ = flat.logpdf(fm, ff, grad=True)
    assert ll == pytest.approx(tot, rel=1e-9)
    np.testing.assert_allclose(grad, og, rtol=1e-9, atol=1e-9 * np.abs(og).max())
    assert W.logpdf(w, ccd) == pytest.approx(tot, rel=1e-9)
    if not branch_rates:
        return
    nr = 58  # non-WGD, non-root nodes carry their own rates; the root's are unused (NaN-safe)
    rng = np.random.default_rng(5)
    r = W.DLWGD(lam=list(rng.normal(np.log(0.15), 0.3, nr + 1)), mu=list(rng.normal(np.log(0.15), 0.3, nr + 1)), q=q, eta=0.67)
    wb = W.WhaleModel(r, newick.readnw(nws), 0.05)
    ccdb = W.read_ale(d, wb)
    ll, grad = W.logpdf_and_gradient(wb, ccdb)
    owb = wo.WhaleModel(wo.DLWGD(lam=r.lam, mu=r.mu, q=q, eta=0.67), wo.readnw(nws), 0.05)
    fm, ff = flat.FlatModel(owb), flat.FlatFams(wo.read_ale(d, owb), len(owb))
    tot, _, og, _ = flat.logpdf(fm, ff, grad=True)
    assert ll == pytest.approx(tot, rel=1e-9)
    np.testing.assert_allclose(grad, og, rtol=1e-9, atol=1e-9 * np.abs(og).max())


def backtrack_uses_kept_parameters(L):
    """A model handle is shared by all wm(θ) copies: backtracking from a data handle must use the parameters of the
    evaluation that produced ITS kept ℓ, not whatever was evaluated last on the model handle (round-1 advisor finding:
    logpdf!(wm, a); logpdf(wm2, b); backtrack(wm, a) failed or sampled from the wrong distribution)."""
    g = load_golden("c1_example1")
    mh = L.model_create(golden_model(g))
    dh = L.data_create(mh, golden_fams(g, [0, 7]))
    dh2 = L.data_create(mh, golden_fams(g, [3]))
    xa = g["xs"][-1]
    xb = xa.copy()
    xb[:34] -= 1.5  # much lower rates: different tables, different q/η below
    xb[-3:] = [0.9, 0.8, 0.3]
    U = np.random.default_rng(4).random((2, 6, 4 * 256))
    try:
        L.logpdf_grad(mh, dh, xa, g["m_pleaf"], 1, keep_ell=True)
        c1, s1, n1 = L.backtrack(mh, dh, 6, U, max_nodes=256)
        L.logpdf_grad(mh, dh2, xb, g["m_pleaf"], 1, want_grad=True)  # another batch, other parameters, same model handle
        c2, s2, n2 = L.backtrack(mh, dh, 6, U, max_nodes=256)
        assert np.all(s1 == 0) and np.all(s2 == 0)
        assert np.array_equal(c1, c2) and np.array_equal(n1, n2)
    finally:
        L.L.whale_data_destroy(dh)
        L.L.whale_data_destroy(dh2)
        L.L.whale_model_destroy(mh)


def multi_device_vs_golden(L, devices):
    """whale_set_devices + whale_multi_*: the families of the C1 fixture sharded over `devices` (a device may be listed
    twice) give the full-batch golden log-likelihood, gradient and per-family values; two evaluations give identical
    bits (fixed summation order)."""
    g = load_golden("c1_example1")
    fl = golden_fams(g)
    h = L.multi_create(golden_model(g), fl, devices)
    try:
        assert L.L.whale_multi_ndev(h) == len(devices)
        assert sum(L.L.whale_multi_shard_size(h, i) for i in range(len(devices))) == fl["n_fam"]
        for kind in ("root", "none"):
            for xi, x in enumerate(g["xs"]):
                ll, grad, lf = L.multi_logpdf_grad(h, x, g["m_pleaf"], COND[kind], fl["n_fam"], want_grad=True, per_family=True)
                assert ll == pytest.approx(g[f"tot_{kind}"][xi], rel=1e-9)
                wg = g[f"grad_{kind}"][xi]
                np.testing.assert_allclose(grad, wg, rtol=1e-9, atol=1e-9 * np.abs(wg).max())
                np.testing.assert_allclose(lf, g["ll_fam"][xi], rtol=1e-9)
                ll2, grad2, _ = L.multi_logpdf_grad(h, x, g["m_pleaf"], COND[kind], fl["n_fam"], want_grad=True)
                assert ll2 == ll and np.array_equal(grad2, grad)
    finally:
        L.L.whale_multi_destroy(h)
        ids = np.zeros(0, np.int32)
        L.L.whale_set_devices(0, None)


def arena_cache_round_trip(L, tmp_path, n_fam=5):
    """whale_data_save / whale_data_load: a handle rebuilt from the binary arena cache holds the same arena bytes and
    gives bit-identical results (both gradient modes are rebuilt from the cached forward arena); a cache is refused by a
    model with another species tree or slicing; a truncated file, a flipped byte (content hash) and a file whose hash was
    "repaired" after an offset inside a node record was bent (structural validation) are errors, not crashes or
    out-of-bounds reads.  Returns (save seconds, load seconds, bytes)."""
    import struct
    import time
    import whale_jl_b200 as W
    from whale_jl_b200 import lib as wlib, synth
    from whale_jl_b200.core import _data_handle
    d = synth.generate(str(tmp_path / "cache"), n_fam, seed=17)
    mk = lambda dt=0.05: W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), synth.c1_species_tree(), dt)
    w = mk()
    a = W.read_ale(d, w)
    path = str(tmp_path / "c.arena")
    t0 = time.perf_counter()
    W.save_arena(a, w, path)
    t_save = time.perf_counter() - t0
    t0 = time.perf_counter()
    b = W.load_arena(path, w)
    t_load = time.perf_counter() - t0
    assert len(b) == len(a)
    _, dha = _data_handle(w, a)
    _, dhb = _data_handle(w, b)
    assert np.array_equal(L.arena_dump(dha), L.arena_dump(dhb))
    la, ga = W.logpdf_and_gradient(w, a)
    lb, gb = W.logpdf_and_gradient(w, b)
    assert la == lb and np.array_equal(ga, gb)
    for mode in ("rev", "fwd"):
        os.environ["WHALE_GRAD_MODE"] = mode
        try:
            w3 = mk()
            l3, g3 = W.logpdf_and_gradient(w3, W.load_arena(path, w3))
        finally:
            del os.environ["WHALE_GRAD_MODE"]
        assert l3 == pytest.approx(la, rel=1e-12)
        np.testing.assert_allclose(g3, ga, rtol=1e-9, atol=1e-12)
    # another slicing (Δt) -> another model structure -> refused
    with pytest.raises(wlib.WhaleCudaError, match="another model"):
        W.load_arena(path, mk(0.1))
    raw = open(path, "rb").read()
    nbytes = len(raw)
    open(path, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(wlib.WhaleCudaError, match="truncated"):
        W.load_arena(path, w)
    # one flipped byte in the arena: content hash
    HDR, FAMHDR = 80, 1320  # sizeof(CacheHdr), sizeof(FamHdr); the content hash is CacheHdr's last field
    F = struct.unpack_from("<Q", raw, 24)[0]
    arena_bytes = struct.unpack_from("<Q", raw, 32)[0]
    assert F == n_fam
    arena_at = HDR + F * FAMHDR
    assert arena_at + arena_bytes < nbytes
    bad = bytearray(raw)
    bad[arena_at + arena_bytes // 2] ^= 0x40
    open(path, "wb").write(bytes(bad))
    with pytest.raises(wlib.WhaleCudaError, match="corrupted"):
        W.load_arena(path, w)
    # an entry-list offset of node 0 bent far outside the blob, hash recomputed: structural validation must refuse it
    bad = bytearray(raw)
    struct.pack_into("<I", bad, arena_at + 12, 0x0FFFFFF0)  # NodeRec[0].dent_off of family 0
    h = 1469598103934665603
    for byte in bytes(bad[HDR:arena_at + arena_bytes]):
        h = ((h ^ byte) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    struct.pack_into("<Q", bad, HDR - 8, h)
    open(path, "wb").write(bytes(bad))
    with pytest.raises(wlib.WhaleCudaError, match="not a valid arena"):
        W.load_arena(path, w)
    open(path, "wb").write(b"not a cache")
    with pytest.raises(wlib.WhaleCudaError, match="not a whalecuda arena cache"):
        W.load_arena(path, w)
    return t_save, t_load, nbytes


def track_sample_and_summary(L, n_samples=24, n_theta=5):
    """whale_track_sample (per-(family, sample) posterior rows, src/track.jl:50-53) against the step-by-step sequence
    logpdf!(θ_i) + whale_backtrack with the same uniforms; the compact tree transfer against the padded one; the device-side
    tree identity / sumtrees table (src/track.jl:95-113, src/rectree.jl:113-133) against the exact host-side `treekey`
    partition; and the device random stream (no host uniforms): every tree valid, the same seed gives the same trees."""
    import whale_jl_b200 as W
    g = load_golden("c1_example1")
    mh = L.model_create(golden_model(g))
    fams = [0, 4, 7, 9]
    dh = L.data_create(mh, golden_fams(g, fams))
    F, S, MN = len(fams), n_samples, 256
    rng = np.random.default_rng(21)
    X = np.stack([g["xs"][i % len(g["xs"])] for i in range(n_theta)])
    X[:, -1] = np.clip(X[:, -1] * np.linspace(1.0, 0.8, n_theta), 0.05, 0.95)
    ti = rng.integers(n_theta, size=(F, S)).astype(np.int32)
    U = rng.random((F, S, 4 * MN))
    try:
        tot = L.track_sample(mh, dh, X, g["m_pleaf"], S, ti, U, max_nodes=MN)
        cnt, st = L.trees_counts(dh, F * S)
        assert np.all(st == 0) and int(cnt.sum()) == tot
        off, nodes = L.trees_get(dh, F * S, tot)
        off2, nodes2 = L.trees_view(dh, F * S)
        assert np.array_equal(off, off2) and np.array_equal(nodes, nodes2)
        nd, h, c, f1, th = L.trees_summary(dh, F, S, with_tree_hash=True)
        trees = [[nodes[off[f * S + s]:off[f * S + s + 1]].copy() for s in range(S)] for f in range(F)]
        # (1) step by step: for every row, logpdf! then walk with the uniforms of the samples that drew it
        for j in range(n_theta):
            L.logpdf_grad(mh, dh, X[j], g["m_pleaf"], 1, keep_ell=True)
            c1, s1, n1 = L.backtrack(mh, dh, S, U, MN)
            for f in range(F):
                for s in range(S):
                    if ti[f, s] == j:
                        assert np.array_equal(n1[f, s, :c1[f, s]], trees[f][s]), (j, f, s)
        # (2) identity classes: device hash partition == exact treekey partition; summary table == host sumtrees
        for f in range(F):
            keys = [W.treekey(t) for t in trees[f]]
            for a in range(S):
                for b in range(a + 1, S):
                    assert (keys[a] == keys[b]) == (th[f, a] == th[f, b]), (f, a, b)
            summ, _ = W.sumtrees(trees[f])
            k = int(nd[f])
            order = sorted(range(k), key=lambda i: (-int(c[f, i]), int(f1[f, i])))
            assert [int(c[f, i]) for i in order] == [x["count"] for x in summ]
            assert all(W.treekey(trees[f][int(f1[f, i])]) == x["key"] for i, x in zip(order, summ))
            assert int(c[f, :k].sum()) == S
        # (3) device random stream
        t1 = L.track_sample(mh, dh, X, g["m_pleaf"], S, ti, None, seed=77, max_nodes=MN)
        cA, sA = L.trees_counts(dh, F * S)
        oA, nA = L.trees_get(dh, F * S, t1)
        t2 = L.track_sample(mh, dh, X, g["m_pleaf"], S, ti, None, seed=77, max_nodes=MN)
        oB, nB = L.trees_get(dh, F * S, t2)
        assert np.all(sA == 0) and t1 == t2 and np.array_equal(oA, oB) and np.array_equal(nA, nB)
        t3 = L.track_sample(mh, dh, X, g["m_pleaf"], S, ti, None, seed=78, max_nodes=MN)
        oC, nC = L.trees_get(dh, F * S, t3)
        assert not (t3 == t1 and np.array_equal(nA, nC))
        nl = [int((g["f_nleaf"][g["f_clade_off"][f]:g["f_clade_off"][f + 1]] == 1).sum()) for f in fams]
        for f in range(F):  # every gene leaf exactly once per tree
            for s in range(S):
                t = nA[oA[f * S + s]:oA[f * S + s + 1]]
                leafbranch = g["m_kind"][t[:, 1]] == 0
                term = t[(t[:, 0] >= 0) & (t[:, 0] < nl[f]) & leafbranch]
                assert sorted(term[:, 0].tolist()) == list(range(nl[f]))
        # (4) walks only, from one kept ℓ, device stream
        L.logpdf_grad(mh, dh, X[0], g["m_pleaf"], 1, keep_ell=True)
        t4 = L.backtrack_device(mh, dh, S, None, seed=5, max_nodes=MN)
        c4, s4 = L.trees_counts(dh, F * S)
        assert np.all(s4 == 0) and int(c4.sum()) == t4
    finally:
        L.L.whale_data_destroy(dh)
        L.L.whale_model_destroy(mh)
