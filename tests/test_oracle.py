"""The oracle against everything the reference pins for the hot path — CPU only."""
import numpy as np
import pytest

from conftest import HAVE_REF, REF, load_golden
from oracle import flat, whale_oracle as wo

needs_ref = pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted (GPU box)")


@needs_ref
def test_known_answer_single_family_maxn5():
    """test/runtests.jl:15-19"""
    w = wo.c1_model(maxn=5)
    ccd = wo.read_ale(f"{REF}/example/example-1/ale", w)
    assert wo.logpdf(w, ccd[0]) == pytest.approx(-60.96367806571888, rel=1e-14)


@needs_ref
def test_known_answer_12_families_rootcondition():
    """test/runtests.jl:21-34 (logpdf and logpdf! agree)"""
    w = wo.c1_model()
    ccd = wo.read_ale(f"{REF}/example/example-1/ale", w)
    assert wo.logpdf(w, ccd) == pytest.approx(-592.0185620440255, rel=1e-14)
    assert wo.logpdf(w, ccd, keep=True) == pytest.approx(-592.0185620440255, rel=1e-14)


@needs_ref
def test_root_rates_have_no_effect():
    """test/runtests.jl:44-49"""
    w = wo.c1_model()
    ccd = wo.read_ale(f"{REF}/example/example-1/ale", w)
    r = w.rates
    root = w.root.id
    lam, mu = list(r.lam), list(r.mu)
    lam[root - 1] = float("nan")
    mu[root - 1] = float("nan")
    w2 = w.with_rates(wo.DLWGD(lam=lam, mu=mu, q=r.q, eta=r.eta))
    assert wo.logpdf(w2, ccd[:3]) == pytest.approx(wo.logpdf(w, ccd[:3]), rel=1e-15)


@needs_ref
def test_model_structure_c1():
    """Node order / ids / wgd ids (src/model.jl:96-145) incl. NewickTree's append-on-insert (runtests.jl:54-55)."""
    w = wo.c1_model()
    rows = [(n.name, n.id, n.wgdid, n.n) for n in w.order]
    assert rows[:9] == [("MPOL", 1, 0, 96), ("PPAT", 2, 0, 96), ("SMOE", 3, 0, 90), ("GBIL", 4, 0, 64),
                        ("PABI", 5, 0, 64), ("OSAT", 6, 0, 32), ("CPAP", 7, 0, 12), ("ATHA", 8, 0, 6),
                        ("ATRI", 9, 0, 46)]
    assert [(r[1], r[2]) for r in rows[9:]] == [(10, 0), (11, 0), (18, 2), (12, 0), (13, 0), (14, 0), (19, 1),
                                                 (15, 0), (16, 0), (17, 0)]
    assert sum(r[3] for r in rows) == 618
    # the reference's own test (runtests.jl:54-55) needs t[1][2] to be the inserted node
    t = wo.readnw(wo.EXTREE)
    wo.insertnode(t[1][1], name="wgd_1")
    assert t[1][2].name == "wgd_1" and not t[1][2].isleaf()
    wo.insertnode(t[1][2][1], name="wgd_2")


def test_cpp_oracle_matches_known_answers():
    g = load_golden("c1_maxn5")
    assert g["tot_none"][0] == pytest.approx(float(g["known_logpdf"]), rel=1e-13)
    g = load_golden("c1_example1")
    assert g["tot_root"][0] == pytest.approx(float(g["known_logpdf"]), rel=1e-13)


@needs_ref
def test_cpp_oracle_matches_python_oracle():
    w = wo.c1_model()
    ccd = wo.read_ale(f"{REF}/example/example-1/ale", w)[:3]
    fm, ff = flat.FlatModel(w), flat.FlatFams(ccd, len(w))
    g = load_golden("c1_example1")
    x = g["xs"][2]
    tot, ll, grad, _ = flat.logpdf(fm, ff, x=x, grad=True)
    v, gp = wo.logpdf_and_gradient(w, ccd, list(x))
    assert tot == pytest.approx(v, rel=1e-13)
    np.testing.assert_allclose(grad, gp, rtol=1e-10, atol=1e-12)
    # full ℓ and backtracking with an explicit uniform stream
    w2 = w.with_rates(wo.rates_from_vector(w.rates, list(x)))
    wo.logpdf(w2, ccd[0], keep=True)
    mats, _ = flat.ell(fm, ff, 0, x=x)
    for e in range(fm.nn):
        a = np.array(ccd[0].ell[e + 1])
        if a.size:
            np.testing.assert_allclose(mats[e], a, rtol=1e-13, atol=0)
    for seed in range(3):
        u = np.random.default_rng(seed).random(2048)
        nodes, used = wo.backtrack(w2, ccd[0], u)
        n, arr, used2 = flat.backtrack(fm, ff, 0, u, x=x)
        assert n == len(nodes) and used == used2
        assert np.array_equal(arr, np.array([(a - 1, b - 1, t, p) for (a, b, t, p) in nodes]))


def test_gradient_matches_finite_differences():
    """ForwardDiff values are not pinned by the reference (runtests.jl:39 only checks finiteness):
    pin the dual-number pass with central differences, away from and at the critical λ=μ branch."""
    g = load_golden("const_wgdturing")
    from conftest import golden_fams
    import ctypes as C

    class M:  # rebuild a FlatModel-like object from the golden arrays
        pass
    fm = _flat_from_golden(g)
    ff = _fams_from_golden(g)
    for x in (g["xs"][1],):
        _, _, grad, _ = flat.logpdf(fm, ff, x=x, grad=True)
        fd = np.zeros_like(x)
        for i in range(len(x)):
            h = 1e-6
            xp, xm = x.copy(), x.copy()
            xp[i] += h
            xm[i] -= h
            fd[i] = (flat.logpdf(fm, ff, x=xp)[0] - flat.logpdf(fm, ff, x=xm)[0]) / (2 * h)
        np.testing.assert_allclose(grad, fd, rtol=2e-6, atol=1e-6)


def test_backtracking_invariants():
    """Every gene leaf appears exactly once; loss nodes have γ=-1,t=0; children follow parents."""
    g = load_golden("c1_example1")
    fm, ff = _flat_from_golden(g), _fams_from_golden(g)
    x = g["xs"][-1]
    nleaves = int((g["f_nleaf"][g["f_clade_off"][0]:g["f_clade_off"][1]] == 1).sum())
    for seed in range(5):
        n, arr, used = flat.backtrack(fm, ff, 0, np.random.default_rng(seed).random(4096), x=x)
        assert n > 0
        leafkind = fm.kind[arr[:, 1]] == 0
        term = arr[(arr[:, 0] >= 0) & leafkind & (arr[:, 0] < nleaves)]
        assert sorted(term[:, 0].tolist()) == list(range(nleaves))
        assert np.all(arr[1:, 3] < np.arange(1, n)) and arr[0, 3] == -1
        assert np.all(arr[arr[:, 0] < 0, 2] == 0)


def test_discretisation_stability():
    """test/runtests.jl:128-144: ℓ within 0.1 between Δt = 0.1 and 0.01 (critical getα branch, MUL tree)."""
    a, b = load_golden("ex5_dt0.1"), load_golden("ex5_dt0.01")
    assert abs(a["tot_root"][0] - b["tot_root"][0]) < 0.1


def _flat_from_golden(g):
    import types
    from oracle.flat import OModel, _p, i32p, f64p
    fm = types.SimpleNamespace()
    keep = {k: np.ascontiguousarray(g[k]) for k in g if k.startswith("m_")}
    fm.keep = keep
    fm.nn = len(keep["m_order"])
    fm.P = int(g["m_P"])
    fm.kind = keep["m_kind"]
    fm.nslices = keep["m_nslices"]
    fm.x = g["xs"][0]
    fm.c = OModel(fm.nn, _p(keep["m_order"], i32p), _p(keep["m_child0"], i32p), _p(keep["m_child1"], i32p),
                  _p(keep["m_kind"], i32p), _p(keep["m_nslices"], i32p), _p(keep["m_dt"], f64p),
                  _p(keep["m_leafP"], f64p), _p(keep["m_pleaf"], f64p), _p(keep["m_lam_slot"], i32p),
                  _p(keep["m_mu_slot"], i32p), _p(keep["m_q_slot"], i32p), int(g["m_eta_slot"]),
                  int(g["m_log_scale"]), 1)
    fm.row_off = np.concatenate([[0], np.cumsum(fm.nslices + 1)]).astype(np.int64)
    return fm


def _fams_from_golden(g):
    import types
    from oracle.flat import OFams, _p, i32p, i64p, f64p
    ff = types.SimpleNamespace()
    keep = {k: np.ascontiguousarray(g[k]) for k in g if k.startswith("f_")}
    ff.keep = keep
    ff.F = len(keep["f_clade_off"]) - 1
    ff.nn = len(g["m_order"])
    ff.compat_off = keep["f_compat_off"]
    ff.ncompat = lambda f, e: int(ff.compat_off[f * ff.nn + e + 1] - ff.compat_off[f * ff.nn + e])
    ff.c = OFams(ff.F, _p(keep["f_clade_off"], i64p), _p(keep["f_nleaf"], i32p), _p(keep["f_split_off"], i64p),
                 _p(keep["f_g1"], i32p), _p(keep["f_g2"], i32p), _p(keep["f_p"], f64p),
                 _p(keep["f_compat_off"], i64p), _p(keep["f_compat"], i32p))
    return ff


def test_nowhere_extinct_condition_reference_tests():
    """test/runtests.jl:52-77: the probability of being extinct nowhere grows with the retention rates, and agrees
    with a Monte-Carlo simulation of the DL+WGD process (the reference allows 20 %; this uses 200k lineages, 3 %)."""
    p = -np.inf
    for q in np.arange(0.0, 1.01, 0.1):
        w = wo.WhaleModel(wo.ConstantDLWGD(lam=0.3, mu=0.4, q=[q, q], eta=0.66), wo.c1_tree(), 0.05, condition="nowhere")
        c = wo.condition(w)
        assert c > p
        p = c
    rng = np.random.default_rng(0)
    lam, mu, q, eta, N = 0.3, 0.4, [0.35, 0.6], 0.66, 200000
    w = wo.WhaleModel(wo.ConstantDLWGD(lam=lam, mu=mu, q=q, eta=eta), wo.c1_tree(), 0.05, condition="nowhere")

    def evolve(n, t):  # linear BDP transition: each lineage dies out w.p. α, else leaves Geometric(1−β) copies
        a = float(wo.getalpha(lam, mu, t))
        b = lam / mu * a
        surv = rng.binomial(n, 1.0 - a)
        return surv + rng.negative_binomial(np.maximum(surv, 1), 1.0 - b) * (surv > 0)

    def walk(node, n):
        if node is not w.root:
            n = evolve(n, node.dist)
        if node.isleaf():
            return n > 0
        if node.iswgd():
            n = n + rng.binomial(n, q[node.wgdid - 1])
        ok = np.ones(N, bool)
        for c in node.children:
            ok &= walk(c, n)
        return ok

    ok = walk(w.root, rng.geometric(eta, N))
    assert np.exp(wo.condition(w)) == pytest.approx(ok.mean(), rel=0.03)


def test_nowhere_extinct_condition_gradient_by_differences():
    # moderate rates: at λ ≈ μ ≈ e (the C1 test point) the probability is ~1e-8 and central differences of the
    # alternating 2^9-term sum are pure cancellation noise
    x0 = np.log(0.3) + np.linspace(-0.3, 0.2, 37)
    x0[-3:] = [0.3, 0.15, 0.8]

    def cond(x):
        base = wo.c1_model(condition="nowhere")
        m = base.with_rates(wo.rates_from_vector(base.rates, list(x)))
        wo.setmodel(m)
        return wo.condition(m)

    P = len(x0)
    c = cond([wo.Dual(v, np.eye(P)[i]) for i, v in enumerate(x0)])
    for i in (0, 5, 16, 17, 30, 34, 35, 36):
        h = 1e-5
        xp, xm = x0.copy(), x0.copy()
        xp[i] += h
        xm[i] -= h
        assert c.d[i] == pytest.approx((cond(xp) - cond(xm)) / (2 * h), rel=1e-5, abs=1e-8)


def test_slice_recursion_equals_the_bdp_pgf_off_the_critical_branch():
    """SURVEY §8c gap: both reference known answers use λ = μ, so the general branch of `getα` (src/bdputil.jl:6-7) is
    pinned by the formula text alone.  Cross-check it against the reference's OTHER statement of the same process:
    the per-slice ϵ recursion over a whole branch (src/model.jl:182-191) must equal the linear-BDP pgf at the branch
    length, pgf(LinearBDP(λ, μ, t), ϵ₀) (src/bdputil.jl:58-64) — and stay continuous across the 1e-6 window."""
    rng = np.random.default_rng(5)
    for lam, mu in [(0.2, 0.3), (0.45, 0.1), (1.3, 1.1), (0.3, 0.3 + 2e-6), (0.3, 0.3 - 2e-6)] + \
                   [tuple(rng.uniform(0.05, 1.5, 2)) for _ in range(5)]:
        w = wo.WhaleModel(wo.ConstantDLWGD(lam=lam, mu=mu, q=[0.2, 0.1], eta=0.67), wo.c1_tree(), 0.05)
        wo.setmodel(w)
        for n in w.order:
            if n.isroot() or len(n) == 1:
                continue
            want = wo.bdp_pgf(lam, mu, n.dist, n.eps[0])
            # just outside the window exp(Δt(λ−μ)) − 1 cancels: the reference's recursion itself carries ~1e-16/(Δt|λ−μ|)
            # per slice (2.5e-10 over a 64-slice branch at |λ−μ| = 2e-6) — why k_tables keeps the recurrence there
            assert n.eps[-1] == pytest.approx(want, rel=1e-8 if abs(lam - mu) < 1e-5 else 1e-12), (lam, mu, n.name)
    # continuity of getα across the isapprox window (critical formula on one side, general on the other)
    a_in, a_out = wo.getalpha(0.3, 0.3 + 0.9e-6, 0.05), wo.getalpha(0.3, 0.3 + 1.1e-6, 0.05)
    assert a_out == pytest.approx(a_in, rel=1e-5)  # the window itself is a 4e-6 step (λt/(1+λt) ignores μ)
