"""Parity of the CUDA path against the oracle, through the C ABI, on the B200 (-m gpu).

Bar (north star): log-likelihoods and gradients within 1e-9 relative in fp64; integer packing bit-exact."""
import os

import numpy as np
import pytest

import whale_jl_b200 as W
from whale_jl_b200 import lib as wlib, synth
from conftest import run_parity, load_golden, golden_model, golden_fams

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    lib = wlib.Lib()  # the in-tree CUDA library; raises if it was not built
    assert lib.L.whale_device_count() > 0, "no CUDA device"
    wlib.use(lib)
    return lib


def test_known_answer_single_family(L):
    """test/runtests.jl:15-19"""
    g = run_parity(L, "c1_maxn5")
    mh = L.model_create(golden_model(g))
    dh = L.data_create(mh, golden_fams(g))
    ll, *_ = L.logpdf_grad(mh, dh, g["xs"][0], g["m_pleaf"], 0)
    assert ll == pytest.approx(-60.96367806571888, rel=1e-12)


def test_known_answer_12_families(L):
    """test/runtests.jl:21-34, all conditions, three parameter points incl. the critical λ=μ branch"""
    g = run_parity(L, "c1_example1")
    assert g["tot_root"][0] == pytest.approx(-592.0185620440255, rel=1e-12)


def test_constant_rates_wgd_model(L):
    run_parity(L, "const_wgdturing")


def test_mul_tree(L):
    """test/runtests.jl:111-122"""
    run_parity(L, "mul_tree")


def test_discretisation_fixture(L):
    """test/runtests.jl:128-144 (critical getα branch on every branch)"""
    a = run_parity(L, "ex5_dt0.1")
    b = run_parity(L, "ex5_dt0.01")
    assert abs(a["tot_root"][0] - b["tot_root"][0]) < 0.1


def test_landplant_100_families(L):
    """docs/src/tutorial.md:103-131: 15..1025 clades per family (ragged sizes, the largest CCD we have)"""
    run_parity(L, "landplant100")


def test_landplant_tutorial_discretisation(L):
    """docs/src/tutorial.md:103-104,131 at the tutorial's own Δt = 0.01 (21 nodes, 3 765 slices — SURVEY §8d C1): 17 of the
    100 landplant families incl. the smallest and the largest (1 025 clades)."""
    g = run_parity(L, "landplant_dt0.01")
    assert int(g["m_nslices"].sum()) == 3765


def test_slices_tables(L):
    for name in ("c1_example1", "const_wgdturing"):
        g = load_golden(name)
        mh = L.model_create(golden_model(g))
        n = int((g["m_nslices"] + 1).sum())
        for xi, x in enumerate(g["xs"]):
            eps, phi, psi = L.slices(mh, x, g["m_pleaf"], n)
            np.testing.assert_allclose(np.stack([eps, phi, psi], 1), g["slices"][xi], rtol=1e-12)


def test_keep_ell(L):
    """logpdf! keeps the full ℓ (src/core.jl:29): compare every cell with the oracle's ℓ"""
    g = load_golden("c1_example1")
    mh = L.model_create(golden_model(g))
    dh = L.data_create(mh, golden_fams(g))
    x = g["xs"][-1]
    ll, *_ = L.logpdf_grad(mh, dh, x, g["m_pleaf"], 1, keep_ell=True)
    assert ll == pytest.approx(g["tot_root"][-1], rel=1e-9)
    for f in (0, 3):
        np.testing.assert_allclose(L.ell_get(dh, f), g[f"ell_{f}"], rtol=1e-9, atol=0)


def test_nan_root_rates_and_neg_inf(L):
    """root rates do not matter (runtests.jl:44-49); impossible data gives -Inf not an error (src/core.jl:36)"""
    g = load_golden("c1_example1")
    mh = L.model_create(golden_model(g))
    dh = L.data_create(mh, golden_fams(g))
    x = g["xs"][0].copy()
    ll0, *_ = L.logpdf_grad(mh, dh, x, g["m_pleaf"], 1)
    root = int(g["m_order"][-1])
    x[root] = np.nan
    x[17 + root] = np.nan
    ll1, *_ = L.logpdf_grad(mh, dh, x, g["m_pleaf"], 1)
    assert ll1 == ll0
    x = g["xs"][0].copy()
    x[-1] = 1.0  # η = 1: no duplication at the root -> families needing root duplications get L = 0
    x[:34] = -40.0  # and (almost) no duplication anywhere
    ll2, grad, *_ = L.logpdf_grad(mh, dh, x, g["m_pleaf"], 1, want_grad=True)
    assert ll2 == -np.inf and np.all(grad == 0.0)


def test_synthetic_c2_shape_vs_oracle(L, tmp_path):
    """BASELINE config 1 shape (9 taxa + 2 WGD, ~200 clades, constant rates) at a size the oracle does in
    seconds, through the package API (read_ale -> pack -> logpdf_and_gradient)."""
    from conftest import synthetic_c2_shape_vs_oracle
    synthetic_c2_shape_vs_oracle(tmp_path)


@pytest.mark.parametrize("stage_max", ["24576", "163840"])
def test_c4_shape_vs_oracle(L, tmp_path, monkeypatch, stage_max):
    """BASELINE config 3 shape: 30-taxon tree with 5 WGDs (64 nodes), CCDs of ~2,000 clades — lists too long to
    stage in shared memory (read in place) and a gradient computed in parameter chunks; constant rates (P = 8)
    and branch-wise rates (P = 122) against the oracle."""
    from conftest import c4_shape_vs_oracle
    monkeypatch.setenv("WHALE_STAGE_MAX", stage_max)  # lists read in place (large batches) / staged (small batches)
    c4_shape_vs_oracle(tmp_path)


def test_full_size_properties(L, tmp_path):
    """BASELINE config 1 at full size (1000 families): size-independent properties — the batch sum equals the
    sum of per-family values; shuffling families permutes per-family results and leaves the total unchanged
    (up to summation order); the gradient of the total is the sum of per-family gradients."""
    d = synth.generate(str(tmp_path / "c2full"), 1000, seed=2)
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67, ), synth.c1_species_tree(), 0.05,
                     condition="none")
    ccd = W.read_ale(d, w)
    ll, grad = W.logpdf_and_gradient(w, ccd)
    lf, gf = W.logpdf_per_family(w, ccd, grad=True)
    assert np.all(np.isfinite(lf))
    assert ll == pytest.approx(lf.sum(), rel=1e-12)
    np.testing.assert_allclose(grad, gf.sum(0), rtol=1e-11)
    perm = np.random.default_rng(0).permutation(len(ccd))
    ccd2 = W.CCDVector([ccd[i] for i in perm])
    lf2, _ = W.logpdf_per_family(w, ccd2)
    assert np.array_equal(lf2, lf[perm])  # bit-identical per family, independent of batch position
    assert W.logpdf(w, ccd2) == pytest.approx(ll, rel=1e-13)


@pytest.mark.parametrize("name", ["c1_example1", "const_wgdturing", "mul_tree"])
def test_backtrack_identical_to_oracle(L, name):
    """src/track.jl:190-414: same uniforms => identical reconciled trees (γ, e, t, parent) as the oracle, for
    every family and sample of the golden fixture (DLWGD with 2 WGDs, constant rates, MUL tree)."""
    g = load_golden(name)
    seed, F, nbt, stride = (int(v) for v in g["bt_seed"])
    U = np.random.default_rng(seed).random((F, nbt, stride))
    starts = np.concatenate([[0], np.cumsum(g["bt_counts"][:, 0])])
    mh = L.model_create(golden_model(g))
    dh = L.data_create(mh, golden_fams(g))
    L.logpdf_grad(mh, dh, g["xs"][-1], g["m_pleaf"], 1, keep_ell=True)
    cnt, st, nodes = L.backtrack(mh, dh, nbt, U, max_nodes=256)
    assert np.all(st == 0)
    for f in range(F):
        for s in range(nbt):
            want = g["bt_nodes"][starts[f * nbt + s]:starts[f * nbt + s + 1]]
            assert np.array_equal(nodes[f, s, :cnt[f, s]], want), (name, f, s)


def test_backtrack_api_and_invariants(L, tmp_path):
    """Package-level `backtrack` on synthetic families: every gene leaf exactly once per tree, parents precede
    children, loss nodes carry t = 0; needs logpdf_ first (WHALE_ERR_STATE otherwise)."""
    d = synth.generate(str(tmp_path / "bt"), 16, seed=5)
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), synth.c1_species_tree(), 0.05)
    ccd = W.read_ale(d, w)
    with pytest.raises(wlib.WhaleCudaError):
        W.backtrack(w, ccd, n_samples=2, seed=1)
    W.logpdf_(w, ccd)
    trees = W.backtrack(w, ccd, n_samples=8, seed=1)
    assert len(trees) == 16 and len(trees[0]) == 8
    for f, fam in enumerate(trees):
        nl = len(ccd[f].leaves)
        for t in fam:
            leafbranch = w.kind[t[:, 1]] == 0
            term = t[(t[:, 0] >= 0) & (t[:, 0] < nl) & leafbranch]
            assert sorted(term[:, 0].tolist()) == list(range(nl))
            assert np.all(t[1:, 3] < np.arange(1, len(t))) and t[0, 3] == -1
            assert np.all(t[t[:, 0] < 0, 2] == 0)


def test_mixture_on_device_with_gradient(L, tmp_path):
    """src/core.jl:66-76 on the device (whale_mixture_logpdf_grad): value, ∂/∂ raw parameters of every component and
    ∂/∂ log weights against the oracle's per-family outputs."""
    from conftest import mixture_vs_oracle
    mixture_vs_oracle(tmp_path, n_fam=24)


def test_native_read_ale(L, tmp_path):
    """whale_read_ale (native parse + build + pack) gives the arena and results of the Python read_ale path."""
    from whale_jl_b200.core import _data_handle
    d = synth.generate(str(tmp_path / "nat"), 40, seed=21)
    w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1], eta=0.67), synth.c1_species_tree(), 0.05)
    a, b = W.read_ale(d, w), W.read_ale_native(d, w)
    _, dha = _data_handle(w, a)
    _, dhb = _data_handle(w, b)
    assert np.array_equal(L.arena_dump(dha), L.arena_dump(dhb))
    la, ga = W.logpdf_and_gradient(w, a)
    lb, gb = W.logpdf_and_gradient(w, b)
    assert la == lb and np.array_equal(ga, gb)


def test_nowhere_extinct_condition(L):
    """NowhereExtinctCondition (src/condition.jl:31-36, 2^9 inclusion–exclusion terms) on the device."""
    from conftest import nowhere_condition_vs_oracle
    nowhere_condition_vs_oracle(L)


def test_fused_track_equals_stepwise(L):
    """src/track.jl:47-63 fused on the device (whale_track) against the step-by-step C-ABI sequence."""
    from conftest import fused_track_equals_stepwise
    fused_track_equals_stepwise(L, n_theta=5)


def test_mixture_modelarray_and_track(L, tmp_path):
    """src/core.jl:66-79 (mixture / ModelArray on top of per-family device outputs) against the oracle, and the
    `track` driver loop (src/track.jl:30-63)."""
    from oracle import whale_oracle as wo
    d = synth.generate(str(tmp_path / "mix"), 6, seed=11)
    tree = synth.c1_species_tree
    comps = [W.WhaleModel(W.ConstantDLWGD(lam=l, mu=m, q=[0.2, 0.1], eta=0.67), tree(), 0.05) for l, m in
             ((0.1, 0.2), (0.4, 0.3))]
    ccd = W.read_ale(d, comps[0])
    for c in comps:
        W.set_probe(c, ccd[0])
    got = W.logpdf_mixture(comps, [0.3, 0.7], ccd)
    ocomps = [wo.WhaleModel(wo.ConstantDLWGD(lam=l, mu=m, q=[0.2, 0.1], eta=0.67), wo.c1_tree(), 0.05) for l, m in
              ((0.1, 0.2), (0.4, 0.3))]
    occd = wo.read_ale(d, ocomps[0])
    M = np.array([[wo.logpdf(om, x) + np.log(p) - wo.condition(om) for om, p in zip(ocomps, (0.3, 0.7))] for x in occd])
    want = float(np.sum(np.log(np.exp(M - M.max(1, keepdims=True)).sum(1)) + M.max(1)))
    assert got == pytest.approx(want, rel=1e-9)
    assert W.condition(comps[0]) == pytest.approx(wo.condition(ocomps[0]), rel=1e-9)
    ma = W.logpdf_modelarray([comps[i % 2] for i in range(6)], ccd)
    assert ma == pytest.approx(sum(wo.logpdf(ocomps[i % 2], occd[i]) for i in range(6)), rel=1e-9)
    post = [dict(lam=0.15, mu=0.2, q=[0.1, 0.3], eta=0.7), dict(lam=0.3, mu=0.25, q=[0.5, 0.05], eta=0.6)]
    trees = W.track(comps[0], ccd, post, 4, seed=3)
    assert len(trees) == 6 and all(len(t) == 4 for t in trees)
    assert all(t[0][0, 3] == -1 for t in trees)


def test_near_critical_rates(L, tmp_path):
    """λ ≈ μ: the reference's isapprox window, its cancellation regime just outside it (k_tables keeps the
    per-slice recurrence there) and the closed-form rows, against the oracle."""
    from conftest import near_critical_vs_oracle
    near_critical_vs_oracle(tmp_path, n_fam=6)


@pytest.mark.parametrize("mode", ["fwd", "rev"])
def test_gradient_modes(L, tmp_path, monkeypatch, mode):
    """Both gradient paths against the oracle on every fixture: k_dp's forward tangents (WHALE_GRAD_MODE=fwd, parameter
    chunks where needed) and the reverse-mode kernel k_dp_rev (=rev; the default picks per data set, see
    whale_data_grad_mode)."""
    monkeypatch.setenv("WHALE_GRAD_MODE", mode)
    g = load_golden("c1_example1")
    mh = L.model_create(golden_model(g))
    dh = L.data_create(mh, golden_fams(g))
    assert L.L.whale_data_grad_mode(dh) == (1 if mode == "rev" else 0)
    L.L.whale_data_destroy(dh)
    L.L.whale_model_destroy(mh)
    for name in ("c1_maxn5", "c1_example1", "const_wgdturing", "mul_tree", "landplant100", "ex5_dt0.01", "landplant_dt0.01"):
        run_parity(L, name)
    from conftest import synthetic_c2_shape_vs_oracle, near_critical_vs_oracle, nowhere_condition_vs_oracle, mixture_vs_oracle
    synthetic_c2_shape_vs_oracle(tmp_path)
    near_critical_vs_oracle(tmp_path, n_fam=4)
    nowhere_condition_vs_oracle(L)
    mixture_vs_oracle(tmp_path, n_fam=8)


def test_reverse_mode_one_persistent_cta(L, monkeypatch):
    """k_dp_rev with a single persistent CTA: all families go through one CTA's loop (history slot and shared-memory
    carve-up reused from family to family) — same per-family results as the default launch."""
    monkeypatch.setenv("WHALE_GRAD_MODE", "rev")
    monkeypatch.setenv("WHALE_REV_GRID", "1")
    run_parity(L, "c1_example1", conds=["root"])
    run_parity(L, "landplant100", sel=list(range(0, 100, 7)), conds=["root"])


def test_backtrack_uses_kept_parameters(L):
    from conftest import backtrack_uses_kept_parameters
    backtrack_uses_kept_parameters(L)


def test_multi_device_handle(L):
    """whale_set_devices(n, ids[]) + one handle sharded over the devices (all GPUs of the box; on a one-GPU
    box two shards on device 0) against the full-batch golden values."""
    from conftest import multi_device_vs_golden
    n = L.L.whale_device_count()
    multi_device_vs_golden(L, list(range(n)) if n > 1 else [0, 0])


@pytest.mark.parametrize("graphs", ["auto", "1", "0"])
def test_repeated_calls_graph_replay(graphs):
    """Host-pointer evaluations are replayed as one CUDA graph from the third call with the same (flags, condition) on
    (automatic for single-bin handles; WHALE_GRAPHS=1 / 0 forces it): repeated gradient / value / ℓ-keeping calls with
    changing θ on one handle stay on the golden values in every mode, reverse (C1, P = 37) and forward (constant rates)
    kernels.  Process-wide switch: child processes."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from whale_jl_b200 import lib as wlib\nfrom conftest import repeated_calls_parity\n"
            "L = wlib.Lib()\nrepeated_calls_parity(L, 'c1_example1')\nrepeated_calls_parity(L, 'const_wgdturing')\nprint('ok')\n"
            % (os.path.dirname(here), here))
    env = dict(os.environ)
    env.pop("WHALE_GRAPHS", None)
    if graphs != "auto":
        env["WHALE_GRAPHS"] = graphs
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout[-1500:] + out.stderr[-3000:]


@pytest.mark.parametrize("env", [{}, {"WHALE_PEER_FUSE": "0"}, {"WHALE_GRAD_MODE": "fwd"}])
def test_peer_sum_between_processes(tmp_path, env):
    """SURVEY §8e, one process per GPU: whale_peer_export / whale_peer_import + WHALE_PEER_SUM between two (one-GPU box:
    both on device 0, time-sliced) or four processes; every rank must see the full-batch golden total with identical bits.
    Default = the exchange in the tail of the DP kernel; WHALE_PEER_FUSE=0 = its own launch; forward-tangent kernel."""
    import subprocess
    import sys
    import glob
    world = 2
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "peer_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), str(tmp_path)], env=dict(os.environ, **env),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    for r, p in enumerate(procs):
        assert p.returncode == 0 and f"rank {r} ok" in outs[r], outs[r][-3000:]
    for f0 in glob.glob(str(tmp_path / "ll_r0_*.txt")):  # same bits on every rank (fixed summation order)
        f1 = f0.replace("ll_r0_", "ll_r1_")
        assert open(f0).read() == open(f1).read(), (f0, open(f0).read(), open(f1).read())


def test_arena_cache_round_trip_on_device(L, tmp_path):
    """whale_data_save / whale_data_load on the B200: byte-identical arena, bit-identical results, refusal of
    foreign / truncated / corrupted / structurally invalid caches; prints the save and load times."""
    from conftest import arena_cache_round_trip
    t_save, t_load, nbytes = arena_cache_round_trip(L, tmp_path, n_fam=200)
    print(f"arena cache: {nbytes / 1e6:.1f} MB for 200 families, save {t_save * 1e3:.1f} ms, load + repack plans {t_load * 1e3:.1f} ms")


@pytest.mark.parametrize("slots", ["16", "3", "1"])
def test_track_sample_and_summary(L, monkeypatch, slots):
    """src/track.jl:47-63 with per-(family, sample) posterior rows, compact tree transfer, device tree identity / sumtrees.
    slots = posterior draws per batch: one batch / three batches (last one short) / draw by draw."""
    monkeypatch.setenv("WHALE_TRACK_SLOTS", slots)
    from conftest import track_sample_and_summary
    track_sample_and_summary(L, n_samples=64, n_theta=7)


def test_c4_shape_16_families(L, tmp_path):
    """BASELINE config 3 shape at 16 families (~2,000 clades each, 30 taxa + 5 WGDs): constant rates (P = 8) and
    branch-wise rates (P = 124), both through the reverse-mode kernel, against the oracle."""
    from conftest import c4_shape_vs_oracle
    c4_shape_vs_oracle(tmp_path, n_fam=16)
