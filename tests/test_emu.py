"""Kernel logic exercised on the CPU through the emulation build (tests/emu): the same whalecuda.cu compiled
against a host-thread CUDA shim.  Catches packer / indexing / tangent-plan bugs without a GPU; the parity tests
proper are tests/test_gpu_parity.py (-m gpu)."""
import os
import subprocess

import numpy as np
import pytest

from whale_jl_b200 import lib as wlib
from conftest import ROOT, run_parity, load_golden, golden_model, golden_fams

# WHALE_EMU_LIB: run this file against another emulation build (tools/emu_asan.sh: AddressSanitizer)
EMU = os.environ.get("WHALE_EMU_LIB") or os.path.join(ROOT, "tests", "emu", "libwhalecuda_emu.so")


@pytest.fixture(scope="module")
def L():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s"])
    return wlib.Lib(EMU)


def test_emu_known_answer_maxn5(L):
    g = run_parity(L, "c1_maxn5")
    assert g["tot_none"][0] == pytest.approx(-60.96367806571888, rel=1e-12)


def test_emu_c1_subset(L):
    run_parity(L, "c1_example1", sel=[0, 3], conds=["root"])


def test_emu_constant_rates(L):
    run_parity(L, "const_wgdturing", sel=[1, 7], conds=["nonextinct"])


def test_emu_mul_tree(L):
    run_parity(L, "mul_tree", sel=[2], conds=["root"])


def test_emu_slices_and_ell(L):
    g = load_golden("c1_example1")
    mh = L.model_create(golden_model(g))
    x = g["xs"][2]
    n = int((g["m_nslices"] + 1).sum())
    eps, phi, psi = L.slices(mh, x, g["m_pleaf"], n)
    np.testing.assert_allclose(np.stack([eps, phi, psi], 1), g["slices"][2], rtol=1e-12)
    dh = L.data_create(mh, golden_fams(g, [3]))
    L.logpdf_grad(mh, dh, x, g["m_pleaf"], 1, keep_ell=True)
    np.testing.assert_allclose(L.ell_get(dh, 0), g["ell_3"], rtol=1e-11, atol=0)


def _check_backtrack(L, name, fams, nbt_check=None):
    """Backtracked trees must be identical to the oracle's for the same uniforms (golden bt_* arrays)."""
    g = load_golden(name)
    seed, F, nbt, stride = (int(v) for v in g["bt_seed"])
    U = np.random.default_rng(seed).random((F, nbt, stride))
    counts = g["bt_counts"].reshape(F, nbt, 2)
    starts = np.concatenate([[0], np.cumsum(g["bt_counts"][:, 0])])
    mh = L.model_create(golden_model(g))
    dh = L.data_create(mh, golden_fams(g, fams))
    L.logpdf_grad(mh, dh, g["xs"][-1], g["m_pleaf"], 1, keep_ell=True)
    cnt, st, nodes = L.backtrack(mh, dh, nbt, U[fams], max_nodes=256)
    assert np.all(st == 0)
    for i, f in enumerate(fams):
        for s in range(nbt):
            want = g["bt_nodes"][starts[f * nbt + s]:starts[f * nbt + s + 1]]
            assert cnt[i, s] == counts[f, s, 0]
            assert np.array_equal(nodes[i, s, :cnt[i, s]], want), (name, f, s)
        import whale_jl_b200 as W  # sumtrees over the family's samples (src/rectree.jl:113-133)
        summ, clades = W.sumtrees([nodes[i, s, :cnt[i, s]] for s in range(nbt)])
        assert sum(x["count"] for x in summ) == nbt and summ[0]["count"] >= summ[-1]["count"]
        assert max(clades.values()) == nbt  # the root clade's reconciliation node is in every sample


def test_emu_backtrack_matches_oracle(L):
    _check_backtrack(L, "c1_example1", [0, 7])
    _check_backtrack(L, "const_wgdturing", [2])
    _check_backtrack(L, "mul_tree", [5])


def test_emu_parameter_chunks(L, monkeypatch):
    """Families whose full-tangent working set exceeds the shared-memory goal get their gradient in several
    passes over parameter chunks; force that with a tiny goal and compare with the oracle (37 parameters)."""
    monkeypatch.setenv("WHALE_GRAD_MODE", "fwd")
    monkeypatch.setenv("WHALE_SMEM_GOAL", "25000")  # -> 6 passes
    run_parity(L, "c1_example1", sel=[3], conds=["root"])
    monkeypatch.setenv("WHALE_SMEM_GOAL", "6000")  # -> one pass per parameter (37)
    run_parity(L, "c1_example1", sel=[3], conds=["root"])


def test_emu_unstaged_lists(L, monkeypatch):
    """Families whose lists exceed the staging limit (large CCDs, BASELINE config 3) are read from global memory
    in place; force that path with a tiny limit."""
    monkeypatch.setenv("WHALE_STAGE_MAX", "64")
    run_parity(L, "c1_example1", sel=[5], conds=["root"])
    _check_backtrack(L, "c1_example1", [5])


def test_emu_mixture(L, tmp_path):
    from conftest import mixture_vs_oracle
    wlib.use(L)
    try:
        mixture_vs_oracle(tmp_path, n_fam=3)
    finally:
        wlib.use(None)


def test_emu_fused_track(L):
    from conftest import fused_track_equals_stepwise
    fused_track_equals_stepwise(L)


def test_emu_nowhere_extinct_condition(L):
    from conftest import nowhere_condition_vs_oracle
    nowhere_condition_vs_oracle(L)


def test_emu_native_read_ale_is_bit_identical(L, tmp_path):
    """whale_read_ale (C++ parser + CCD builder + packer, all host threads) produces byte-for-byte the arena that
    the Python read_ale -> whale_data_create path produces: synthetic families on the C1 tree, and (when the
    reference checkout is present) its example-1 / landplant / MUL-tree fixtures."""
    import whale_jl_b200 as W
    from whale_jl_b200 import synth, newick
    from conftest import HAVE_REF, REF
    wlib.use(L)
    try:
        cases = [(synth.generate(str(tmp_path / "nat"), 12, seed=21), synth.c1_species_tree(), 0.05)]
        if HAVE_REF:
            cases.append((os.path.join(REF, "example", "example-1", "ale"), synth.c1_species_tree(), 0.05))
            t5 = newick.readnw(open(os.path.join(REF, "example", "example-5", "tree.nw")).read().strip())
            lst = tmp_path / "ex5.txt"  # a text file listing paths (src/ccd.jl:131-133)
            lst.write_text("".join(os.path.join(REF, "example", "example-5", f) + "\n" for f in ("OG0006450.ale", "OG0014587.ale")))
            cases.append((str(lst), t5, 0.1))
            cases.append((os.path.join(REF, "docs", "data", "landplant", "100fams"), None, 0.05))
        for d, tree, dt in cases:
            if tree is None:  # landplant fixture: 100 families, 15..1025 clades
                tree = newick.readnw(open(os.path.join(REF, "docs", "data", "landplant", "speciestree.nw")).read().strip())
            w = W.WhaleModel(W.ConstantDLWGD(lam=0.2, mu=0.3, q=[0.2, 0.1][:sum(1 for n in newick.postwalk(tree) if n.name.startswith("wgd"))], eta=0.67), tree, dt)
            a = W.read_ale(d, w)
            b = W.read_ale_native(d, w, n_threads=3)
            from whale_jl_b200.core import _data_handle
            mh, dha = _data_handle(w, a)
            _, dhb = _data_handle(w, b)
            assert len(a) == len(b) and [len(x.nleaf) for x in a] == b.n_clades.tolist()
            assert np.array_equal(L.arena_dump(dha), L.arena_dump(dhb))
            if len(a) <= 12:  # the emulated kernels are slow: evaluate only the small batches
                assert W.logpdf(w, b) == W.logpdf(w, a)
    finally:
        wlib.use(None)


def test_emu_near_critical_rates(L, tmp_path):
    from conftest import near_critical_vs_oracle
    wlib.use(L)
    try:
        near_critical_vs_oracle(tmp_path)
    finally:
        wlib.use(None)


def test_emu_landplant_100_families(L):
    """docs/src/tutorial.md:103-131 fixture: 100 families of 15..1025 clades, 21 nodes, 3,765 slices (dt = 0.01)."""
    run_parity(L, "landplant100")


def test_emu_discretisation_fixture(L):
    """test/runtests.jl:128-144"""
    a = run_parity(L, "ex5_dt0.1")
    b = run_parity(L, "ex5_dt0.01")
    assert abs(a["tot_root"][0] - b["tot_root"][0]) < 0.1


def test_emu_synthetic_c2_shape(L, tmp_path):
    """BASELINE config 1/2 shapes (~200-clade families: full 125-lane rows, constant and branch-wise rates)."""
    from conftest import synthetic_c2_shape_vs_oracle
    wlib.use(L)
    try:
        synthetic_c2_shape_vs_oracle(tmp_path, n_fam=12)
    finally:
        wlib.use(None)


@pytest.mark.parametrize("stage_max", ["24576", "163840"])
def test_emu_c4_shape(L, tmp_path, monkeypatch, stage_max):
    """BASELINE config 3 shape (64 nodes, ~2,000 clades): lists read in place / staged, constant rates (P = 8)."""
    from conftest import c4_shape_vs_oracle
    monkeypatch.setenv("WHALE_STAGE_MAX", stage_max)
    wlib.use(L)
    try:
        c4_shape_vs_oracle(tmp_path, n_fam=1, branch_rates=False)  # P = 122 chunking: GPU test; P = 37 here above
    finally:
        wlib.use(None)


def test_emu_chain_tables_and_unfused_reduction():
    """The fallbacks behind WHALE_TABLES_CHAIN=1 (per-slice recurrence on every branch) and WHALE_FUSED_REDUCE=0
    (K3 as separate kernels) are process-wide switches: exercise them in a child process."""
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from whale_jl_b200 import lib as wlib\nfrom conftest import run_parity\n"
            "run_parity(wlib.Lib(%r), 'c1_example1', sel=[0, 3], conds=['root'])\nprint('ok')\n"
            % (ROOT, os.path.join(ROOT, "tests"), EMU))
    env = dict(os.environ, WHALE_TABLES_CHAIN="1", WHALE_FUSED_REDUCE="0")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("order", ["reverse", "random"])
def test_emu_results_independent_of_thread_order(order):
    """The shim runs a CTA's threads one after another between two barriers, in index order by default;
    WHALE_EMU_SCHED changes that order.  A likelihood, gradient or sampled tree that moved with it would mean a
    missing barrier or a warp-lockstep assumption in a kernel.  (Process-wide switch: child process.)"""
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from whale_jl_b200 import lib as wlib\nfrom conftest import run_parity\nimport test_emu\n"
            "L = wlib.Lib(%r)\n"
            "run_parity(L, 'c1_example1', sel=[0, 3], conds=['root'])\n"
            "run_parity(L, 'const_wgdturing', sel=[1, 7], conds=['nonextinct'])\n"
            "run_parity(L, 'landplant100', sel=list(range(0, 100, 7)))\n"
            "test_emu._check_backtrack(L, 'const_wgdturing', [2])\nprint('ok')\n"
            % (ROOT, os.path.join(ROOT, "tests"), EMU))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, WHALE_EMU_SCHED=order),
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_emu_mle_lbfgs_like_the_reference_mle_test(L, tmp_path):
    """test/mle.jl's pattern (`optimize(f, g!, x0, LBFGS())` with f = −logpdf(model(rates), ccd) over (log λ, log μ),
    η and q fixed, gradient from AD) with the fused loglik+∇ call in place of ForwardDiff: the optimiser must converge
    to a stationary point, and the oracle must agree with the value there."""
    from scipy.optimize import minimize
    import whale_jl_b200 as W
    from whale_jl_b200 import synth
    from oracle import whale_oracle as wo, flat
    wlib.use(L)
    try:
        d = synth.generate(str(tmp_path / "mle"), 4, seed=31)
        tree = synth.c1_species_tree()
        w0 = W.WhaleModel(W.ConstantDLWGD(lam=0.5, mu=0.4, q=[0.2, 0.1], eta=0.66), tree, 0.05)
        ccd = W.read_ale(d, w0)
        evals = []

        def fg(x):
            lam, mu = np.exp(x)
            w = W.WhaleModel(W.ConstantDLWGD(lam=lam, mu=mu, q=[0.2, 0.1], eta=0.66), tree, 0.05)
            ll, g = W.logpdf_and_gradient(w, ccd)  # raw gradient: (λ, μ, q1, q2, η) on the natural scale
            evals.append(ll)
            return -ll, -np.array([g[0] * lam, g[1] * mu])

        f0, g0 = fg(np.log([0.5, 0.4]))
        res = minimize(fg, np.log([0.5, 0.4]), jac=True, method="L-BFGS-B", options={"maxiter": 25, "gtol": 1e-6})
        assert res.fun < f0 - 1e-3
        assert np.abs(res.jac).max() < 1e-3 * max(1.0, np.abs(g0).max())
        lam, mu = np.exp(res.x)
        ow = wo.WhaleModel(wo.ConstantDLWGD(lam=lam, mu=mu, q=[0.2, 0.1], eta=0.66), wo.c1_tree(), 0.05)
        ff = flat.FlatFams(wo.read_ale(d, ow), len(ow))
        assert -res.fun == pytest.approx(flat.logpdf(flat.FlatModel(ow), ff, grad=False)[0], rel=1e-9)
    finally:
        wlib.use(None)


def test_emu_arena_cache_round_trip(L, tmp_path):
    from conftest import arena_cache_round_trip
    wlib.use(L)
    try:
        arena_cache_round_trip(L, tmp_path)
    finally:
        wlib.use(None)


def test_emu_forward_tangent_mode(monkeypatch, tmp_path):
    """WHALE_GRAD_MODE=fwd keeps k_dp's forward tangents (the round-1 gradient path, still the fallback for families that
    do not fit the reverse kernel's working set): same numbers as the default reverse-mode gradient."""
    monkeypatch.setenv("WHALE_GRAD_MODE", "fwd")
    L2 = wlib.Lib(EMU)
    g = load_golden("c1_example1")
    mh = L2.model_create(golden_model(g))
    dh = L2.data_create(mh, golden_fams(g, [3]))
    assert L2.L.whale_data_grad_mode(dh) == 0
    L2.L.whale_data_destroy(dh)
    L2.L.whale_model_destroy(mh)
    run_parity(L2, "c1_example1", sel=[0, 3], conds=["root"])
    run_parity(L2, "const_wgdturing", sel=[1, 7], conds=["nonextinct"])
    run_parity(L2, "mul_tree", sel=[2], conds=["root"])
    from conftest import near_critical_vs_oracle
    wlib.use(L2)
    try:
        near_critical_vs_oracle(tmp_path, n_fam=2)
    finally:
        wlib.use(None)


def test_emu_reverse_mode_is_default_and_persistent(monkeypatch):
    """The default gradient is the reverse-mode kernel; with ONE persistent CTA (WHALE_REV_GRID=1) every family goes
    through the same CTA's loop (history slot, shared-memory carve-up and adjoint rows reused from family to family)."""
    monkeypatch.setenv("WHALE_REV_GRID", "1")
    monkeypatch.setenv("WHALE_GRAD_MODE", "rev")
    L2 = wlib.Lib(EMU)
    g = load_golden("c1_example1")
    mh = L2.model_create(golden_model(g))
    dh = L2.data_create(mh, golden_fams(g, [0, 1]))
    assert L2.L.whale_data_grad_mode(dh) == 1 and L2.L.whale_data_grad_passes(dh) == 1
    L2.L.whale_data_destroy(dh)
    L2.L.whale_model_destroy(mh)
    run_parity(L2, "c1_example1", sel=[5, 0, 9, 3], conds=["root"])
    run_parity(L2, "const_wgdturing", sel=[1, 7, 2], conds=["none"])


def test_emu_even_row_stride_variant(tmp_path):
    """The dense row layout of round 1 (WHALE_EVEN_STRIDE; the product build pads rows of even K to K+1 doubles per cell
    against shared-memory bank conflicts) must give the same numbers: forward-tangent gradient with 37 parameters,
    constant-rates WGD model, kept ℓ, backtracking."""
    L2 = wlib.Lib(os.path.join(ROOT, "tests", "emu", "libwhalecuda_emu_evenstride.so"))
    os.environ["WHALE_GRAD_MODE"] = "fwd"
    try:
        run_parity(L2, "c1_example1", sel=[0, 3], conds=["root"])
        run_parity(L2, "const_wgdturing", sel=[1, 7], conds=["nonextinct"])
    finally:
        del os.environ["WHALE_GRAD_MODE"]
    _check_backtrack(L2, "const_wgdturing", [2])
    g = load_golden("c1_example1")
    mh = L2.model_create(golden_model(g))
    dh = L2.data_create(mh, golden_fams(g, [3]))
    L2.logpdf_grad(mh, dh, g["xs"][2], g["m_pleaf"], 1, keep_ell=True)
    np.testing.assert_allclose(L2.ell_get(dh, 0), g["ell_3"], rtol=1e-11, atol=0)


def test_emu_backtrack_uses_kept_parameters(L):
    from conftest import backtrack_uses_kept_parameters
    backtrack_uses_kept_parameters(L)


def test_emu_multi_device_handle(L):
    from conftest import multi_device_vs_golden
    multi_device_vs_golden(L, [0, 0, 0])


def test_emu_peer_sum_two_ranks_in_one_process(L):
    """whale_peer_export / whale_peer_import / WHALE_PEER_SUM: two "ranks" (two data handles holding the two halves of the
    C1 families) exchange through each other's buffers; evaluated back to back each ends with the total of both — the
    full-batch golden value.  (In the emulation the exchange kernel does not wait for the peer's flag: rank 0's first
    total is incomplete by construction, so rank 0 is evaluated again after rank 1 has published.)"""
    g = load_golden("c1_example1")
    mh = L.model_create(golden_model(g))
    halves = [list(range(0, 6)), list(range(6, 12))]
    dhs = [L.data_create(mh, golden_fams(g, h)) for h in halves]
    hnd = [L.peer_export(dhs[r], r, 2) for r in range(2)]
    for r in range(2):
        L.peer_import(dhs[r], 1 - r, hnd[1 - r])
        assert L.L.whale_peer_ready(dhs[r]) == 1
    x = g["xs"][1]
    P = len(x)
    import ctypes as C
    f64p = C.POINTER(C.c_double)

    def ev(r, flags):
        ll = C.c_double()
        grad = np.zeros(P)
        xx = np.ascontiguousarray(x)
        pl = np.ascontiguousarray(g["m_pleaf"])
        L.check(L.L.whale_logpdf_grad(mh, dhs[r], xx.ctypes.data_as(f64p), pl.ctypes.data_as(f64p), 1, flags, C.byref(ll),
                                      grad.ctypes.data_as(f64p), None, None))
        return ll.value, grad

    ev(0, wlib.WANT_GRAD | wlib.PEER_SUM)              # step 1, rank 0 publishes (its own total is still partial)
    ll1, g1 = ev(1, wlib.WANT_GRAD | wlib.PEER_SUM)    # step 1, rank 1: sees both contributions
    assert ll1 == pytest.approx(g["tot_root"][1], rel=1e-9)
    np.testing.assert_allclose(g1, g["grad_root"][1], rtol=1e-9, atol=1e-9 * np.abs(g["grad_root"][1]).max())
    for dh in dhs:
        L.L.whale_data_destroy(dh)
    L.L.whale_model_destroy(mh)


def test_emu_packer_rejects_malformed_input(L):
    """whale_data_create validates what it is handed instead of reading out of bounds: empty batch, more than 65 535 clades
    (the reference's UInt16 clade ids, src/ccd.jl:13-33), clades not sorted by size, a compat list that is not ascending,
    a clade incompatible with the root, a triple pointing outside the family — each WHALE_ERR_ARG with a message."""
    g = load_golden("c1_example1")
    mh = L.model_create(golden_model(g))
    nn = len(g["m_order"])
    good = golden_fams(g, [3])

    def expect(fl, msg):
        with pytest.raises(wlib.WhaleCudaError, match=msg):
            L.data_create(mh, fl)

    try:
        expect(dict(good, n_fam=0), "n_fam must be positive")
        big = 70000
        expect(dict(n_fam=1, clade_off=np.array([0, big], np.int64), clade_nleaf=np.ones(big, np.int32),
                    split_off=np.zeros(big + 1, np.int64), g1=np.zeros(1, np.int32), g2=np.zeros(1, np.int32),
                    p=np.zeros(1), compat_off=np.zeros(nn + 1, np.int64), compat=np.zeros(1, np.int32)), "65535")
        bad = dict(good, clade_nleaf=good["clade_nleaf"].copy())
        bad["clade_nleaf"][0], bad["clade_nleaf"][-1] = bad["clade_nleaf"][-1], bad["clade_nleaf"][0]
        expect(bad, "sorted by size")
        bad = dict(good, compat=good["compat"].copy())
        # swap two entries of the first node's list that has at least two
        for e in range(nn):
            a, b = int(good["compat_off"][e]), int(good["compat_off"][e + 1])
            if b - a >= 2:
                bad["compat"][a], bad["compat"][a + 1] = bad["compat"][a + 1], bad["compat"][a]
                break
        expect(bad, "ascending")
        bad = dict(good, g1=good["g1"].copy())
        bad["g1"][0] = 10 ** 6
        expect(bad, "triple out of range")
    finally:
        L.L.whale_model_destroy(mh)


@pytest.mark.parametrize("mode", ["fwd", "rev"])
def test_emu_occupancy_line_unstages_outliers(mode, tmp_path):
    """A launch requests its bin's largest shared-memory need for every CTA, so the few families just above the size that
    still lets four of them share an SM read their lists in place (set_budgets / set_budgets_rev).  With the line lowered
    (WHALE_OCC_LINE) to a size that exactly two of twelve C2-shaped families exceed, the batch must become ONE bin at or
    below the line and the results must still match the oracle (staged and unstaged families in the same launch).
    Process-wide switches: child processes."""
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import pathlib\nfrom whale_jl_b200 import lib as wlib\nfrom conftest import synthetic_c2_shape_vs_oracle\n"
            "wlib.use(wlib.Lib(%r))\nsynthetic_c2_shape_vs_oracle(pathlib.Path(%r), n_fam=12)\nprint('ok')\n"
            % (ROOT, os.path.join(ROOT, "tests"), EMU, str(tmp_path)))

    def run(env):
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900,
                             env=dict(os.environ, WHALE_GRAD_MODE=mode, WHALE_DEBUG="2", WHALE_DEBUG_BINS="1", WHALE_CALIBRATE="0", **env))
        assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
        fam_key, bin_key = (" reverse: ", "plan 1 (reverse) bin") if mode == "rev" else (" plan 1: ", "plan 1 bin")
        lines = out.stderr.splitlines()
        needs = [int(ln.split("-> ")[1].split()[0]) for ln in lines if fam_key in ln and "-> " in ln][:12]  # first handle
        bins = [ln for ln in lines if bin_key in ln]
        return needs, bins

    needs, _ = run({})
    assert len(needs) == 12
    top = sorted(needs)
    line = (top[-3] + top[-2]) // 2  # exactly two families above
    _, bins2 = run({"WHALE_OCC_LINE": str(line)})  # (the per-family debug lines show the needs BEFORE the rule)
    assert line < max(needs)
    first = bins2[0]
    assert "count 12" in first and int(first.split(" smem ")[1].split()[0]) <= line, bins2


@pytest.mark.parametrize("env", [{"WHALE_PEER_FUSE": "0"}, {"WHALE_GRAD_MODE": "fwd"}, {"WHALE_GRAD_MODE": "fwd", "WHALE_PEER_FUSE": "0"}])
def test_emu_peer_sum_variants(env):
    """The exchange as its own launch (WHALE_PEER_FUSE=0) and behind the forward-tangent kernel: process-wide switches,
    so the two-rank test above runs again in a child process."""
    import sys
    out = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", __file__, "-k", "test_emu_peer_sum_two_ranks_in_one_process"],
                         env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "1 passed" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_emu_landplant_dt001(L, monkeypatch):
    """The tutorial's Δt = 0.01 discretisation (3 765 slices): the two smallest families of the fixture, both gradient modes."""
    g = load_golden("landplant_dt0.01")
    assert int(g["m_nslices"].sum()) == 3765
    size = np.diff(g["f_clade_off"])
    sel = [int(i) for i in np.argsort(size)[:2]]
    for mode in ("rev", "fwd"):
        monkeypatch.setenv("WHALE_GRAD_MODE", mode)
        run_parity(L, "landplant_dt0.01", sel=sel, conds=["root"])


@pytest.mark.parametrize("slots", ["16", "2", "1"])
def test_emu_track_sample_and_summary(L, monkeypatch, slots):
    """slots: posterior draws per batch of whale_track_sample (one walk launch per batch) — all draws in one batch, two
    batches with a short last one, and the draw-by-draw path."""
    from conftest import track_sample_and_summary
    monkeypatch.setenv("WHALE_TRACK_SLOTS", slots)
    track_sample_and_summary(L, n_samples=12, n_theta=3)
