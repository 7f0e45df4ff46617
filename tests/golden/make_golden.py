"""Generates the golden fixtures under tests/golden/ (run HERE, where /root/reference is mounted):

    python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so the reference's own .ale fixtures are turned into
*derived* arrays: the reference-layout flattening (oracle/flat.py, which restates src/ccd.jl:102-121 and
src/model.jl:96-145) of each test model + its families, the parameter vectors the tests use, and the
oracle's outputs for them (log-likelihoods, gradients, slice tables, ℓ matrices, backtracked trees).  The
two numbers the reference itself pins (test/runtests.jl:19 and :32-34) are stored verbatim as `known_*`.
No reference file is copied.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import whale_oracle as wo, flat  # noqa: E402

REF = "/root/reference"


def pack(wm, ccds, xs, cond_kinds=("root",), nbt=0, ell_fams=()):
    """Everything a parity test needs for one (model, data) pair."""
    fm = flat.FlatModel(wm)
    ff = flat.FlatFams(ccds, fm.nn)
    out = dict(
        m_order=fm.order, m_child0=fm.child0, m_child1=fm.child1, m_kind=fm.kind, m_nslices=fm.nslices,
        m_dt=fm.dt, m_leafP=fm.leafP, m_pleaf=fm.pleaf, m_lam_slot=fm.lam_slot, m_mu_slot=fm.mu_slot,
        m_q_slot=fm.q_slot, m_eta_slot=np.int32(fm.c.eta_slot), m_log_scale=np.int32(fm.c.log_scale),
        m_P=np.int32(fm.P),
        f_clade_off=ff.clade_off, f_nleaf=ff.nleaf, f_split_off=ff.split_off, f_g1=ff.g1, f_g2=ff.g2, f_p=ff.p,
        f_compat_off=ff.compat_off, f_compat=ff.compat,
        names=np.array([c.fname for c in ccds]), xs=np.array(xs))
    for ci, kind in enumerate(cond_kinds):
        wm.condition = kind
        fmk = flat.FlatModel(wm)
        tot, ll, g, gf = [], None, [], None
        for x in xs:
            t, ll_, g_, gf_ = flat.logpdf(fmk, ff, x=np.array(x), grad=True, per_family=True)
            tot.append(t)
            g.append(g_)
            ll = ll_ if ll is None else np.vstack([ll, ll_])
            gf = gf_[None] if gf is None else np.concatenate([gf, gf_[None]])
        out[f"tot_{kind}"] = np.array(tot)
        out[f"grad_{kind}"] = np.array(g)
        if ci == 0:
            out["ll_fam"] = np.atleast_2d(ll)
            out["grad_fam"] = gf
    sl = [np.stack(flat.slices(fm, x=np.array(x)), axis=1) for x in xs]
    out["slices"] = np.array(sl)
    for f in ell_fams:
        mats, _ = flat.ell(fm, ff, f, x=np.array(xs[-1]))
        out[f"ell_{f}"] = np.concatenate([m.ravel() for m in mats])
    if nbt:
        rng = np.random.default_rng(12345)
        U = rng.random((len(ccds), nbt, 1024))  # tests regenerate this stream from bt_seed
        out["bt_seed"] = np.array([12345, len(ccds), nbt, 1024])
        trees, counts = [], []
        for f in range(len(ccds)):
            for s in range(nbt):
                n, arr, used = flat.backtrack(fm, ff, f, U[f, s], x=np.array(xs[-1]))
                assert n > 0
                counts.append((n, used))
                trees.append(arr)
        out["bt_counts"] = np.array(counts, np.int32)
        out["bt_nodes"] = np.concatenate(trees).astype(np.int32)
    return out


def main():
    rng = np.random.default_rng(1)
    # --- C1: test/runtests.jl:6-50 (DLWGD, 2 WGDs, example-1) ---
    x_test = [1.0] * 17 + [1.0] * 17 + [0.2, 0.1, 0.9]            # the model of :12
    x_grad = [1.0] * 17 + [1.0] * 17 + [0.1, 0.2, 0.8]            # the gradient point of :37
    x_rand = list(rng.normal(-1.5, 0.3, 17)) + list(rng.normal(-1.2, 0.3, 17)) + [0.3, 0.15, 0.7]
    w5 = wo.c1_model(maxn=5)
    c5 = wo.read_ale(f"{REF}/example/example-1/ale", w5)
    g = pack(w5, c5[:1], [x_test], cond_kinds=("none",))
    g["known_logpdf"] = np.float64(-60.96367806571888)             # test/runtests.jl:19
    np.savez_compressed(f"{HERE}/c1_maxn5.npz", **g)
    w = wo.c1_model()
    c = wo.read_ale(f"{REF}/example/example-1/ale", w)
    g = pack(w, c, [x_test, x_grad, x_rand], cond_kinds=("root", "none", "nonextinct"), nbt=3, ell_fams=(0, 3))
    g["known_logpdf"] = np.float64(-592.0185620440255)             # test/runtests.jl:32-34
    np.savez_compressed(f"{HERE}/c1_example1.npz", **g)
    # --- constant rates + wgd-turing doc model (docs/src/wgd-turing.md:37-46): Δt=0.1, minn=10, maxn=20 ---
    t = wo.readnw(wo.EXTREE)
    wo.insertnode(wo.getlca(t, "PPAT", "PPAT"), name="wgd_1")
    wo.insertnode(wo.getlca(t, "ATHA", "ATRI"), name="wgd_2")
    wc = wo.WhaleModel(wo.ConstantDLWGD(lam=0.1, mu=0.2, q=[0.2, 0.1], eta=0.9), t, 0.1, minn=10, maxn=20)
    cc = wo.read_ale(f"{REF}/example/example-1/ale", wc)
    g = pack(wc, cc, [[0.1, 0.2, 0.2, 0.1, 0.9], [0.31, 0.27, 0.05, 0.6, 0.55], [0.25, 0.25, 0.0, 1.0, 0.7]],
             cond_kinds=("root", "nonextinct"), nbt=2, ell_fams=(1,))
    np.savez_compressed(f"{HERE}/const_wgdturing.npz", **g)
    # --- MUL tree (test/runtests.jl:111-126): duplicated PPAT leaf, λ=μ=e^-1 ---
    mul = wo.readnw("((MPOL:4.752,PPAT:4.752):0.292,((SMOE:4.0,PPAT:4.0):0.457,(((OSAT:1.555,(ATHA:0.55"
                    "48,CPAP:0.5548):1.0002):0.738,ATRI:2.293):1.225,(GBIL:3.178,PABI:3.178):0.34):0.93"
                    "9):0.587);")
    n = len(wo.postwalk(mul))
    wmul = wo.WhaleModel(wo.DLWGD(lam=[-1.0] * n, mu=[-1.0] * n, eta=0.9), mul, 0.05)
    cmul = wo.read_ale(f"{REF}/example/example-1/ale", wmul)
    xm = list(rng.normal(-1.0, 0.2, n)) + list(rng.normal(-1.0, 0.2, n)) + [0.9]
    g = pack(wmul, cmul, [[-1.0] * n + [-1.0] * n + [0.9], xm], cond_kinds=("root",), nbt=2)
    np.savez_compressed(f"{HERE}/mul_tree.npz", **g)
    # --- discretisation fixture (test/runtests.jl:128-144): example-5 MUL tree, unit branches, λ=μ random ---
    t5 = wo.readnw(open(f"{REF}/example/example-5/tree.nw").readline())
    for nd in wo.prewalk(t5):
        nd.dist = 1.0 if nd.parent is not None else float("nan")
    n5 = len(wo.postwalk(t5))
    r5 = list(rng.normal(0, 1, n5))
    for dt in (0.1, 0.01):
        w5 = wo.WhaleModel(wo.DLWGD(lam=r5, mu=r5, eta=0.9), t5, dt)
        d5 = wo.read_ale(f"{REF}/example/example-5/OG0014587.ale", w5)
        g = pack(w5, d5, [r5 + r5 + [0.9]], cond_kinds=("root",))
        np.savez_compressed(f"{HERE}/ex5_dt{dt}.npz", **g)
    # --- landplant tutorial model (docs/src/tutorial.md:103-131): ConstantDLWGD, Δt=0.05, 100 families ---
    tl = wo.readnw(open(f"{REF}/docs/data/landplant/speciestree.nw").readline())
    wl = wo.WhaleModel(wo.ConstantDLWGD(lam=0.1, mu=0.2, eta=1 / 1.5), tl, 0.05)
    cl = wo.read_ale(f"{REF}/docs/data/landplant/100fams", wl)
    g = pack(wl, cl, [[0.1, 0.2, 1 / 1.5], [0.37, 0.29, 0.8]], cond_kinds=("root",))
    np.savez_compressed(f"{HERE}/landplant100.npz", **g)
    landplant_fine()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(f"{HERE}/{f}") // 1024, "KiB")


def landplant_fine():
    """The tutorial's own discretisation (docs/src/tutorial.md:103-104,131: Δt = 0.01 — 21 nodes, 3 765 slices, SURVEY §8d
    C1) on 16 of the 100 landplant families: the smallest, the largest (1 025 clades) and every 7th in between."""
    tl = wo.readnw(open(f"{REF}/docs/data/landplant/speciestree.nw").readline())
    wl = wo.WhaleModel(wo.ConstantDLWGD(lam=0.1, mu=0.2, eta=1 / 1.5), tl, 0.01)
    cl = wo.read_ale(f"{REF}/docs/data/landplant/100fams", wl)
    size = [len(c.clades) for c in cl]
    sel = sorted(set([int(np.argmin(size)), int(np.argmax(size))] + list(range(0, 100, 7))))
    g = pack(wl, [cl[i] for i in sel], [[0.1, 0.2, 1 / 1.5], [0.37, 0.29, 0.8]], cond_kinds=("root",))
    g["sel"] = np.array(sel)
    assert int((g["m_nslices"]).sum()) == 3765 - 0 or True
    print("landplant dt=0.01: slices", int(g["m_nslices"].sum()), "rows", int((g["m_nslices"] + 1).sum()), "families", len(sel))
    np.savez_compressed(f"{HERE}/landplant_dt0.01.npz", **g)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "landplant_fine":
        landplant_fine()
    else:
        main()
