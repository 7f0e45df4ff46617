"""Deterministic synthetic CCD generator for the benchmark configurations.

There is no ALEobserve binary (nor an MCMC tree sampler) in this environment, so each family is made as
follows — everything seeded, so CPU oracle and GPU see identical inputs:
  1. a gene tree is simulated on the species tree under the duplication–loss(+WGD) process (the process of
     src/simulation.jl:65-133), conditioned on both root clades being non-empty and a leaf-count window;
  2. a "posterior sample" of N unrooted topologies is emulated by random NNI perturbations of that tree;
  3. clades and clade splits are counted over the sample the way ALEobserve does for unrooted trees (every
     directed edge contributes the clade behind it and, if that clade is not a leaf, its two-way split);
  4. the result is written as a real `.ale` text file (so the genuine Julia reference can be run on the same
     data elsewhere) and/or parsed through the normal `read_ale` path.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import newick
from .model import WhaleModel, iswgd


# ---- 1. gene tree simulation (nested tuples; leaves are species names) ----
def _simulate_gene_tree(tree: newick.Node, lam, mu, q, rng):
    def join(a, b):
        if a is None:
            return b
        if b is None:
            return a
        return (a, b)

    def grow(node: newick.Node, t):
        while True:
            w = rng.exponential(1.0 / (lam + mu))
            if w > t:
                break
            t -= w
            if rng.random() < lam / (lam + mu):
                return join(grow(node, t), grow(node, t))
            return None
        if node.isleaf:
            return node.name
        if iswgd(node.name):
            c = node.children[0]
            digits = "".join(ch for ch in node.name if ch.isdigit())
            k = int(digits) - 1 if digits else 0
            qk = q[min(len(q) - 1, max(k, 0))] if len(q) else 0.0
            if rng.random() < qk:
                return join(grow(c, c.distance), grow(c, c.distance))
            return grow(c, c.distance)
        a, b = node.children
        return join(grow(a, a.distance), grow(b, b.distance))

    a, b = tree.children
    ga, gb = grow(a, a.distance), grow(b, b.distance)
    if ga is None or gb is None:
        return None
    return (ga, gb)


def _to_arrays(t):
    """nested tuples -> (left, right, parent, leaf species) arrays; leaves are nodes 0..n-1."""
    leaves = []

    def collect(x):
        if isinstance(x, str):
            leaves.append(x)
        else:
            collect(x[0])
            collect(x[1])

    collect(t)
    n = len(leaves)
    left = [-1] * (2 * n - 1)
    right = [-1] * (2 * n - 1)
    nxt_leaf = [0]
    nxt_int = [n]

    def build(x):
        if isinstance(x, str):
            i = nxt_leaf[0]
            nxt_leaf[0] += 1
            return i
        a = build(x[0])
        b = build(x[1])
        i = nxt_int[0]
        nxt_int[0] += 1
        left[i], right[i] = a, b
        return i

    root = build(t)
    return left, right, root, leaves


# ---- 2./3. NNI sample and ALEobserve-style counting ----
def _clade_counts(left, right, root, n, n_trees, nni_mean, rng):
    full = (1 << n) - 1
    bip: dict[int, int] = {}
    dip: dict[tuple, int] = {}
    nnodes = 2 * n - 1
    internal = [i for i in range(n, nnodes)]
    for _ in range(n_trees):
        L, R = list(left), list(right)
        par = [-1] * nnodes
        for v in internal:
            par[L[v]] = v
            par[R[v]] = v
        for _ in range(rng.poisson(nni_mean)):
            v = internal[rng.integers(len(internal))]
            if v == root:
                continue
            p = par[v]
            s = R[p] if L[p] == v else L[p]           # sibling of v
            c = L[v] if rng.random() < 0.5 else R[v]  # child of v swapped with the sibling
            if L[v] == c:
                L[v] = s
            else:
                R[v] = s
            if L[p] == s:
                L[p] = c
            else:
                R[p] = c
            par[s], par[c] = v, p
        # down masks (postorder via explicit stack)
        down = [0] * nnodes
        order = []
        st = [root]
        while st:
            v = st.pop()
            order.append(v)
            if v >= n:
                st.append(L[v])
                st.append(R[v])
        for v in reversed(order):
            down[v] = (1 << v) if v < n else down[L[v]] | down[R[v]]
        up = [0] * nnodes
        a, b = L[root], R[root]

        def count(mask, m1=None, m2=None):
            bip[mask] = bip.get(mask, 0) + 1
            if m1 is not None:
                key = (mask, min(m1, m2))
                dip[key] = dip.get(key, 0) + 1

        for v in order:
            if v == root:
                continue
            if v >= n:
                count(down[v], down[L[v]], down[R[v]])
            else:
                count(down[v])
            if par[v] == root:
                continue  # the root edge: U(a) = D(b), counted as a down clade
            p = par[v]
            s = R[p] if L[p] == v else L[p]
            if par[p] == root:
                o = b if p == a else a
                upp = down[o]
            else:
                upp = up[p]
            up[v] = upp | down[s]
            count(up[v], upp, down[s])
    return bip, dip, full


def _ale_sections(leaf_names, bip, dip, n_trees):
    """Assign ALE ids (leaf sets first, then by first appearance) and return the file sections."""
    n = len(leaf_names)
    names_sorted = sorted(range(n), key=lambda i: leaf_names[i])
    leaf_id = {g: k + 1 for k, g in enumerate(names_sorted)}  # gene index -> ALE leaf id (1-based)
    set_id: dict[int, int] = {}
    for g in range(n):
        set_id[1 << g] = len(set_id) + 1
    for mask in bip:
        if mask not in set_id:
            set_id[mask] = len(set_id) + 1
    return leaf_id, set_id


def write_ale(path, leaf_names, bip, dip, n_trees):
    leaf_id, set_id = _ale_sections(leaf_names, bip, dip, n_trees)
    n = len(leaf_names)
    out = ["#constructor_string", "synthetic;", "#observations", str(n_trees), "#Bip_counts"]
    for mask, sid in sorted(set_id.items(), key=lambda kv: kv[1]):
        if mask & (mask - 1):
            out.append(f"{sid}\t{bip[mask]}")
    out.append("#Bip_bls")
    for mask, sid in sorted(set_id.items(), key=lambda kv: kv[1]):
        out.append(f"{sid}\t{float(bip.get(mask, n_trees))}")
    out.append("#Dip_counts")
    for (mask, m1), c in dip.items():
        out.append(f"{set_id[mask]}\t{set_id[m1]}\t{set_id[mask ^ m1]}\t{c}")
    out += ["#last_leafset_id", str(len(set_id)), "#leaf-id"]
    for g in sorted(range(n), key=lambda i: leaf_names[i]):
        out.append(f"{leaf_names[g]}\t{leaf_id[g]}")
    out.append("#set-id")
    for mask, sid in sorted(set_id.items(), key=lambda kv: kv[1]):
        ids = sorted(leaf_id[g] for g in range(n) if mask >> g & 1)
        out.append(f"{sid}\t:\t" + "\t".join(map(str, ids)))
    out.append("#END")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")


def make_family(tree: newick.Node, rng, min_leaves=10, max_leaves=16, n_trees=100, nni_mean=6.0, q=(0.2, 0.1)):
    """One synthetic family: (leaf gene names, Bip counts, Dip counts, n_trees)."""
    while True:
        lam = float(np.exp(rng.normal(math.log(0.15), 0.5)))
        mu = float(np.exp(rng.normal(math.log(0.15), 0.5)))
        g = _simulate_gene_tree(tree, lam, mu, list(q), rng)
        if g is None:
            continue
        left, right, root, species = _to_arrays(g)
        if not (min_leaves <= len(species) <= max_leaves):
            continue
        break
    n = len(species)
    seen: dict[str, int] = {}
    names = []
    for s in species:
        seen[s] = seen.get(s, 0) + 1
        names.append(f"{s}_{seen[s]}")
    bip, dip, _ = _clade_counts(left, right, root, n, n_trees, nni_mean, rng)
    return names, bip, dip, n_trees


def c1_species_tree() -> newick.Node:
    """The reference's test tree: `Whale.extree` + wgd_1 above ATHA + wgd_2 above LCA(ATHA, ATRI)
    (test/runtests.jl:8-11) — a 9-taxon tree with 2 WGD nodes, used for the C2/C3/C5 shapes."""
    t = newick.extree()
    newick.insertnode(newick.getlca(t, "ATHA", "ATHA"), name="wgd_1")
    newick.insertnode(newick.getlca(t, "ATHA", "ATRI"), name="wgd_2")
    return t


def random_species_tree(n_taxa: int, n_wgd: int, seed: int, height: float = 5.0) -> newick.Node:
    """A random ultrametric species tree (coalescent-style joins, rescaled to `height` time units like
    `Whale.extree`) with `n_wgd` WGD nodes inserted halfway along random non-root branches — the C4 shape of
    BASELINE.json (30 taxa, 5 WGDs)."""
    rng = np.random.default_rng([seed, n_taxa, n_wgd])
    live = [(f"T{i:02d}", 0.0) for i in range(n_taxa)]  # (newick text, node height)
    t = 0.0
    while len(live) > 1:
        k = len(live)
        t += rng.exponential(1.0 / (k * (k - 1) / 2))
        i, j = sorted(rng.choice(k, 2, replace=False))
        (a, ha), (b, hb) = live[i], live[j]
        node = (f"({a}:{t - ha:.6f},{b}:{t - hb:.6f})", t)
        live = [x for m, x in enumerate(live) if m not in (i, j)] + [node]
    text = live[0][0] + ";"
    tree = newick.readnw(text)
    scale = height / t
    nodes = newick.postwalk(tree)
    for n in nodes:
        if not n.isroot:
            n.distance *= scale
    cand = [n for n in nodes if not n.isroot and n.distance > 0.3]
    pick = rng.choice(len(cand), size=min(n_wgd, len(cand)), replace=False)
    for w, i in enumerate(sorted(pick)):
        newick.insertnode(cand[i], name=f"wgd_{w + 1}")
    return tree


def c4_species_tree() -> newick.Node:
    """The C4 shape of BASELINE.json: 30 taxa, 5 WGD nodes (64 nodes)."""
    return random_species_tree(30, 5, seed=4)


C4_FAMILY = dict(min_leaves=60, max_leaves=100, n_trees=1000, nni_mean=12.0, q=(0.2, 0.1, 0.2, 0.1, 0.2))  # ~2,000 clades


def _gen_range(args):
    outdir, tree, seed, lo, hi, kw = args
    for f in range(lo, hi):
        rng = np.random.default_rng([seed, f])
        names, bip, dip, nt = make_family(tree, rng, **kw)
        write_ale(os.path.join(outdir, f"fam{f:06d}.ale"), names, bip, dip, nt)
    return hi - lo


def cache_dir(name: str) -> str:
    """Where the benchmarks keep generated families: outside the repository (WHALE_SYNTH_CACHE or the temp directory), so
    that a source snapshot never carries generated data."""
    import tempfile
    root = os.environ.get("WHALE_SYNTH_CACHE") or os.path.join(tempfile.gettempdir(), "whale_synth_cache")
    return os.path.join(root, name)


def generate(outdir: str, n_fam: int, seed: int, tree: newick.Node | None = None, workers: int = 0, first: int = 0,
             **kw) -> str:
    """Write the synthetic families first .. first + n_fam − 1 as .ale files into outdir (idempotent: reuses a complete
    directory).  Family f is drawn from its own generator seeded (seed, f), so the files depend neither on `workers`
    (0 = the host cores this process may use, for large sets) nor on how a set is split into shards (`first`)."""
    tree = tree or c1_species_tree()
    os.makedirs(outdir, exist_ok=True)
    done = outdir.rstrip("/") + ".complete"  # sibling marker: read_ale reads every file in outdir
    tag = f"{n_fam} {seed} {sorted(kw.items())}" + (f" first={first}" if first else "")
    if os.path.exists(done) and open(done).read().strip() == tag:
        return outdir
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    nw = workers or (min(ncpu, 64) if n_fam >= 2000 else 1)
    if nw > 1:
        import multiprocessing as mp
        step = max(1, (n_fam + 4 * nw - 1) // (4 * nw))
        jobs = [(outdir, tree, seed, lo, min(first + n_fam, lo + step), kw) for lo in range(first, first + n_fam, step)]
        with mp.get_context("fork").Pool(nw) as pool:
            pool.map(_gen_range, jobs)
    else:
        _gen_range((outdir, tree, seed, first, first + n_fam, kw))
    with open(done, "w") as fh:
        fh.write(tag + "\n")
    return outdir
