"""CCD data: `.ale` ingest and the reference's clade layout (src/ccd.jl).

`read_ale` parses ALEobserve output (src/ccd.jl:147-185), adds leaf clades and the ubiquitous clade
(:199-248), renumbers clades by (size, old id) (:108-110) and computes, per species-tree node, the
ascending list of compatible clades (:39,59-75).  The result is stored directly in the flattened form that
crosses the C ABI (`whale_ccd_desc`); `whale_data_create` then builds the device arena once.
"""
from __future__ import annotations

import os

import numpy as np

from .model import WhaleModel


class CCD:
    """One gene family.  Clades are 0-based here (Julia id − 1), sorted by size, leaves first, the
    ubiquitous clade last; `g1/g2/p[split_off[c]:split_off[c+1]]` are clade c's triples in file order;
    `compat[e]` is the ascending array of clades whose species set ⊆ clade(e)."""

    def __init__(self, fname, total, leaves, nleaf, species, split_off, g1, g2, p, counts, model: WhaleModel):
        self.fname = fname
        self.total = total
        self.leaves = leaves          # gene names, indexed by leaf-clade id
        self.nleaf = nleaf            # int32[Γ] clade sizes
        self.species = species        # list[frozenset] species ids per clade
        self.split_off, self.g1, self.g2, self.p = split_off, g1, g2, p
        self.counts = counts
        self.compat = compat_lists(species, model)
        self._batch = None

    def __len__(self):
        return len(self.nleaf)

    def index(self, model_nn=None) -> np.ndarray:
        """The reference's `index[γ,e]` matrix (src/ccd.jl:59-75), 0-based, −1 = incompatible."""
        nn = len(self.compat)
        idx = np.full((len(self), nn), -1, np.int32)
        for e, comp in enumerate(self.compat):
            idx[comp, e] = np.arange(len(comp), dtype=np.int32)
        return idx

    def __repr__(self):
        return f"CCD(Γ={len(self)}, 𝓛={len(self.leaves)})"


def compat_lists(species, model: WhaleModel):
    out = []
    for e in range(model.nn):
        cl = model.clade[e]
        out.append(np.array([c for c, s in enumerate(species) if s <= cl], np.int32))
    return out


def _num(tok: str):
    try:
        return int(tok)
    except ValueError:
        return float(tok)


def parse_ale(path: str) -> dict:
    """Sections of an ALEobserve file as plain dicts keyed by the file's own (1-based) ids."""
    with open(path) as fh:
        text = fh.read()
    parts = text.split("#")
    if len(parts) != 10:
        raise ValueError(f"Not a valid .ale file {path}")
    sec = {}
    for part in parts[1:-1]:
        lines = [l for l in part.replace(":\t", "").split("\n") if l != ""]
        sec[lines[0].strip().replace("-", "_")] = lines[1:]
    obs = _num(sec["observations"][0])
    bip = {int(l.split()[0]): _num(l.split()[1]) for l in sec["Bip_counts"]}
    dip: dict[int, list] = {}
    for l in sec["Dip_counts"]:
        a, b, c, n = l.split()
        dip.setdefault(int(a), []).append((int(b), int(c), _num(n)))
    leaf_name = {int(l.split()[1]): l.split()[0] for l in sec["leaf_id"]}  # leaf id -> gene name
    sets = {int(l.split()[0]): [int(t) for t in l.split()[1:]] for l in sec["set_id"]}
    return dict(observations=obs, bip=bip, dip=dip, leaf_name=leaf_name, sets=sets)


def build_ccd(path: str, model: WhaleModel) -> CCD:
    a = parse_ale(path)
    obs, bip, dip, sets = a["observations"], a["bip"], a["dip"], a["sets"]
    # addleafclades! (src/ccd.jl:199-216): leaf ids -> their set ids; leaf clades get count = observations
    leaf2set, leafname = {}, {}
    for k in sorted(sets):
        v = sets[k]
        if len(v) == 1:
            bip[k] = obs
            dip[k] = []
            leafname[k] = a["leaf_name"][v[0]]
            leaf2set[v[0]] = k
        else:
            sets[k] = [leaf2set[i] for i in v]
    for lid, k in leaf2set.items():
        sets[k] = [k]
    # addubiquitous! (src/ccd.jl:219-248): complementary pairs (i<j) split the root clade
    nleaves = len(leafname)
    ns = len(sets)
    G = ns + 1
    fs = {k: frozenset(v) for k, v in sets.items()}
    bysize: dict[int, list] = {}
    for k in range(1, ns + 1):
        bysize.setdefault(len(fs[k]), []).append(k)
    rootsplits, N = [], 0
    for i in range(1, ns + 1):
        for j in bysize.get(nleaves - len(fs[i]), []):
            if j > i and not (fs[i] & fs[j]):
                if bip[i] != bip[j]:
                    raise ValueError(f"{path}: complementary clades {i},{j} have different counts")
                N += bip[i]
                rootsplits.append((i, j, bip[j]))
    dip[G] = rootsplits
    bip[G] = N
    sets[G] = sorted(fs[rootsplits[-1][0]] | fs[rootsplits[-1][1]])
    # CCD ctor (src/ccd.jl:102-121): new ids by (size, old id); p = count / count(parent clade)
    order = sorted(sets, key=lambda k: (len(sets[k]), k))
    newid = {k: i for i, k in enumerate(order)}
    nleaf = np.array([len(sets[k]) for k in order], np.int32)
    species = [frozenset(model.spmap[leafname[g].split("_")[0]] for g in sets[k]) for k in order]
    split_off = [0]
    g1, g2, p, counts = [], [], [], []
    for k in order:
        for (x, y, c) in dip.get(k, []):
            g1.append(newid[x])
            g2.append(newid[y])
            p.append(c / bip[k])
        split_off.append(len(g1))
        counts.append(bip[k])
    leaves = [leafname[k] for k in sorted(leafname)]
    return CCD(os.path.basename(path), obs, leaves, nleaf, species, np.array(split_off, np.int64),
               np.array(g1, np.int32), np.array(g2, np.int32), np.array(p, np.float64),
               np.array(counts, np.int64), model)


class CCDVector(list):
    """`Vector{CCD}` with the lazily-built device arena attached (one `whale_data` handle per model
    structure)."""

    def __init__(self, items=()):
        super().__init__(items)
        self._data = {}

    def __getitem__(self, i):
        """A slice is again a batch (`ccds[:10]` keeps its own device arena across calls instead of being re-packed on
        every evaluation)."""
        r = super().__getitem__(i)
        return CCDVector(r) if isinstance(i, slice) else r

    def close(self):
        """Release the device arenas of this batch (`whale_data_destroy`); they are rebuilt on the next evaluation."""
        _release_data(self._data)

    def __del__(self):
        try:
            _release_data(self._data)
        except Exception:
            pass

    def flatten(self, nn: int):
        """Concatenate the families into the `whale_ccd_desc` arrays (include/whalecuda.h)."""
        clade_off = np.zeros(len(self) + 1, np.int64)
        np.cumsum([len(c) for c in self], out=clade_off[1:])
        nleaf = np.concatenate([c.nleaf for c in self]).astype(np.int32)
        soffs, base = [np.zeros(1, np.int64)], 0
        for c in self:
            soffs.append(c.split_off[1:] + base)
            base += int(c.split_off[-1])
        split_off = np.concatenate(soffs).astype(np.int64)
        g1 = np.concatenate([c.g1 for c in self]).astype(np.int32)
        g2 = np.concatenate([c.g2 for c in self]).astype(np.int32)
        p = np.concatenate([c.p for c in self]).astype(np.float64)
        lens = np.array([len(comp) for c in self for comp in c.compat], np.int64)
        compat_off = np.zeros(len(lens) + 1, np.int64)
        np.cumsum(lens, out=compat_off[1:])
        compat = np.concatenate([comp for c in self for comp in c.compat]).astype(np.int32)
        assert len(lens) == len(self) * nn
        return dict(n_fam=len(self), clade_off=clade_off, clade_nleaf=nleaf, split_off=split_off, g1=g1, g2=g2,
                    p=p, compat_off=compat_off, compat=compat)


def _release_data(data: dict):
    """Destroy the data handles a batch holds (keys: (id(lib), model handle))."""
    from . import lib as _lib
    for (lid, _), h in list(data.items()):
        L = _lib.lib_by_id(lid)
        if L is not None and h:
            L.L.whale_data_destroy(h)
    data.clear()


def _ale_files(path: str) -> list[str]:
    if not os.path.exists(path):
        raise FileNotFoundError(f"Not a file nor directory `{path}`")
    if os.path.isfile(path) and path.endswith(".ale"):
        files = [path]
    elif os.path.isfile(path):
        files = [l.strip() for l in open(path) if l.strip()]
    else:
        files = [os.path.join(path, f) for f in sorted(os.listdir(path))]
    return [f for f in files if not f.startswith("#")]


class NativeCCDVector:
    """`Vector{CCD}` whose families were parsed, built and packed by the library itself (`whale_read_ale`): the
    CCDs exist only as the device arena.  Usable wherever a batch is evaluated (`logpdf`, `logpdf_and_gradient`,
    `logpdf_mixture`, `backtrack`, `track`); per-CCD host detail (`ell`, leaf names) needs `read_ale`."""

    def __init__(self, files, n_clades, model, lib, handle):
        self.files, self.n_clades = list(files), np.asarray(n_clades)
        self._model_nn = model.nn
        self._data = {(id(lib), model._handle[1]): handle}

    def __len__(self):
        return len(self.files)

    def close(self):
        _release_data(self._data)

    def __del__(self):
        try:
            _release_data(self._data)
        except Exception:
            pass


def read_ale_native(path: str, model: WhaleModel, n_threads: int = 0) -> NativeCCDVector:
    """`read_ale(path, wm)` (src/ccd.jl:126-137) done natively: files parsed on all host threads straight into the
    packed device arena (the reference parses under `tmap`, src/ccd.jl:134)."""
    from . import lib as _lib
    from .core import _model_handle
    files = _ale_files(path)
    L = _lib.get()
    mh = _model_handle(model)
    h, ncl = L.read_ale(mh, model, files, n_threads)
    return NativeCCDVector(files, ncl, model, L, h)


def save_arena(xs, model: WhaleModel, path: str) -> None:
    """Write the packed device arena of a batch (`read_ale` / `read_ale_native` result) as a binary cache
    (`whale_data_save`), so a later process can skip parsing and packing."""
    from . import lib as _lib
    from .core import _data_handle, _as_vector
    xs, _ = _as_vector(xs)
    _, dh = _data_handle(model, xs)
    _lib.get().data_save(dh, path)


def load_arena(path: str, model: WhaleModel) -> NativeCCDVector:
    """A batch straight from an arena cache (`whale_data_load`) for a model with the same species tree and slicing
    (a different one is refused).  Like `read_ale_native`, the CCDs then exist only as the device arena."""
    from . import lib as _lib
    from .core import _model_handle
    L = _lib.get()
    mh = _model_handle(model)
    h = L.data_load(mh, path)
    n = int(L.L.whale_data_nfam(h))
    return NativeCCDVector([f"{path}#{i}" for i in range(n)], np.zeros(n, np.int32), model, L, h)


def read_ale(path: str, model: WhaleModel) -> CCDVector:
    """`read_ale(path, wm)` (src/ccd.jl:126-137): a `.ale` file, a directory of them (sorted like
    `readdir`), or a text file listing paths (lines starting with # skipped)."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"Not a file nor directory `{path}`")
    if os.path.isfile(path) and path.endswith(".ale"):
        files = [path]
    elif os.path.isfile(path):
        files = [l.strip() for l in open(path) if l.strip()]
    else:
        files = [os.path.join(path, f) for f in sorted(os.listdir(path))]
    files = [f for f in files if not f.startswith("#")]
    return CCDVector(build_ccd(f, model) for f in files)
