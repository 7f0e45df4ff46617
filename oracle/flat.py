"""
ORACLE glue — TEST INFRASTRUCTURE ONLY (see oracle/whale_oracle.py header).

Flattens the reference-layout structures of ``whale_oracle`` (WhaleModel, CCD) into the plain-C
arrays ``oracle/whale_oracle.cpp`` consumes, and wraps ``liboracle.so`` with ctypes.
Build the library with ``make -C oracle`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import whale_oracle as wo

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)


class OModel(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("order", i32p), ("child0", i32p), ("child1", i32p),
                ("kind", i32p), ("nslices", i32p), ("dt", f64p), ("leafP", f64p), ("pleaf", f64p),
                ("lam_slot", i32p), ("mu_slot", i32p), ("q_slot", i32p), ("eta_slot", C.c_int32),
                ("log_scale", C.c_int32), ("condition", C.c_int32)]


class OFams(C.Structure):
    _fields_ = [("n_fam", C.c_int32), ("clade_off", i64p), ("clade_nleaf", i32p), ("split_off", i64p),
                ("g1", i32p), ("g2", i32p), ("p", f64p), ("compat_off", i64p), ("compat", i32p)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "whale_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oracle_logpdf.restype = C.c_double
        L.oracle_logpdf.argtypes = [C.POINTER(OModel), C.POINTER(OFams), f64p, C.c_int32, f64p, f64p, f64p,
                                    C.c_int32]
        L.oracle_slices.restype = None
        L.oracle_slices.argtypes = [C.POINTER(OModel), f64p, f64p, f64p, f64p]
        L.oracle_ell.restype = C.c_double
        L.oracle_ell.argtypes = [C.POINTER(OModel), C.POINTER(OFams), C.c_int32, f64p, f64p]
        L.oracle_backtrack.restype = C.c_int32
        L.oracle_backtrack.argtypes = [C.POINTER(OModel), C.POINTER(OFams), C.c_int32, f64p, f64p, C.c_int32,
                                       C.c_int32, i32p, i32p, i32p, i32p, i32p]
        L.oracle_max_threads.restype = C.c_int32
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


COND = {"none": 0, "root": 1, "nonextinct": 2}


class FlatModel:
    """0-based flattening of a whale_oracle.WhaleModel (node index = id-1)."""

    def __init__(self, wm: wo.WhaleModel):
        nn = len(wm)
        self.nn = nn
        self.order = np.array([n.id - 1 for n in wm.order], np.int32)
        self.child0 = np.full(nn, -1, np.int32)
        self.child1 = np.full(nn, -1, np.int32)
        self.kind = np.zeros(nn, np.int32)
        self.nslices = np.zeros(nn, np.int32)
        self.dt = np.zeros(nn)
        self.leafP = np.zeros(nn)
        self.pleaf = np.zeros(nn)
        self.lam_slot = np.full(nn, -1, np.int32)
        self.mu_slot = np.full(nn, -1, np.int32)
        self.q_slot = np.full(nn, -1, np.int32)
        r = wm.rates
        const = isinstance(r, wo.ConstantDLWGD)
        nq = len(r.q)
        nr = 1 if const else len(r.lam)
        for n in wm.order:
            e = n.id - 1
            if n.children:
                self.child0[e] = n.children[0].id - 1
                if len(n.children) > 1:
                    self.child1[e] = n.children[1].id - 1
            self.kind[e] = 2 if n.iswgd() else 3 if n.isroot() else 0 if n.isleaf() else 1
            self.nslices[e] = n.n
            self.dt[e] = n.dts[-1] if n.n > 0 else 0.0
            self.leafP[e] = n.leafP
            self.pleaf[e] = wo.getp(r, n)
            if const:
                self.lam_slot[e], self.mu_slot[e] = 0, 1
            else:
                c = wo.nonwgdchild(n)
                if c.id <= nr:
                    self.lam_slot[e], self.mu_slot[e] = c.id - 1, nr + c.id - 1
            if n.iswgd():
                self.q_slot[e] = 2 * nr + n.wgdid - 1
        self.P = 2 * nr + nq + 1
        self.x = np.array([float(v) for v in wo.vector_from_rates(r)])
        assert len(self.x) == self.P
        self.c = OModel(nn, _p(self.order, i32p), _p(self.child0, i32p), _p(self.child1, i32p),
                        _p(self.kind, i32p), _p(self.nslices, i32p), _p(self.dt, f64p), _p(self.leafP, f64p),
                        _p(self.pleaf, f64p), _p(self.lam_slot, i32p), _p(self.mu_slot, i32p),
                        _p(self.q_slot, i32p), 2 * nr + nq, 0 if const else 1, COND[wm.condition])
        self.row_off = np.concatenate([[0], np.cumsum(self.nslices + 1)]).astype(np.int64)


def model_from_arrays(g, condition="none"):
    """A FlatModel from the flattened arrays of a golden fixture (tests/golden/*.npz: m_* keys) — lets the live C++ oracle
    run where /root/reference is absent (the GPU box)."""
    fm = FlatModel.__new__(FlatModel)
    fm.nn = len(g["m_order"])
    for k, src, dt in (("order", "m_order", np.int32), ("child0", "m_child0", np.int32), ("child1", "m_child1", np.int32),
                       ("kind", "m_kind", np.int32), ("nslices", "m_nslices", np.int32), ("dt", "m_dt", np.float64),
                       ("leafP", "m_leafP", np.float64), ("pleaf", "m_pleaf", np.float64), ("lam_slot", "m_lam_slot", np.int32),
                       ("mu_slot", "m_mu_slot", np.int32), ("q_slot", "m_q_slot", np.int32)):
        setattr(fm, k, np.ascontiguousarray(g[src], dt))
    fm.P = int(g["m_P"])
    fm.x = np.ascontiguousarray(g["xs"][0], np.float64)
    fm.c = OModel(fm.nn, _p(fm.order, i32p), _p(fm.child0, i32p), _p(fm.child1, i32p), _p(fm.kind, i32p),
                  _p(fm.nslices, i32p), _p(fm.dt, f64p), _p(fm.leafP, f64p), _p(fm.pleaf, f64p), _p(fm.lam_slot, i32p),
                  _p(fm.mu_slot, i32p), _p(fm.q_slot, i32p), int(g["m_eta_slot"]), int(g["m_log_scale"]), COND[condition])
    fm.row_off = np.concatenate([[0], np.cumsum(fm.nslices + 1)]).astype(np.int64)
    return fm


def fams_from_arrays(g):
    """FlatFams from a golden fixture's f_* arrays."""
    ff = FlatFams.__new__(FlatFams)
    ff.nn = len(g["m_order"])
    ff.F = len(g["f_clade_off"]) - 1
    ff.clade_off = np.ascontiguousarray(g["f_clade_off"], np.int64)
    ff.nleaf = np.ascontiguousarray(g["f_nleaf"], np.int32)
    ff.split_off = np.ascontiguousarray(g["f_split_off"], np.int64)
    ff.g1 = np.ascontiguousarray(g["f_g1"], np.int32)
    ff.g2 = np.ascontiguousarray(g["f_g2"], np.int32)
    ff.p = np.ascontiguousarray(g["f_p"], np.float64)
    ff.compat_off = np.ascontiguousarray(g["f_compat_off"], np.int64)
    ff.compat = np.ascontiguousarray(g["f_compat"], np.int32)
    ff.c = OFams(ff.F, _p(ff.clade_off, i64p), _p(ff.nleaf, i32p), _p(ff.split_off, i64p), _p(ff.g1, i32p), _p(ff.g2, i32p),
                 _p(ff.p, f64p), _p(ff.compat_off, i64p), _p(ff.compat, i32p))
    return ff


class FlatFams:
    """Reference-layout CCD arena: clades sorted by size, triples in file order (0-based ids)."""

    def __init__(self, ccds, nn):
        self.F = len(ccds)
        clade_off = [0]
        nleaf, split_off, g1, g2, p, compat_off, compat = [], [0], [], [], [], [0], []
        for x in ccds:
            G = len(x)
            clade_off.append(clade_off[-1] + G)
            for c in x.clades[1:]:
                nleaf.append(len(c.leaves))
                for (a, b, pr) in c.splits:
                    g1.append(a - 1)
                    g2.append(b - 1)
                    p.append(pr)
                split_off.append(len(g1))
            for e in range(1, nn + 1):
                compat.extend(c - 1 for c in x.compat[e])
                compat_off.append(len(compat))
        self.clade_off = np.array(clade_off, np.int64)
        self.nleaf = np.array(nleaf, np.int32)
        self.split_off = np.array(split_off, np.int64)
        self.g1 = np.array(g1, np.int32)
        self.g2 = np.array(g2, np.int32)
        self.p = np.array(p, np.float64)
        self.compat_off = np.array(compat_off, np.int64)
        self.compat = np.array(compat, np.int32)
        self.nn = nn
        self.c = OFams(self.F, _p(self.clade_off, i64p), _p(self.nleaf, i32p), _p(self.split_off, i64p),
                       _p(self.g1, i32p), _p(self.g2, i32p), _p(self.p, f64p), _p(self.compat_off, i64p),
                       _p(self.compat, i32p))

    def ncompat(self, f, e):
        k = f * self.nn + e
        return int(self.compat_off[k + 1] - self.compat_off[k])


def logpdf(fm: FlatModel, ff: FlatFams, x=None, grad=False, per_family=False, nthreads=0):
    """oracle_logpdf: returns (total, ll_fam or None, grad or None, grad_fam or None)."""
    x = np.ascontiguousarray(fm.x if x is None else x, dtype=np.float64)
    ll = np.zeros(ff.F)
    g = np.zeros(fm.P) if grad else None
    gf = np.zeros((ff.F, fm.P)) if (grad and per_family) else None
    tot = lib().oracle_logpdf(C.byref(fm.c), C.byref(ff.c), _p(x, f64p), fm.P, _p(ll, f64p),
                              _p(g, f64p) if grad else None, _p(gf, f64p) if gf is not None else None, nthreads)
    return tot, ll, g, gf


def slices(fm: FlatModel, x=None):
    x = np.ascontiguousarray(fm.x if x is None else x, dtype=np.float64)
    n = int(fm.row_off[-1])
    eps, phi, psi = np.zeros(n), np.zeros(n), np.zeros(n)
    lib().oracle_slices(C.byref(fm.c), _p(x, f64p), _p(eps, f64p), _p(phi, f64p), _p(psi, f64p))
    return eps, phi, psi


def ell(fm: FlatModel, ff: FlatFams, f: int, x=None):
    """Full ℓ of family f: list over nodes (id order) of (n_e+1, C_e) arrays, and log L."""
    x = np.ascontiguousarray(fm.x if x is None else x, dtype=np.float64)
    sizes = [(int(fm.nslices[e]) + 1, ff.ncompat(f, e)) for e in range(fm.nn)]
    out = np.zeros(sum(r * c for r, c in sizes))
    l = lib().oracle_ell(C.byref(fm.c), C.byref(ff.c), f, _p(x, f64p), _p(out, f64p))
    mats, o = [], 0
    for r, c in sizes:
        mats.append(out[o:o + r * c].reshape(r, c))
        o += r * c
    return mats, l


def backtrack(fm: FlatModel, ff: FlatFams, f: int, u, x=None, max_nodes=1 << 16):
    """Returns (status/n_nodes, nodes[n,4] = (gamma, e, t, parent) 0-based ids (gamma -1 = loss), used)."""
    x = np.ascontiguousarray(fm.x if x is None else x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    g = np.zeros(max_nodes, np.int32)
    e = np.zeros(max_nodes, np.int32)
    t = np.zeros(max_nodes, np.int32)
    par = np.zeros(max_nodes, np.int32)
    used = C.c_int32(0)
    n = lib().oracle_backtrack(C.byref(fm.c), C.byref(ff.c), f, _p(x, f64p), _p(u, f64p), len(u), max_nodes,
                               _p(g, i32p), _p(e, i32p), _p(t, i32p), _p(par, i32p), C.byref(used))
    if n < 0:
        return n, None, used.value
    return n, np.stack([g[:n], e[:n], t[:n], par[:n]], axis=1), used.value
