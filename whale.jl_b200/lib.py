"""ctypes binding of libwhalecuda (include/whalecuda.h).

The library is built in-tree (`whale.jl_b200/libwhalecuda.so`, see build.py) and is the ONLY compute path:
if it is missing, or no CUDA device is usable, every call raises — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwhalecuda.so")

i32p, i64p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)

WANT_GRAD, KEEP_ELL, PROFILE, PEER_SUM = 1, 2, 4, 8

# every symbol include/whalecuda.h declares (tests check the built library exports all of them)
SYMBOLS = ["whale_version", "whale_last_error", "whale_device_count", "whale_set_device", "whale_model_create",
           "whale_model_destroy", "whale_data_create", "whale_read_ale", "whale_data_destroy", "whale_data_nfam",
           "whale_data_arena_bytes", "whale_data_arena_dump", "whale_data_save", "whale_data_load", "whale_logpdf_grad", "whale_logpdf_grad_async", "whale_mixture_logpdf_grad",
           "whale_slices", "whale_ell_size", "whale_ell_get", "whale_backtrack", "whale_track", "whale_launch_count",
           "whale_work_estimate", "whale_last_kernel_ms", "whale_last_phase_cycles", "whale_last_node_cycles", "whale_last_family_cycles", "whale_last_tables_cycles", "whale_last_backtrack_ms", "whale_fp64_peak",
           "whale_data_grad_mode", "whale_data_grad_passes", "whale_set_devices", "whale_multi_create", "whale_multi_destroy",
           "whale_multi_ndev", "whale_multi_shard_size", "whale_multi_logpdf_grad", "whale_peer_export", "whale_peer_import",
           "whale_peer_ready", "whale_peer_sum_async", "whale_backtrack_device", "whale_track_sample", "whale_trees_counts", "whale_trees_view",
           "whale_trees_get", "whale_trees_summary"]


class ModelDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("order", i32p), ("child0", i32p), ("child1", i32p), ("kind", i32p),
                ("n_slices", i32p), ("slice_dt", f64p), ("leafP", f64p), ("n_params", C.c_int32),
                ("lam_slot", i32p), ("mu_slot", i32p), ("q_slot", i32p), ("eta_slot", C.c_int32),
                ("log_scale", C.c_int32)]


class CCDDesc(C.Structure):
    _fields_ = [("n_fam", C.c_int32), ("clade_off", i64p), ("clade_nleaf", i32p), ("split_off", i64p),
                ("g1", i32p), ("g2", i32p), ("p", f64p), ("compat_off", i64p), ("compat", i32p)]


class WhaleCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libwhalecuda error {code}: {msg}")
        self.code = code


def _ptr(a, t):
    return a.ctypes.data_as(t)


class Lib:
    """A loaded libwhalecuda with typed entry points."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build the CUDA library first (python -m whale_jl_b200.build or "
                f"__graft_entry__.build()); there is no CPU fallback")
        self.path = path
        self._ctx = {}  # (model handle, data handle) -> reusable argument buffers of logpdf_grad
        _LIBS[id(self)] = self
        L = self.L = C.CDLL(path)
        vp = C.c_void_p
        L.whale_version.restype = C.c_int32
        L.whale_last_error.argtypes = [C.c_char_p, C.c_size_t]
        L.whale_device_count.restype = C.c_int32
        L.whale_set_device.argtypes = [C.c_int32]
        L.whale_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(vp)]
        L.whale_model_destroy.argtypes = [vp]
        L.whale_data_create.argtypes = [vp, C.POINTER(CCDDesc), C.POINTER(vp)]
        L.whale_data_destroy.argtypes = [vp]
        L.whale_data_nfam.argtypes = [vp]
        L.whale_data_grad_mode.argtypes = [vp]
        L.whale_data_grad_passes.argtypes = [vp]
        L.whale_data_arena_bytes.argtypes = [vp]
        L.whale_data_arena_bytes.restype = C.c_int64
        L.whale_data_arena_dump.argtypes = [vp, vp, C.c_int64]
        L.whale_data_arena_dump.restype = C.c_int64
        L.whale_data_save.argtypes = [vp, C.c_char_p]
        L.whale_data_load.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
        L.whale_logpdf_grad.argtypes = [vp, vp, f64p, f64p, C.c_int32, C.c_uint32, f64p, f64p, f64p, f64p]
        L.whale_logpdf_grad_async.argtypes = [vp, vp, vp, C.c_int32, C.c_uint32, vp, vp]
        L.whale_slices.argtypes = [vp, f64p, f64p, f64p, f64p, f64p]
        L.whale_ell_size.argtypes = [vp, C.c_int32]
        L.whale_ell_size.restype = C.c_int64
        L.whale_ell_get.argtypes = [vp, C.c_int32, f64p]
        L.whale_track.argtypes = [vp, vp, C.c_int32, f64p, f64p, C.c_int32, f64p, C.c_int64, C.c_int32, i32p, i32p, i32p,
                                  i32p, i32p, i32p, f64p]
        L.whale_backtrack.argtypes = [vp, vp, C.c_int32, f64p, C.c_int64, C.c_int32, i32p, i32p, i32p, i32p, i32p,
                                      i32p]
        L.whale_launch_count.restype = C.c_int64
        L.whale_work_estimate.argtypes = [vp, vp, C.c_uint32, f64p, f64p]
        L.whale_fp64_peak.argtypes = [f64p]
        L.whale_last_kernel_ms.argtypes = [vp, f64p, f64p, f64p]
        L.whale_read_ale.argtypes = [vp, C.c_int32, C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_char_p), i32p, i64p, i32p,
                                     C.c_int32, C.POINTER(vp), i32p]
        L.whale_mixture_logpdf_grad.argtypes = [vp, vp, C.c_int32, f64p, f64p, f64p, C.c_int32, C.c_uint32,
                                                C.POINTER(C.c_double), f64p, f64p]
        L.whale_last_phase_cycles.argtypes = [vp, f64p, f64p]
        L.whale_last_tables_cycles.argtypes = [vp, C.c_int32, f64p]
        L.whale_last_node_cycles.argtypes = [vp, f64p, f64p, C.c_int32]
        L.whale_last_family_cycles.argtypes = [vp, f64p]
        L.whale_last_backtrack_ms.argtypes = [vp, f64p]
        u64p = C.POINTER(C.c_uint64)
        L.whale_backtrack_device.argtypes = [vp, vp, C.c_int32, f64p, C.c_int64, C.c_uint64, C.c_int32, i64p]
        L.whale_track_sample.argtypes = [vp, vp, C.c_int32, f64p, f64p, C.c_int32, i32p, f64p, C.c_int64, C.c_uint64,
                                         C.c_int32, i64p]
        L.whale_trees_counts.argtypes = [vp, i32p, i32p]
        L.whale_trees_view.argtypes = [vp, C.POINTER(i64p), C.POINTER(i32p), i64p]
        L.whale_trees_get.argtypes = [vp, i64p, i32p]
        L.whale_trees_summary.argtypes = [vp, i32p, u64p, i32p, i32p, u64p]
        L.whale_peer_export.argtypes = [vp, C.c_int32, C.c_int32, vp]
        L.whale_peer_import.argtypes = [vp, C.c_int32, vp]
        L.whale_peer_ready.argtypes = [vp]
        L.whale_peer_sum_async.argtypes = [vp, vp, vp]
        L.whale_set_devices.argtypes = [C.c_int32, i32p]
        L.whale_multi_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(CCDDesc), C.POINTER(vp)]
        L.whale_multi_destroy.argtypes = [vp]
        L.whale_multi_ndev.argtypes = [vp]
        L.whale_multi_shard_size.argtypes = [vp, C.c_int32]
        L.whale_multi_logpdf_grad.argtypes = [vp, f64p, f64p, C.c_int32, C.c_uint32, f64p, f64p, f64p]

    def check(self, rc):
        if rc != 0:
            buf = C.create_string_buffer(512)
            self.L.whale_last_error(buf, 512)
            raise WhaleCudaError(rc, buf.value.decode(errors="replace"))

    # ---- handles ----
    def model_create(self, m) -> int:
        keep = [np.ascontiguousarray(a) for a in
                (m.order, m.child0, m.child1, m.kind, m.n_slices, m.slice_dt, m.leafP, m.lam_slot, m.mu_slot,
                 m.q_slot)]
        d = ModelDesc(m.nn, _ptr(keep[0], i32p), _ptr(keep[1], i32p), _ptr(keep[2], i32p), _ptr(keep[3], i32p),
                      _ptr(keep[4], i32p), _ptr(keep[5], f64p), _ptr(keep[6], f64p), m.n_params,
                      _ptr(keep[7], i32p), _ptr(keep[8], i32p), _ptr(keep[9], i32p), m.eta_slot, m.log_scale)
        h = C.c_void_p()
        self.check(self.L.whale_model_create(C.byref(d), C.byref(h)))
        return h.value

    def _model_desc(self, m):
        keep = [np.ascontiguousarray(a) for a in
                (m.order, m.child0, m.child1, m.kind, m.n_slices, m.slice_dt, m.leafP, m.lam_slot, m.mu_slot,
                 m.q_slot)]
        d = ModelDesc(m.nn, _ptr(keep[0], i32p), _ptr(keep[1], i32p), _ptr(keep[2], i32p), _ptr(keep[3], i32p),
                      _ptr(keep[4], i32p), _ptr(keep[5], f64p), _ptr(keep[6], f64p), m.n_params,
                      _ptr(keep[7], i32p), _ptr(keep[8], i32p), _ptr(keep[9], i32p), m.eta_slot, m.log_scale)
        return d, keep

    def multi_create(self, m, flat: dict, devices) -> int:
        """whale_set_devices + whale_multi_create: one handle over several GPUs (families sharded by predicted work)."""
        ids = np.ascontiguousarray(devices, np.int32)
        self.check(self.L.whale_set_devices(len(ids), _ptr(ids, i32p)))
        md, keep = self._model_desc(m)
        cd = CCDDesc(flat["n_fam"], _ptr(flat["clade_off"], i64p), _ptr(flat["clade_nleaf"], i32p),
                     _ptr(flat["split_off"], i64p), _ptr(flat["g1"], i32p), _ptr(flat["g2"], i32p),
                     _ptr(flat["p"], f64p), _ptr(flat["compat_off"], i64p), _ptr(flat["compat"], i32p))
        h = C.c_void_p()
        self.check(self.L.whale_multi_create(C.byref(md), C.byref(cd), C.byref(h)))
        return h.value

    def multi_logpdf_grad(self, h, x, p_leaf, condition, n_fam, want_grad=False, per_family=False):
        x = np.ascontiguousarray(x, np.float64)
        pl = np.ascontiguousarray(p_leaf, np.float64)
        ll = C.c_double()
        g = np.zeros(len(x)) if want_grad else None
        lf = np.zeros(n_fam) if per_family else None
        self.check(self.L.whale_multi_logpdf_grad(h, _ptr(x, f64p), _ptr(pl, f64p), condition, WANT_GRAD if want_grad else 0,
                                                  C.byref(ll), _ptr(g, f64p) if want_grad else None,
                                                  _ptr(lf, f64p) if per_family else None))
        return ll.value, g, lf

    def peer_export(self, dh, rank, world) -> bytes:
        buf = C.create_string_buffer(64)
        self.check(self.L.whale_peer_export(dh, rank, world, buf))
        return buf.raw

    def peer_import(self, dh, peer, handle: bytes):
        buf = C.create_string_buffer(handle, 64)
        self.check(self.L.whale_peer_import(dh, peer, buf))

    def data_create(self, mh: int, flat: dict) -> int:
        d = CCDDesc(flat["n_fam"], _ptr(flat["clade_off"], i64p), _ptr(flat["clade_nleaf"], i32p),
                    _ptr(flat["split_off"], i64p), _ptr(flat["g1"], i32p), _ptr(flat["g2"], i32p),
                    _ptr(flat["p"], f64p), _ptr(flat["compat_off"], i64p), _ptr(flat["compat"], i32p))
        h = C.c_void_p()
        self.check(self.L.whale_data_create(mh, C.byref(d), C.byref(h)))
        return h.value

    def data_save(self, dh: int, path: str):
        """whale_data_save: the packed state of a data handle as a binary arena cache."""
        self.check(self.L.whale_data_save(dh, path.encode()))

    def data_load(self, mh: int, path: str) -> int:
        """whale_data_load: a data handle from an arena cache written for the same species tree and slicing."""
        h = C.c_void_p()
        self.check(self.L.whale_data_load(mh, path.encode(), C.byref(h)))
        return h.value

    def read_ale(self, mh, model, files, n_threads=0):
        """whale_read_ale: parse + build + pack natively; returns (data handle, clades per family)."""
        names = list(model.spmap)
        ids = np.array([model.spmap[n] for n in names], np.int32)
        off = np.zeros(model.nn + 1, np.int64)
        np.cumsum([len(model.clade[e]) for e in range(model.nn)], out=off[1:])
        cl = np.array([s for e in range(model.nn) for s in sorted(model.clade[e])], np.int32)
        cpaths = (C.c_char_p * len(files))(*[f.encode() for f in files])
        cnames = (C.c_char_p * len(names))(*[n.encode() for n in names])
        ncl = np.zeros(len(files), np.int32)
        h = C.c_void_p()
        self.check(self.L.whale_read_ale(mh, len(files), cpaths, len(names), cnames, _ptr(ids, i32p), _ptr(off, i64p),
                                         _ptr(cl, i32p), n_threads, C.byref(h), _ptr(ncl, i32p)))
        return h.value, ncl

    def logpdf_grad(self, mh, dh, x, p_leaf, condition, want_grad=False, keep_ell=False, per_family=False,
                    per_family_grad=False, profile=False, peer_sum=False):
        """whale_logpdf_grad with host buffers.  peer_sum: one process per GPU — the returned (loglik, grad) is the sum over
        all ranks (whale_peer_export / whale_peer_import done before; every rank must make the same calls)."""
        flags = ((WANT_GRAD if want_grad else 0) | (KEEP_ELL if keep_ell else 0) | (PROFILE if profile else 0) |
                 (PEER_SUM if peer_sum else 0))
        if not (per_family or per_family_grad):
            # the call an optimiser / sampler makes every iteration: argument buffers are kept per handle pair (building
            # ctypes pointers from numpy arrays costs ~3 µs each, three of them per call — 3 % of a 0.3 ms evaluation)
            P, nn = len(x), len(p_leaf)
            ctx = self._ctx.get((mh, dh))
            if ctx is None or ctx[0] != P or ctx[1] != nn:
                xb, plb, gb, llv = (C.c_double * max(P, 1))(), (C.c_double * max(nn, 1))(), (C.c_double * max(P, 1))(), C.c_double()
                ctx = (P, nn, xb, np.frombuffer(xb, np.float64, P), plb, np.frombuffer(plb, np.float64, nn), gb,
                       np.frombuffer(gb, np.float64, P), llv, C.byref(llv))
                if len(self._ctx) > 64:
                    self._ctx.clear()
                self._ctx[(mh, dh)] = ctx
            ctx[3][:] = x
            ctx[5][:] = p_leaf
            rc = self.L.whale_logpdf_grad(mh, dh, ctx[2], ctx[4], condition, flags, ctx[9], ctx[6] if want_grad else None, None, None)
            if rc:
                self.check(rc)
            return ctx[8].value, (ctx[7].copy() if want_grad else None), None, None
        x = np.ascontiguousarray(x, np.float64)
        pl = np.ascontiguousarray(p_leaf, np.float64)
        F = self.L.whale_data_nfam(dh)
        P = len(x)
        ll = C.c_double()
        g = np.zeros(P) if want_grad else None
        lf = np.zeros(F) if per_family else None
        gf = np.zeros((F, P)) if per_family_grad else None
        self.check(self.L.whale_logpdf_grad(mh, dh, _ptr(x, f64p), _ptr(pl, f64p), condition, flags, C.byref(ll),
                                            _ptr(g, f64p) if want_grad else None,
                                            _ptr(lf, f64p) if per_family else None,
                                            _ptr(gf, f64p) if per_family_grad else None))
        return ll.value, g, lf, gf

    def mixture_logpdf_grad(self, mh, dh, xs, log_w, p_leaf, condition, want_grad=False):
        xs = np.ascontiguousarray(xs, np.float64)
        lw = np.ascontiguousarray(log_w, np.float64)
        pl = np.ascontiguousarray(p_leaf, np.float64)
        J, P = xs.shape
        ll = C.c_double()
        gx = np.zeros((J, P)) if want_grad else None
        gw = np.zeros(J) if want_grad else None
        self.check(self.L.whale_mixture_logpdf_grad(mh, dh, J, _ptr(xs, f64p), _ptr(lw, f64p), _ptr(pl, f64p), condition,
                                                    WANT_GRAD if want_grad else 0, C.byref(ll),
                                                    _ptr(gx, f64p) if want_grad else None,
                                                    _ptr(gw, f64p) if want_grad else None))
        return ll.value, gx, gw

    def slices(self, mh, x, p_leaf, nrows):
        x = np.ascontiguousarray(x, np.float64)
        pl = np.ascontiguousarray(p_leaf, np.float64)
        eps, phi, psi = np.zeros(nrows), np.zeros(nrows), np.zeros(nrows)
        self.check(self.L.whale_slices(mh, _ptr(x, f64p), _ptr(pl, f64p), _ptr(eps, f64p), _ptr(phi, f64p),
                                       _ptr(psi, f64p)))
        return eps, phi, psi

    def ell_get(self, dh, fam):
        n = self.L.whale_ell_size(dh, fam)
        if n < 0:
            raise IndexError(fam)
        out = np.zeros(n)
        self.check(self.L.whale_ell_get(dh, fam, _ptr(out, f64p)))
        return out

    def arena_dump(self, dh) -> np.ndarray:
        n = self.L.whale_data_arena_dump(dh, None, 0)
        buf = np.zeros(n, np.uint8)
        got = self.L.whale_data_arena_dump(dh, buf.ctypes.data_as(C.c_void_p), n)
        if got != n:
            raise WhaleCudaError(2, "arena dump failed")
        return buf

    def backtrack_device(self, mh, dh, n_samples, uniforms=None, seed=0, max_nodes=512) -> int:
        """whale_backtrack_device: walks from the kept ℓ, results stay on the device; returns the total node count."""
        tot = C.c_int64()
        if uniforms is not None:
            U = np.ascontiguousarray(uniforms, np.float64)
            self.check(self.L.whale_backtrack_device(mh, dh, n_samples, _ptr(U, f64p), U.shape[-1], 0, max_nodes, C.byref(tot)))
        else:
            self.check(self.L.whale_backtrack_device(mh, dh, n_samples, None, 0, seed, max_nodes, C.byref(tot)))
        return tot.value

    def track_sample(self, mh, dh, xs, p_leaf, n_samples, theta_index=None, uniforms=None, seed=0, max_nodes=512) -> int:
        """whale_track_sample: per-(family, sample) posterior rows; results stay on the device; returns the total node count."""
        X = np.ascontiguousarray(xs, np.float64)
        pl = np.ascontiguousarray(p_leaf, np.float64)
        ti = None if theta_index is None else np.ascontiguousarray(theta_index, np.int32)
        tot = C.c_int64()
        if uniforms is not None:
            U = np.ascontiguousarray(uniforms, np.float64)
            up, stride = _ptr(U, f64p), U.shape[-1]
        else:
            up, stride = None, 0
        self.check(self.L.whale_track_sample(mh, dh, X.shape[0], _ptr(X, f64p), _ptr(pl, f64p), n_samples,
                                             None if ti is None else _ptr(ti, i32p), up, stride, seed, max_nodes, C.byref(tot)))
        return tot.value

    def trees_counts(self, dh, W):
        cnt, st = np.zeros(W, np.int32), np.zeros(W, np.int32)
        self.check(self.L.whale_trees_counts(dh, _ptr(cnt, i32p), _ptr(st, i32p)))
        return cnt, st

    def trees_get(self, dh, W, total):
        """Compact trees: (offsets[W+1], nodes[total, 4]) copied out of the library's pinned buffer."""
        off = np.zeros(W + 1, np.int64)
        nodes = np.zeros((max(total, 1), 4), np.int32)
        self.check(self.L.whale_trees_get(dh, _ptr(off, i64p), _ptr(nodes, i32p)))
        return off, nodes[:total]

    def trees_view(self, dh, W):
        """Zero-copy views of the library-owned pinned buffers (valid until the next backtracking call)."""
        po, pn, tot = i64p(), i32p(), C.c_int64()
        self.check(self.L.whale_trees_view(dh, C.byref(po), C.byref(pn), C.byref(tot)))
        off = np.ctypeslib.as_array(po, shape=(W + 1,))
        nodes = np.ctypeslib.as_array(pn, shape=(max(tot.value, 1), 4))[:tot.value]
        return off, nodes

    def trees_summary(self, dh, F, S, with_tree_hash=False):
        nd = np.zeros(F, np.int32)
        h, c, f1 = np.zeros(F * S, np.uint64), np.zeros(F * S, np.int32), np.zeros(F * S, np.int32)
        th = np.zeros(F * S, np.uint64) if with_tree_hash else None
        u64p = C.POINTER(C.c_uint64)
        self.check(self.L.whale_trees_summary(dh, _ptr(nd, i32p), _ptr(h, u64p), _ptr(c, i32p), _ptr(f1, i32p),
                                              _ptr(th, u64p) if with_tree_hash else None))
        return nd, h.reshape(F, S), c.reshape(F, S), f1.reshape(F, S), (th.reshape(F, S) if with_tree_hash else None)

    def backtrack(self, mh, dh, n_samples, uniforms, max_nodes=512):
        """whale_backtrack: uniforms [F, n_samples, stride]; returns (counts[F,S], status[F,S], nodes[F,S,max_nodes,4])
        with node columns (gamma, e, t, parent)."""
        F = self.L.whale_data_nfam(dh)
        u = np.ascontiguousarray(uniforms, np.float64).reshape(F, n_samples, -1)
        stride = u.shape[2]
        W = F * n_samples
        cnt, st = np.zeros(W, np.int32), np.zeros(W, np.int32)
        g, e, t, p = (np.zeros(W * max_nodes, np.int32) for _ in range(4))
        self.check(self.L.whale_backtrack(mh, dh, n_samples, _ptr(u, f64p), stride, max_nodes, _ptr(cnt, i32p),
                                          _ptr(g, i32p), _ptr(e, i32p), _ptr(t, i32p), _ptr(p, i32p), _ptr(st, i32p)))
        nodes = np.stack([g, e, t, p], axis=1).reshape(F, n_samples, max_nodes, 4)
        return cnt.reshape(F, n_samples), st.reshape(F, n_samples), nodes

    def track(self, mh, dh, xs, p_leaf, condition, uniforms, max_nodes=512):
        """whale_track: xs [n_theta, P], uniforms [F, n_theta, stride]; returns (counts, status, nodes, loglik[n_theta])
        laid out like `backtrack` with n_samples = n_theta."""
        F = self.L.whale_data_nfam(dh)
        xs = np.ascontiguousarray(xs, np.float64)
        S = xs.shape[0]
        pl = np.ascontiguousarray(p_leaf, np.float64)
        u = np.ascontiguousarray(uniforms, np.float64).reshape(F, S, -1)
        stride = u.shape[2]
        W = F * S
        cnt, st = np.zeros(W, np.int32), np.zeros(W, np.int32)
        g, e, t, p = (np.zeros(W * max_nodes, np.int32) for _ in range(4))
        ll = np.zeros(S)
        self.check(self.L.whale_track(mh, dh, S, _ptr(xs, f64p), _ptr(pl, f64p), condition, _ptr(u, f64p), stride, max_nodes,
                                      _ptr(cnt, i32p), _ptr(g, i32p), _ptr(e, i32p), _ptr(t, i32p), _ptr(p, i32p),
                                      _ptr(st, i32p), _ptr(ll, f64p)))
        nodes = np.stack([g, e, t, p], axis=1).reshape(F, S, max_nodes, 4)
        return cnt.reshape(F, S), st.reshape(F, S), nodes, ll

    def work_estimate(self, mh, dh, want_grad=True):
        fl, by = C.c_double(), C.c_double()
        self.check(self.L.whale_work_estimate(mh, dh, WANT_GRAD if want_grad else 0, C.byref(fl), C.byref(by)))
        return fl.value, by.value

    def last_kernel_ms(self, dh):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.check(self.L.whale_last_kernel_ms(dh, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def last_phase_cycles(self, dh):
        mean, mx = np.zeros(8), np.zeros(8)
        self.check(self.L.whale_last_phase_cycles(dh, _ptr(mean, f64p), _ptr(mx, f64p)))
        if self.L.whale_data_grad_mode(dh) == 1:  # reverse mode (see whalecuda.h)
            names = ["prologue", "leaf_phase", "fwd_stage_row1", "bwd_row1", "slices_fwd_bwd", "root_fwd_bwd", "total",
                     "contraction"]
        else:
            names = ["prologue", "leaf_phase", "staging", "row1", "slices", "root", "total"]
        return {n: (float(mean[i]), float(mx[i])) for i, n in enumerate(names)}

    def last_family_cycles(self, dh):
        out = np.zeros((self.L.whale_data_nfam(dh), 8))
        self.check(self.L.whale_last_family_cycles(dh, _ptr(out, f64p)))
        return out

    def last_node_cycles(self, dh):
        a, b = np.zeros(32), np.zeros(32)
        n = self.L.whale_last_node_cycles(dh, _ptr(a, f64p), _ptr(b, f64p), 32)
        if n < 0:
            self.check(n)
        return a[:n].tolist(), b[:n].tolist()

    def last_backtrack_ms(self, dh) -> float:
        ms = C.c_double()
        self.check(self.L.whale_last_backtrack_ms(dh, C.byref(ms)))
        return ms.value

    def last_tables_cycles(self, mh, with_grad=True):
        out = np.zeros(32)
        self.check(self.L.whale_last_tables_cycles(mh, 1 if with_grad else 0, _ptr(out, f64p)))
        n = int(out[31])
        return {"meta": out[0], "levels": out[1:1 + n].tolist(), "total": out[30]}

    def logpdf_grad_async(self, mh, dh, d_x_ptr, condition, flags, d_out_ptr, stream_ptr):
        """Device-resident evaluation enqueued on `stream_ptr` (a cudaStream_t as int); not synchronised."""
        self.check(self.L.whale_logpdf_grad_async(mh, dh, d_x_ptr, condition, flags, d_out_ptr, stream_ptr))

    def fp64_peak(self) -> float:
        t = C.c_double()
        self.check(self.L.whale_fp64_peak(C.byref(t)))
        return t.value


_default: Lib | None = None


_LIBS: dict = {}  # id(Lib) -> Lib, for finalisers that only know the id


def lib_by_id(i: int):
    return _LIBS.get(i)


def get() -> Lib:
    """The process-wide library instance (loaded on first use)."""
    global _default
    if _default is None:
        _default = Lib()
    return _default


def use(lib: Lib | None):
    """Install another library instance (tests use this to run the host logic against the emulation build)."""
    global _default
    _default = lib
