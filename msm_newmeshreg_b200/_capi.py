"""ctypes binding of the C ABI in include/msm_b200.h (libmsm_b200.so, built in-tree for sm_100a).

There is no CPU fallback: if the shared library is missing this module raises ImportError-like RuntimeError, and
`Context()` raises when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsm_b200.so")

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)


class MsmParams(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("rexp", C.c_double), ("mu", C.c_double), ("kappa", C.c_double), ("k_exp", C.c_double),
                ("range", C.c_double), ("simmeasure", C.c_int), ("rmode", C.c_int), ("dweight", C.c_int), ("anorm", C.c_int),
                ("meanangle", C.c_double), ("mvdmax", C.c_double), ("group", C.c_int), ("bary_orthogonal", C.c_int)]


EXPORTS = [
    "msm_default_params", "msm_create", "msm_destroy", "msm_last_error", "msm_set_stream", "msm_synchronize", "msm_set_params",
    "msm_set_target", "msm_set_source", "msm_reset_source", "msm_set_cpgrid", "msm_set_abs_weights", "msm_reset_cpgrid", "msm_set_shard", "msm_set_labels",
    "msm_set_pairs", "msm_set_triplets", "msm_set_group", "msm_build_patches", "msm_build_patches_async", "msm_patches_resolve", "msm_patch_count", "msm_get_patches",
    "msm_unary_table", "msm_unary_table_dev", "msm_triplet_costs", "msm_triplet_table", "msm_triplet_table_dev", "msm_triplet_moves",
    "msm_triplet_moves_f32", "msm_triplet_moves_dev", "msm_triplet_moves_range_dev", "msm_unary_table_peers", "msm_triplet_moves_range_peers", "msm_pair_table", "msm_pair_table_dev", "msm_locate", "msm_total_cost", "msm_launch_count", "msm_patch_build_stats",
    "msm_set_timing", "msm_last_kernel_ms", "msm_timing_collect", "msm_set_group_patches", "msm_group_pair_costs",
    "msm_group_pair_moves", "msm_group_pair_table", "msm_mesh_interpolate",
    "msm_smooth_data", "msm_histogram_match", "msm_variance_normalise", "msm_create_exclusion",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). There is no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.msm_last_error.restype = C.c_char_p
        _lib.msm_last_error.argtypes = [C.c_void_p]
        _lib.msm_launch_count.restype = C.c_longlong
        _lib.msm_launch_count.argtypes = [C.c_void_p]
        _lib.msm_last_kernel_ms.restype = C.c_double
        _lib.msm_last_kernel_ms.argtypes = [C.c_void_p, C.c_char_p]
        _lib.msm_destroy.argtypes = [C.c_void_p]
        _lib.msm_destroy.restype = None
    return _lib


class MsmError(RuntimeError):
    """Raised for every non-zero status of the C ABI (mirrors newmeshreg::MeshregException, src/reg_tools.h:15-23)."""


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


class Context:
    """One GPU context (one per process/GPU).  Thin, typed veneer over the C ABI; all arrays are numpy host arrays
    except the *_dev variants, which take raw device pointers (e.g. torch tensors' data_ptr())."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        st = lib().msm_create(int(device), C.byref(self._h))
        if st != 0:
            raise MsmError(lib().msm_last_error(None).decode())
        self.N = self.L = self.T = self.P = 0
        self.k_begin = self.k_end = 0

    @classmethod
    def from_handle(cls, handle, N, L, T=0, P=0):
        """Wrap a context owned by someone else (e.g. a host-side cost function object); close() will not destroy it."""
        self = cls.__new__(cls)
        self._h = C.c_void_p(handle)
        self._borrowed = True
        self.N, self.L, self.T, self.P = N, L, T, P
        self.k_begin, self.k_end = 0, N
        return self

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            if not getattr(self, "_borrowed", False):
                lib().msm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise MsmError(lib().msm_last_error(self._h).decode())

    # ---- configuration -------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._ck(lib().msm_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._ck(lib().msm_synchronize(self._h))

    def set_params(self, lam=1.0, rexp=2.0, mu=0.4, kappa=1.6, k_exp=2.0, rng=1.0, simmeasure=2, rmode=3, dweight=False,
                   anorm=False, meanangle=0.0, mvdmax=0.0, group=False, bary_orthogonal=False):
        p = MsmParams(lam, rexp, mu, kappa, k_exp, rng, simmeasure, rmode, int(dweight), int(anorm), meanangle, mvdmax, int(group),
                      int(bary_orthogonal))
        self._ck(lib().msm_set_params(self._h, C.byref(p)))

    def set_target(self, xyz, tris, feat=None):
        xyz, px = _d(xyz); tris, pt = _i(tris)
        if feat is None:
            self._ck(lib().msm_set_target(self._h, px, len(xyz), pt, len(tris), None, 0))
        else:
            feat = np.atleast_2d(feat); feat, pf = _d(feat)
            assert feat.shape[1] == len(xyz)
            self._ck(lib().msm_set_target(self._h, px, len(xyz), pt, len(tris), pf, feat.shape[0]))
        self.Vt = len(xyz)

    def set_source(self, xyz, feat, cfweight=None, multivariate=None):
        xyz, px = _d(xyz); feat = np.atleast_2d(feat); feat, pf = _d(feat)
        assert feat.shape[1] == len(xyz)
        if multivariate is None:
            multivariate = feat.shape[0] > 1
        if cfweight is None:
            pw, wr = None, 0
        else:
            cfweight = np.atleast_2d(cfweight); cfweight, pw = _d(cfweight); wr = cfweight.shape[0]
        self._ck(lib().msm_set_source(self._h, px, len(xyz), pf, feat.shape[0], pw, wr, int(multivariate)))
        self.Vs = len(xyz)

    def reset_source(self, xyz):
        xyz, px = _d(xyz); assert len(xyz) == self.Vs
        self._ck(lib().msm_reset_source(self._h, px))

    def set_cpgrid(self, xyz, tris, maxsep, orig_xyz=None, abs_weights=None):
        xyz, px = _d(xyz); tris, pt = _i(tris); maxsep, pm = _d(maxsep)
        po = pa = None
        if orig_xyz is not None:
            orig_xyz, po = _d(orig_xyz)
        if abs_weights is not None:
            abs_weights, pa = _d(abs_weights)
        self._ck(lib().msm_set_cpgrid(self._h, px, len(xyz), pt, len(tris), pm, po, pa))
        self.N = len(xyz); self.k_begin, self.k_end = 0, self.N

    def set_abs_weights(self, aw):
        aw, pa = _d(aw); assert len(aw) == self.N
        self._ck(lib().msm_set_abs_weights(self._h, pa))

    def reset_cpgrid(self, xyz):
        xyz, px = _d(xyz); assert len(xyz) == self.N
        self._ck(lib().msm_reset_cpgrid(self._h, px))

    def set_shard(self, k_begin, k_end):
        self._ck(lib().msm_set_shard(self._h, int(k_begin), int(k_end)))
        self.k_begin, self.k_end = int(k_begin), int(k_end)

    def set_labels(self, labels, centre):
        labels, pl = _d(labels); centre, pc = _d(centre)
        self._ck(lib().msm_set_labels(self._h, pl, len(labels), pc))
        self.L = len(labels)

    def set_pairs(self, pairs):
        pairs, pp = _i(pairs)
        self._ck(lib().msm_set_pairs(self._h, pp, len(pairs)))
        self.P = len(pairs)

    def set_triplets(self, triplets):
        triplets, pt = _i(triplets)
        self._ck(lib().msm_set_triplets(self._h, pt, len(triplets)))
        self.T = len(triplets)

    def set_group(self, S, n_subj, t_subj):
        self._ck(lib().msm_set_group(self._h, int(S), int(n_subj), int(t_subj)))

    # ---- hot path ------------------------------------------------------------------------------------------
    def build_patches(self):
        self._ck(lib().msm_build_patches(self._h))

    def build_patches_async(self):
        """get_source_data without a host wait; follow device-resident launches with patches_resolve()."""
        self._ck(lib().msm_build_patches_async(self._h))

    def patches_resolve(self):
        """Reads the status of a deferred build; True = the patches were rebuilt (exact path): repeat the device-resident launches."""
        r = C.c_int(0)
        self._ck(lib().msm_patches_resolve(self._h, C.byref(r)))
        return bool(r.value)

    def patches(self):
        tot = C.c_longlong()
        self._ck(lib().msm_patch_count(self._h, C.byref(tot)))
        off = np.zeros(self.N + 1, np.int32); idx = np.zeros(max(tot.value, 1), np.int32)
        self._ck(lib().msm_get_patches(self._h, off.ctypes.data_as(c_ip), idx.ctypes.data_as(c_ip)))
        return off, idx[:tot.value]

    def patch_build_stats(self):
        """(one-pass builds, exact two-pass builds) of this context so far."""
        a, b = C.c_longlong(), C.c_longlong()
        self._ck(lib().msm_patch_build_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def unary_table(self, out=None):
        if out is None:
            out = np.zeros((self.L, self.N), np.float64)
        self._ck(lib().msm_unary_table(self._h, out.ctypes.data_as(c_dp)))
        return out

    def unary_table_dev(self, dev_ptr):
        self._ck(lib().msm_unary_table_dev(self._h, C.c_void_p(dev_ptr)))

    def triplet_costs(self, q):
        q, pq = _i(q); out = np.empty(len(q), np.float64)
        self._ck(lib().msm_triplet_costs(self._h, len(q), pq, out.ctypes.data_as(c_dp)))
        return out

    def triplet_table(self, t0=0, t1=None):
        t1 = self.T if t1 is None else t1
        out = np.empty((t1 - t0, self.L, self.L, self.L), np.float32)
        self._ck(lib().msm_triplet_table(self._h, t0, t1, out.ctypes.data_as(c_fp)))
        return out

    def triplet_table_dev(self, t0, t1, dev_ptr):
        self._ck(lib().msm_triplet_table_dev(self._h, t0, t1, C.c_void_p(dev_ptr)))

    def triplet_moves(self, labeling, alpha):
        labeling, pl = _i(labeling)
        shape = (self.L, self.T, 8) if alpha < 0 else (self.T, 8)
        out = np.empty(shape, np.float64)
        self._ck(lib().msm_triplet_moves(self._h, pl, int(alpha), out.ctypes.data_as(c_dp)))
        return out

    def triplet_moves_dev(self, labeling_dev_ptr, alpha, out_dev_ptr):
        self._ck(lib().msm_triplet_moves_dev(self._h, C.c_void_p(labeling_dev_ptr), int(alpha), C.c_void_p(out_dev_ptr)))

    def triplet_moves_range_dev(self, labeling_dev_ptr, alpha, t0, t1, out_dev_ptr):
        self._ck(lib().msm_triplet_moves_range_dev(self._h, C.c_void_p(labeling_dev_ptr), int(alpha), int(t0), int(t1), C.c_void_p(out_dev_ptr)))

    # fused exchange (control-point / triplet sharding): peer_ptrs = device pointers to every rank's FULL table
    @staticmethod
    def _peer_array(peer_ptrs):
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        return arr, len(peer_ptrs)

    def unary_table_peers(self, peer_ptrs, multicast_ptr=0):
        arr, n = self._peer_array(peer_ptrs)
        self._ck(lib().msm_unary_table_peers(self._h, arr, n, C.c_void_p(int(multicast_ptr) or None)))

    def triplet_moves_range_peers(self, labeling_dev_ptr, alpha, t0, t1, peer_ptrs, multicast_ptr=0):
        arr, n = self._peer_array(peer_ptrs)
        self._ck(lib().msm_triplet_moves_range_peers(self._h, C.c_void_p(labeling_dev_ptr), int(alpha), int(t0), int(t1), arr, n,
                                                     C.c_void_p(int(multicast_ptr) or None)))

    def pair_table(self):
        out = np.empty((self.P, self.L, self.L), np.float32)
        self._ck(lib().msm_pair_table(self._h, out.ctypes.data_as(c_fp)))
        return out

    def pair_table_dev(self, dev_ptr):
        self._ck(lib().msm_pair_table_dev(self._h, C.c_void_p(dev_ptr)))

    def locate(self, pts):
        pts, pp = _d(pts)
        tri = np.empty(len(pts), np.int32); w = np.empty((len(pts), 3), np.float64)
        self._ck(lib().msm_locate(self._h, pp, len(pts), tri.ctypes.data_as(c_ip), w.ctypes.data_as(c_dp)))
        return tri, w

    def mesh_interpolate(self, pts, from_xyz, tris, to_xyz, radius=100.0, details=False):
        pts, pp = _d(pts); from_xyz, pf = _d(from_xyz); tris, pt = _i(tris); to_xyz, po = _d(to_xyz)
        assert to_xyz.shape == from_xyz.shape
        out = np.empty((len(pts), 3))
        if details:
            tri = np.empty(len(pts), np.int32); w = np.empty((len(pts), 3))
            self._ck(lib().msm_mesh_interpolate(self._h, pp, len(pts), pf, len(from_xyz), pt, len(tris), po, C.c_double(radius),
                                                out.ctypes.data_as(c_dp), tri.ctypes.data_as(c_ip), w.ctypes.data_as(c_dp)))
            return out, tri, w
        self._ck(lib().msm_mesh_interpolate(self._h, pp, len(pts), pf, len(from_xyz), pt, len(tris), po, C.c_double(radius),
                                            out.ctypes.data_as(c_dp), None, None))
        return out

    # ---- feature pre-processing (src/featurespace.cc:18-108) ------------------------------------------------
    @staticmethod
    def _mask(m, n):
        if m is None:
            return None, None, 0
        m = np.ascontiguousarray(np.atleast_2d(m), dtype=np.float64)
        assert m.shape[1] == n
        return m, m.ctypes.data_as(c_dp), m.shape[0]

    def smooth_data(self, in_xyz, values, ref_xyz, sigma, mask=None):
        in_xyz, p1 = _d(in_xyz); values = np.atleast_2d(values); values, p2 = _d(values); ref_xyz, p3 = _d(ref_xyz)
        assert values.shape[1] == len(in_xyz)
        m, pm, _ = self._mask(mask, len(in_xyz))
        out = np.empty((values.shape[0], len(ref_xyz)))
        self._ck(lib().msm_smooth_data(self._h, p1, len(in_xyz), p2, values.shape[0], pm, p3, len(ref_xyz), C.c_double(sigma),
                                       out.ctypes.data_as(c_dp)))
        return out

    def histogram_match(self, data, data_ref, excl=None, excl_ref=None, bins=256):
        a = np.array(np.atleast_2d(data), dtype=np.float64, order="C", copy=True)
        b, pb_ = _d(np.atleast_2d(data_ref))
        assert a.shape[0] == b.shape[0]
        ea, pea, ra = self._mask(excl, a.shape[1]); eb, peb, rb = self._mask(excl_ref, b.shape[1])
        self._ck(lib().msm_histogram_match(self._h, a.ctypes.data_as(c_dp), a.shape[1], pb_, b.shape[1], a.shape[0], pea, ra, peb, rb, int(bins)))
        return a

    def variance_normalise(self, data, mask=None):
        a = np.array(np.atleast_2d(data), dtype=np.float64, order="C", copy=True)
        m, pm, _ = self._mask(mask, a.shape[1])
        self._ck(lib().msm_variance_normalise(self._h, a.ctypes.data_as(c_dp), a.shape[0], a.shape[1], pm))
        return a

    def create_exclusion(self, data, lower, upper):
        a, pa = _d(np.atleast_2d(data))
        out = np.empty(a.shape[1])
        self._ck(lib().msm_create_exclusion(self._h, pa, a.shape[0], a.shape[1], C.c_float(lower), C.c_float(upper), out.ctypes.data_as(c_dp)))
        return out

    def total_cost(self, labeling):
        labeling, pl = _i(labeling); s = (C.c_double * 3)()
        self._ck(lib().msm_total_cost(self._h, pl, s))
        return float(s[0]), float(s[1]), float(s[2])

    # ---- introspection -------------------------------------------------------------------------------------
    def launch_count(self):
        return int(lib().msm_launch_count(self._h))

    def set_group_patches(self, off, keys, vals):
        off = np.ascontiguousarray(off, np.int64); keys, pk = _i(keys); vals, pv = _d(vals)
        self._ck(lib().msm_set_group_patches(self._h, C.c_longlong(len(off) - 1), off.ctypes.data_as(C.POINTER(C.c_longlong)), pk, pv))

    def group_pair_costs(self, q):
        q, pq = _i(q); out = np.empty(len(q), np.float64)
        self._ck(lib().msm_group_pair_costs(self._h, len(q), pq, out.ctypes.data_as(c_dp)))
        return out

    def group_pair_moves(self, labeling, alpha):
        labeling, pl = _i(labeling); out = np.empty((self.P, 4), np.float64)
        self._ck(lib().msm_group_pair_moves(self._h, pl, int(alpha), out.ctypes.data_as(c_dp)))
        return out

    def set_timing(self, mode=1):
        """0 off, 1 synchronous per-family timing, 2 asynchronous (timing_collect)."""
        self._ck(lib().msm_set_timing(self._h, int(mode)))

    def timing_collect(self, family):
        tot = C.c_double(); cnt = C.c_longlong()
        self._ck(lib().msm_timing_collect(self._h, family.encode(), C.byref(tot), C.byref(cnt)))
        return tot.value, cnt.value

    def last_kernel_ms(self, family):
        return float(lib().msm_last_kernel_ms(self._h, family.encode()))
