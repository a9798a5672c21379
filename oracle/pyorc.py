"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liborc.so (the CPU restatement of the reference hot path, oracle/oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Pinned against the reference's own sources compiled in place (oracle/pyrefcost.py, tests/test_reference_pin.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle_capi.cpp", "oracle.h", "geom.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_cf_create.restype = C.c_void_p
        _LIB.orc_gc_create.restype = C.c_void_p
        for f in ("orc_mesh_meanangle", "orc_corr", "orc_strain", "orc_cf_unary_cost", "orc_cf_pair_cost", "orc_cf_total", "orc_cf_meanangle"):
            getattr(_LIB, f).restype = C.c_double
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


def icosphere(level, radius=100.0):
    nv, nf = C.c_int(), C.c_int()
    lib().orc_icosphere_counts(level, C.byref(nv), C.byref(nf))
    xyz = np.empty((nv.value, 3), np.float64)
    tris = np.empty((nf.value, 3), np.int32)
    lib().orc_icosphere(level, C.c_double(radius), xyz.ctypes.data_as(c_dp), tris.ctypes.data_as(c_ip))
    return xyz, tris


def rotation_matrix(a, b):
    a, pa = _d(a); b, pb = _d(b)
    out = np.empty((3, 3), np.float64)
    lib().orc_rotation_matrix(pa, pb, out.ctypes.data_as(c_dp))
    return out


def label_grid(sgres, maxs_dist, cap=4096):
    s = np.empty((cap, 3)); b = np.empty((cap, 3)); c = np.empty(3)
    ns, nb = C.c_int(), C.c_int()
    lib().orc_label_grid(sgres, C.c_double(maxs_dist), cap, s.ctypes.data_as(c_dp), C.byref(ns), b.ctypes.data_as(c_dp),
                         C.byref(nb), c.ctypes.data_as(c_dp))
    assert ns.value <= cap and nb.value <= cap
    return s[:ns.value].copy(), b[:nb.value].copy(), c


def spacings(xyz, tris):
    xyz, px = _d(xyz); tris, pt = _i(tris)
    ms = np.empty(len(xyz)); mx, mvd = C.c_double(), C.c_double()
    lib().orc_spacings(px, len(xyz), pt, len(tris), ms.ctypes.data_as(c_dp), C.byref(mx), C.byref(mvd))
    return ms, mx.value, mvd.value


def triplets(xyz, tris):
    xyz, px = _d(xyz); tris, pt = _i(tris)
    out = np.empty((len(tris), 3), np.int32)
    lib().orc_triplets(px, len(xyz), pt, len(tris), out.ctypes.data_as(c_ip))
    return out


def pairs(xyz, tris):
    xyz, px = _d(xyz); tris, pt = _i(tris)
    cap = 3 * len(xyz)
    out = np.empty((cap, 2), np.int32)
    n = lib().orc_pairs(px, len(xyz), pt, len(tris), out.ctypes.data_as(c_ip), cap)
    return out[:n].copy()


def mesh_meanangle(xyz, tris):
    xyz, px = _d(xyz); tris, pt = _i(tris)
    return lib().orc_mesh_meanangle(px, len(xyz), pt, len(tris))


def rotations(centre, cps):
    centre, pc = _d(centre); cps, pp = _d(cps)
    out = np.empty((len(cps), 3, 3))
    lib().orc_rotations(pc, pp, len(cps), out.ctypes.data_as(c_dp))
    return out


def corr(a, b, w=None):
    a, pa = _d(a); b, pb = _d(b)
    if w is None:
        return lib().orc_corr(pa, pb, None, len(a), 0)
    w, pw = _d(w)
    return lib().orc_corr(pa, pb, pw, len(a), 1)


def strain(t0, t1, mu=0.4, kappa=1.6, k_exp=2.0, which=0):
    t0, p0 = _d(t0); t1, p1 = _d(t1)
    return lib().orc_strain(p0, p1, C.c_double(mu), C.c_double(kappa), C.c_double(k_exp), which)


def locate(xyz, tris, pts, brute=False, orthogonal=False):
    xyz, px = _d(xyz); tris, pt = _i(tris); pts, pp = _d(pts)
    tri = np.empty(len(pts), np.int32); w = np.empty((len(pts), 3))
    lib().orc_locate(px, len(xyz), pt, len(tris), pp, len(pts), tri.ctypes.data_as(c_ip), w.ctypes.data_as(c_dp), int(brute),
                     int(orthogonal))
    return tri, w


def metric_resample(in_xyz, in_tris, values, ref_xyz, ref_tris):
    """[EXT] A8 newresampler::metric_resample (adaptive barycentric), restated in oracle/geom.h: values[d][v_in] -> [d][v_ref]."""
    in_xyz, p1 = _d(in_xyz); in_tris, p2 = _i(in_tris); values = np.atleast_2d(values); values, p3 = _d(values)
    ref_xyz, p4 = _d(ref_xyz); ref_tris, p5 = _i(ref_tris)
    out = np.empty((values.shape[0], len(ref_xyz)))
    lib().orc_metric_resample(p1, len(in_xyz), p2, len(in_tris), p3, values.shape[0], p4, len(ref_xyz), p5, len(ref_tris), out.ctypes.data_as(c_dp))
    return out


# ---- feature pre-processing (oracle/feat.h: src/featurespace.cc:18-108, src/reg_tools.cpp:750-810) -----------------------------
def _mask(m, n):
    if m is None:
        return None, None, 0
    m = np.ascontiguousarray(np.atleast_2d(m), dtype=np.float64)
    assert m.shape[1] == n
    return m, m.ctypes.data_as(c_dp), m.shape[0]


def smooth_data(in_xyz, values, ref_xyz, sigma, mask=None):
    """[EXT] A9 newresampler::smooth_data restated: Gaussian of the geodesic distance, support 2 asin(sigma/RAD)."""
    in_xyz, p1 = _d(in_xyz); values = np.atleast_2d(values); values, p2 = _d(values); ref_xyz, p3 = _d(ref_xyz)
    m, pm, _ = _mask(mask, len(in_xyz))
    out = np.empty((values.shape[0], len(ref_xyz)))
    lib().orc_smooth_data(p1, len(in_xyz), p2, values.shape[0], p3, len(ref_xyz), C.c_double(sigma), pm, out.ctypes.data_as(c_dp))
    return out


def histogram_normalization(data_in, data_ref, excl_in=None, excl_ref=None):
    """multivariate_histogram_normalization (src/reg_tools.cpp:750-810) over the restated MISCMATHS::Histogram (A10)."""
    a = np.array(np.atleast_2d(data_in), dtype=np.float64, order="C", copy=True)
    b, pb_ = _d(np.atleast_2d(data_ref))
    ea, pea, ra = _mask(excl_in, a.shape[1]); eb, peb, rb = _mask(excl_ref, b.shape[1])
    lib().orc_histogram_normalization(a.ctypes.data_as(c_dp), a.shape[1], pb_, b.shape[1], a.shape[0], pea, ra, peb, rb)
    return a


def variance_normalise(data, mask=None):
    """featurespace::variance_normalise (src/featurespace.cc:67-108)."""
    a = np.array(np.atleast_2d(data), dtype=np.float64, order="C", copy=True)
    m, pm, _ = _mask(mask, a.shape[1])
    lib().orc_variance_normalise(a.ctypes.data_as(c_dp), a.shape[0], a.shape[1], pm)
    return a


def create_exclusion(data, lower, upper):
    """[EXT] A11 newresampler::create_exclusion restated."""
    a, pa = _d(np.atleast_2d(data))
    out = np.empty(a.shape[1])
    lib().orc_create_exclusion(pa, a.shape[0], a.shape[1], C.c_float(lower), C.c_float(upper), out.ctypes.data_as(c_dp))
    return out


def metric_resample_excl(in_xyz, in_tris, values, mask, ref_xyz, ref_tris):
    in_xyz, p1 = _d(in_xyz); in_tris, p2 = _i(in_tris); values = np.atleast_2d(values); values, p3 = _d(values)
    ref_xyz, p4 = _d(ref_xyz); ref_tris, p5 = _i(ref_tris)
    m, pm, _ = _mask(mask, len(in_xyz))
    out = np.empty((values.shape[0], len(ref_xyz)))
    lib().orc_metric_resample_excl(p1, len(in_xyz), p2, len(in_tris), p3, values.shape[0], pm, p4, len(ref_xyz), p5, len(ref_tris),
                                   out.ctypes.data_as(c_dp))
    return out


def featurespace_initialize(ico, mesh_in, data_in, mesh_ref, data_ref, sigma=(5.0, 5.0), thresholds=(0.0, 0.0), intensitynorm=False,
                            varnorm=False, cut=False, exclude=False):
    """featurespace::initialize (src/featurespace.cc:18-65) for the pairwise driver's (input, reference) pair.
    mesh_* = (xyz, tris); data_* = [D][V].  Returns (DATA[0], DATA[1], EXCL[0] or None, EXCL[1] or None)."""
    x0, p0 = _d(mesh_in[0]); t0, q0 = _i(mesh_in[1]); x1, p1 = _d(mesh_ref[0]); t1, q1 = _i(mesh_ref[1])
    d0 = np.atleast_2d(data_in); d0, pd0 = _d(d0); d1 = np.atleast_2d(data_ref); d1, pd1 = _d(d1)
    D = d0.shape[0]
    if ico > 0:
        nv, nf = C.c_int(), C.c_int()
        lib().orc_icosphere_counts(ico, C.byref(nv), C.byref(nf))
        v0 = v1 = nv.value
    else:
        v0, v1 = len(x0), len(x1)
    o0, o1 = np.empty((D, v0)), np.empty((D, v1))
    e0, e1 = np.full(len(x0), np.nan), np.full(len(x1), np.nan)
    lib().orc_featurespace_initialize.restype = C.c_int
    rc = lib().orc_featurespace_initialize(int(ico), p0, len(x0), q0, len(t0), pd0, p1, len(x1), q1, len(t1), pd1, D, C.c_double(sigma[0]),
                                           C.c_double(sigma[1]), C.c_float(thresholds[0]), C.c_float(thresholds[1]), int(intensitynorm),
                                           int(varnorm), int(cut), int(exclude), o0.ctypes.data_as(c_dp), o1.ctypes.data_as(c_dp),
                                           e0.ctypes.data_as(c_dp), e1.ctypes.data_as(c_dp))
    if rc < 0:
        raise RuntimeError("featurespace_initialize failed")
    has = exclude or cut
    return o0, o1, (e0 if has else None), (e1 if has else None)


def mesh_interpolate(pts, from_xyz, tris, to_xyz):
    """[EXT] barycentric_mesh_interpolation(SPH_up, SPH_low_init, SPH_low_final) restated (src/mesh_registration.cpp:390)."""
    pts, pp = _d(pts); from_xyz, pf = _d(from_xyz); tris, pt = _i(tris); to_xyz, po = _d(to_xyz)
    out = np.empty((len(pts), 3))
    lib().orc_mesh_interpolate(pp, len(pts), pf, len(from_xyz), pt, len(tris), po, out.ctypes.data_as(c_dp))
    return out


def unfold(xyz, tris):
    """unfold (src/reg_tools.cpp:134-181): returns (new coordinates, sweeps performed)."""
    xyz = np.array(xyz, dtype=np.float64, order="C", copy=True); tris, pt = _i(tris)
    it = lib().orc_unfold(xyz.ctypes.data_as(c_dp), len(xyz), pt, len(tris))
    return xyz, it


def count_folded(xyz, tris):
    xyz, px = _d(xyz); tris, pt = _i(tris)
    return lib().orc_count_folded(px, len(xyz), pt, len(tris))


class CostFunction:
    """Mirror of NonLinearSRegDiscreteCostFunction (+Uni/Multi variants), src/DiscreteCostFunction.{h,cpp}."""

    def __init__(self):
        self.h = C.c_void_p(lib().orc_cf_create())
        self.N = self.L = self.T = self.P = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_cf_destroy(self.h); self.h = None

    def set_params(self, lam=1.0, rexp=2.0, mu=0.4, kappa=1.6, k_exp=2.0, rng=1.0, simmeasure=2, rmode=3, dweight=False,
                   anorm=False, threads=1, bary_orth=False):
        lib().orc_cf_set_params(self.h, C.c_double(lam), C.c_double(rexp), C.c_double(mu), C.c_double(kappa), C.c_double(k_exp),
                                C.c_double(rng), simmeasure, rmode, int(dweight), int(anorm), threads, int(bary_orth))

    def set_meshes(self, txyz, ttris, sxyz, cxyz, ctris):
        txyz, p1 = _d(txyz); ttris, p2 = _i(ttris); sxyz, p3 = _d(sxyz); cxyz, p4 = _d(cxyz); ctris, p5 = _i(ctris)
        self.N, self.T, self.Vs, self.Vt = len(cxyz), len(ctris), len(sxyz), len(txyz)
        lib().orc_cf_set_meshes(self.h, p1, len(txyz), p2, len(ttris), p3, len(sxyz), p4, len(cxyz), p5, len(ctris))

    def set_orig(self, xyz):
        xyz, p = _d(xyz); lib().orc_cf_set_orig(self.h, p, len(xyz))

    def set_features(self, ref, inp, multivariate=None):
        ref = np.atleast_2d(ref); inp = np.atleast_2d(inp)
        ref, pr = _d(ref); inp, pi = _d(inp)
        if multivariate is None:
            multivariate = ref.shape[0] > 1
        self.C = ref.shape[0]
        lib().orc_cf_set_features(self.h, ref.shape[0], pr, ref.shape[1], pi, inp.shape[1], int(multivariate))

    def set_cfweight(self, w):
        w = np.atleast_2d(w); w, p = _d(w)
        lib().orc_cf_set_cfweight(self.h, w.shape[0], p, w.shape[1])

    def set_absweights(self, aw):
        aw, p = _d(aw); lib().orc_cf_set_absweights(self.h, p, len(aw))

    def absweights(self):
        """AbsoluteWeights after get_source_data (resample_weights, src/DiscreteCostFunction.cpp:364-383)."""
        out = np.empty(self.N); lib().orc_cf_get_absweights(self.h, out.ctypes.data_as(c_dp)); return out

    def set_spacings(self, maxsep, mvdmax):
        maxsep, p = _d(maxsep); lib().orc_cf_set_spacings(self.h, p, len(maxsep), C.c_double(mvdmax))

    def set_labels(self, labels, rot):
        labels, pl = _d(labels); rot, pr = _d(rot)
        self.L = len(labels)
        lib().orc_cf_set_labels(self.h, pl, len(labels), pr, len(rot))

    def reset_source(self, xyz):
        xyz, p = _d(xyz); lib().orc_cf_reset_source(self.h, p)

    def reset_cpgrid(self, xyz):
        xyz, p = _d(xyz); lib().orc_cf_reset_cpgrid(self.h, p)

    def set_pairs(self, pairs_):
        pairs_, p = _i(pairs_); self.P = len(pairs_); lib().orc_cf_set_pairs(self.h, p, len(pairs_))

    def set_triplets(self, t):
        t, p = _i(t); self.T = len(t); lib().orc_cf_set_triplets(self.h, p, len(t))

    def get_source_data(self, k0=None, k1=None):
        if k0 is None:
            lib().orc_cf_get_source_data(self.h)
        else:
            lib().orc_cf_get_source_data_range(self.h, k0, k1)

    def patches(self):
        sizes = np.empty(self.N, np.int32)
        lib().orc_cf_patch_sizes(self.h, sizes.ctypes.data_as(c_ip))
        out = []
        for k in range(self.N):
            a = np.empty(sizes[k], np.int32)
            lib().orc_cf_patch(self.h, k, a.ctypes.data_as(c_ip))
            out.append(a)
        return out

    def unary_costs(self, k0=None, k1=None):
        if k0 is None:
            out = np.empty((self.L, self.N))
            lib().orc_cf_unary_costs(self.h, out.ctypes.data_as(c_dp))
        else:
            out = np.empty((self.L, k1 - k0))
            lib().orc_cf_unary_costs_range(self.h, k0, k1, out.ctypes.data_as(c_dp))
        return out

    def unary_cost(self, node, label):
        return lib().orc_cf_unary_cost(self.h, node, label)

    def triplet_costs(self, q):
        q, p = _i(q)
        out = np.empty(len(q))
        lib().orc_cf_triplet_costs(self.h, len(q), p, out.ctypes.data_as(c_dp))
        return out

    def triplet_table(self, t0=0, t1=None):
        t1 = self.T if t1 is None else t1
        out = np.empty((t1 - t0, self.L, self.L, self.L))
        lib().orc_cf_triplet_table(self.h, t0, t1, out.ctypes.data_as(c_dp))
        return out

    def pair_table(self):
        out = np.empty((self.P, self.L, self.L))
        lib().orc_cf_pair_table(self.h, out.ctypes.data_as(c_dp))
        return out

    def pair_cost(self, p, la, lb):
        return lib().orc_cf_pair_cost(self.h, p, la, lb)

    def total(self, labeling, use_unary=True):
        labeling, p = _i(labeling)
        return lib().orc_cf_total(self.h, p, int(use_unary))

    def meanangle(self):
        return lib().orc_cf_meanangle(self.h)


class GroupCost:
    """Mirror of DiscreteGroupCostFunction, src/DiscreteGroupCostFunction.{h,cpp}."""

    def __init__(self):
        self.h = C.c_void_p(lib().orc_gc_create())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_gc_destroy(self.h); self.h = None

    def set(self, lam, rexp, mu, kappa, cp_xyz, cp_tris, orig_xyz, labels, rot, triplets_):
        cp_xyz, p1 = _d(cp_xyz); cp_tris, p2 = _i(cp_tris); orig_xyz, p3 = _d(orig_xyz)
        labels, p4 = _d(labels); rot, p5 = _d(rot); triplets_, p6 = _i(triplets_)
        S, N = cp_xyz.shape[0], cp_xyz.shape[1]
        self.L = len(labels)
        lib().orc_gc_set(self.h, C.c_double(lam), C.c_double(rexp), C.c_double(mu), C.c_double(kappa), S, p1, N, p2, len(cp_tris),
                         p3, p4, len(labels), p5, p6)

    def triplet_costs(self, q):
        q, p = _i(q); out = np.empty(len(q))
        lib().orc_gc_triplet_costs(self.h, len(q), p, out.ctypes.data_as(c_dp))
        return out

    def set_patches(self, off, keys, vals, pairs_):
        off = np.ascontiguousarray(off, np.int64); keys, pk = _i(keys); vals, pv = _d(vals); pairs_, pp = _i(pairs_)
        lib().orc_gc_set_patches(self.h, len(off) - 1, off.ctypes.data_as(C.POINTER(C.c_longlong)), pk, pv, pp, len(pairs_))

    def pair_costs(self, q):
        q, p = _i(q); out = np.empty(len(q))
        lib().orc_gc_pair_costs(self.h, len(q), p, self.L, out.ctypes.data_as(c_dp))
        return out
