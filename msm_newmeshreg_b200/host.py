"""ctypes binding of libmsm_host.so — the host-side C++ mirror of the reference's DiscreteModel / DiscreteCostFunction
classes (msm_newmeshreg_b200/host/src), driven through the flat entry points of host/harness.cpp.

`Model` follows the reference driver's call sequence (src/mesh_registration.cpp:316-397): set_data -> initialize ->
[reset_meshspace -> setup -> unary / per-element costs -> (optimiser) -> apply_labeling]*.
Every cost is evaluated on the GPU (no CPU fallback); geometry helpers (icosphere, label grid, cliques) are host code.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libmsm_host.so")
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
_lib = None


def bind(lib):
    lib.msmhost_last_error.restype = C.c_char_p
    lib.msmhost_model_create.restype = C.c_void_p
    lib.msmhost_model_create.argtypes = [C.c_double] * 6 + [C.c_int] * 6 + [C.c_double, C.c_int, C.c_char_p, C.c_int]
    lib.msmhost_model_total.restype = C.c_double
    for f in ("msmhost_model_destroy", "msmhost_model_set_data", "msmhost_model_initialize", "msmhost_model_sizes",
              "msmhost_model_reset_meshspace", "msmhost_model_reset_cpgrid", "msmhost_model_setup", "msmhost_model_unary",
              "msmhost_model_get", "msmhost_model_costs", "msmhost_model_set_labeling", "msmhost_model_get_labeling",
              "msmhost_model_total", "msmhost_model_apply_labeling", "msmhost_model_triplet_moves"):
        getattr(lib, f).argtypes = None
    return lib


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
        from . import _capi
        _capi.lib()  # libmsm_b200.so first (dependency)
        _lib = bind(C.CDLL(HOST_LIB_PATH))
    return _lib


class HostError(RuntimeError):
    pass


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


def icosphere(level, radius=100.0, _lib_override=None):
    L = _lib_override or lib()
    nv, nf = C.c_int(), C.c_int()
    L.msmhost_icosphere_counts(level, C.byref(nv), C.byref(nf))
    xyz = np.empty((nv.value, 3)); tris = np.empty((nf.value, 3), np.int32)
    L.msmhost_icosphere(level, C.c_double(radius), xyz.ctypes.data_as(c_dp), tris.ctypes.data_as(c_ip))
    return xyz, tris


def mesh_interpolate(up, init_xyz, tris, final_xyz, _lib_override=None):
    """newresampler::barycentric_mesh_interpolation(SPH_up, SPH_low_init, SPH_low_final) — GPU; returns the moved points."""
    L = _lib_override or lib()
    up = np.array(up, dtype=np.float64, order="C", copy=True)
    init_xyz, pi = _d(init_xyz); tris, pt = _i(tris); final_xyz, pf = _d(final_xyz)
    if L.msmhost_mesh_interpolate(up.ctypes.data_as(c_dp), len(up), pi, len(init_xyz), pt, len(tris), pf) != 0:
        raise HostError(L.msmhost_last_error().decode())
    return up


def metric_resample(in_xyz, in_tris, values, ref_xyz, ref_tris, _lib_override=None):
    """newresampler::metric_resample ([EXT] A8, adaptive barycentric): values[d][v_in] on mesh `in` -> [d][v_ref] on mesh
    `ref`; both batches of point locations on the GPU."""
    L = _lib_override or lib()
    in_xyz, p1 = _d(in_xyz); in_tris, p2 = _i(in_tris); values = np.atleast_2d(values); values, p3 = _d(values)
    ref_xyz, p4 = _d(ref_xyz); ref_tris, p5 = _i(ref_tris)
    out = np.empty((values.shape[0], len(ref_xyz)))
    if L.msmhost_metric_resample(p1, len(in_xyz), p2, len(in_tris), p3, values.shape[0], p4, len(ref_xyz), p5, len(ref_tris), out.ctypes.data_as(c_dp)) != 0:
        raise HostError(L.msmhost_last_error().decode())
    return out


def featurespace_initialize(ico, mesh_in, data_in, mesh_ref, data_ref, sigma=(5.0, 5.0), thresholds=(0.0, 0.0), intensitynorm=False,
                            varnorm=False, cut=False, exclude=False):
    """newmeshreg::featurespace + initialize(ico, IN, exclude) (src/featurespace.cc:18-65), every step on the GPU.
    mesh_* = (xyz, tris); data_* = [D][V].  Returns (DATA[0], DATA[1], EXCL[0] or None, EXCL[1] or None), data as [D][V_template]."""
    L = lib()
    x0, p0 = _d(mesh_in[0]); t0, q0 = _i(mesh_in[1]); x1, p1 = _d(mesh_ref[0]); t1, q1 = _i(mesh_ref[1])
    d0, pd0 = _d(np.atleast_2d(data_in)); d1, pd1 = _d(np.atleast_2d(data_ref))
    D = d0.shape[0]
    assert d0.shape == (D, len(x0)) and d1.shape == (D, len(x1))
    v0, v1 = (10 * 4 ** ico + 2,) * 2 if ico > 0 else (len(x0), len(x1))
    o0, o1 = np.empty((D, v0)), np.empty((D, v1))
    e0, e1 = np.full(len(x0), np.nan), np.full(len(x1), np.nan)
    L.msmhost_featurespace_initialize.restype = C.c_int
    rc = L.msmhost_featurespace_initialize(int(ico), p0, len(x0), q0, len(t0), pd0, p1, len(x1), q1, len(t1), pd1, D, C.c_double(sigma[0]),
                                           C.c_double(sigma[1]), C.c_float(thresholds[0]), C.c_float(thresholds[1]), int(intensitynorm),
                                           int(varnorm), int(cut), int(exclude), 1, o0.ctypes.data_as(c_dp), o1.ctypes.data_as(c_dp),
                                           e0.ctypes.data_as(c_dp), e1.ctypes.data_as(c_dp))
    if rc < 0:
        raise HostError(L.msmhost_last_error().decode())
    has = exclude or cut
    return o0, o1, (e0 if has else None), (e1 if has else None)


def unfold(xyz, tris, _lib_override=None):
    """newmeshreg::unfold (src/reg_tools.cpp:134-181), host side."""
    L = _lib_override or lib()
    xyz = np.array(xyz, dtype=np.float64, order="C", copy=True); tris, pt = _i(tris)
    if L.msmhost_unfold(xyz.ctypes.data_as(c_dp), len(xyz), pt, len(tris)) != 0:
        raise HostError(L.msmhost_last_error().decode())
    return xyz


class Model:
    """NonLinearSRegDiscreteModel (+ its GPU-backed cost function) behind the harness."""

    def __init__(self, lam=0.1, rexp=2.0, mu=0.4, kappa=1.6, k_exp=2.0, rng=1.0, simmeasure=2, rmode=3, dweight=False, anorm=False,
                 sgres=4, cpres=2, labeldist=0.5, multivariate=False, dopt="HOCR", threads=1, _lib_override=None):
        self._L = _lib_override or lib()
        self.h = self._L.msmhost_model_create(lam, rexp, mu, kappa, k_exp, rng, simmeasure, rmode, int(dweight), int(anorm), sgres, cpres,
                                              labeldist, int(multivariate), dopt.encode(), threads)
        if not self.h:
            raise HostError(self._L.msmhost_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.cpres = cpres

    def close(self):
        if getattr(self, "h", None):
            self._L.msmhost_model_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise HostError(self._L.msmhost_last_error().decode())

    def set_data(self, txyz, ttris, sxyz, stris, inp, ref):
        txyz, p1 = _d(txyz); ttris, p2 = _i(ttris); sxyz, p3 = _d(sxyz); stris, p4 = _i(stris)
        inp = np.atleast_2d(inp); ref = np.atleast_2d(ref); inp, p5 = _d(inp); ref, p6 = _d(ref)
        self.Vs = len(sxyz)
        self._ck(self._L.msmhost_model_set_data(self.h, p1, len(txyz), p2, len(ttris), p3, len(sxyz), p4, len(stris), p5, p6, inp.shape[0]))

    def initialize(self):
        self._ck(self._L.msmhost_model_initialize(self.h, self.cpres))
        self._sizes()

    def _sizes(self):
        n, l, p, t = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._L.msmhost_model_sizes(self.h, C.byref(n), C.byref(l), C.byref(p), C.byref(t))
        self.N, self.L, self.P, self.T = n.value, l.value, p.value, t.value

    def reset_meshspace(self, sxyz):
        sxyz, p = _d(sxyz); self._ck(self._L.msmhost_model_reset_meshspace(self.h, p))

    def reset_cpgrid(self, cxyz):
        cxyz, p = _d(cxyz); self._ck(self._L.msmhost_model_reset_cpgrid(self.h, p))

    def set_weights(self, w):
        """cost-function weighting (setupCostFunctionWeighting, src/mesh_registration.cpp:323-334): w[rows][source vertices]."""
        w = np.atleast_2d(w); w, p = _d(w)
        self._ck(self._L.msmhost_model_set_weights(self.h, p, w.shape[0], w.shape[1]))

    def setup(self):
        self._ck(self._L.msmhost_model_setup(self.h)); self._sizes()

    def unary(self):
        out = np.empty((self.L, self.N))
        self._ck(self._L.msmhost_model_unary(self.h, out.ctypes.data_as(c_dp)))
        return out

    def cpgrid(self):
        out = np.empty((self.N, 3)); self._ck(self._L.msmhost_model_get(self.h, out.ctypes.data_as(c_dp), None, None, None)); return out

    def labels(self):
        out = np.empty((self.L, 3)); self._ck(self._L.msmhost_model_get(self.h, None, out.ctypes.data_as(c_dp), None, None)); return out

    def pairs(self):
        out = np.empty((self.P, 2), np.int32); self._ck(self._L.msmhost_model_get(self.h, None, None, out.ctypes.data_as(c_ip), None)); return out

    def triplets(self):
        out = np.empty((self.T, 3), np.int32); self._ck(self._L.msmhost_model_get(self.h, None, None, None, out.ctypes.data_as(c_ip))); return out

    def costs(self, kind, q):
        q, pq = _i(q); out = np.empty(len(q))
        self._ck(self._L.msmhost_model_costs(self.h, {"unary": 0, "pair": 1, "triplet": 2}[kind], len(q), pq, out.ctypes.data_as(c_dp)))
        return out

    def set_labeling(self, lab):
        lab, p = _i(lab); self._L.msmhost_model_set_labeling(self.h, p)

    def labeling(self):
        out = np.empty(self.N, np.int32); self._L.msmhost_model_get_labeling(self.h, out.ctypes.data_as(c_ip)); return out

    def total(self):
        return float(self._L.msmhost_model_total(self.h))

    def apply_labeling(self):
        out = np.empty((self.N, 3)); self._ck(self._L.msmhost_model_apply_labeling(self.h, out.ctypes.data_as(c_dp))); return out

    def context(self):
        """The cost function's device context as a (borrowed) msm_newmeshreg_b200.Context."""
        from ._capi import Context
        self._L.msmhost_model_context.restype = C.c_void_p
        return Context.from_handle(self._L.msmhost_model_context(self.h), self.N, self.L, self.T, self.P)

    def triplet_moves(self, labeling, alpha, out=None):
        """`out`: optional caller-owned float64 buffer (e.g. pinned memory) of shape [L][T][8] / [T][8]."""
        labeling, p = _i(labeling)
        if out is None:
            out = np.empty((self.L, self.T, 8) if alpha < 0 else (self.T, 8))
        if out.dtype == np.float32:   # fp32 table (GPU-backed library only): half the read-back
            self._ck(self._L.msmhost_model_triplet_moves_f32(self.h, p, int(alpha), out.ctypes.data_as(C.POINTER(C.c_float))))
        else:
            self._ck(self._L.msmhost_model_triplet_moves(self.h, p, int(alpha), out.ctypes.data_as(c_dp)))
        return out


class GroupModel:
    """DiscreteGroupModel (+ its GPU-backed DiscreteGroupCostFunction) behind the harness: the call sequence of
    Group_Mesh_registration (src/group_mesh_registration.cpp:20-36,49-108).  Global node id = subject * N_subj + vertex."""

    def __init__(self, lam=0.1, rexp=2.0, mu=0.4, kappa=1.6, rng=1.0, sgres=4, cpres=2, labeldist=0.5, threads=1, _lib_override=None):
        self._L = _lib_override or lib()
        self._L.msmhost_group_create.restype = C.c_void_p
        self._L.msmhost_group_create.argtypes = [C.c_double] * 5 + [C.c_int, C.c_int, C.c_double, C.c_int]
        self._L.msmhost_group_total.restype = C.c_double
        h = self._L.msmhost_group_create(lam, rexp, mu, kappa, rng, sgres, cpres, labeldist, threads)
        if not h:
            raise HostError(self._L.msmhost_last_error().decode())
        self.h = C.c_void_p(h)
        self.cpres = cpres

    def close(self):
        if getattr(self, "h", None):
            self._L.msmhost_group_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise HostError(self._L.msmhost_last_error().decode())

    def set_data(self, txyz, ttris, data):
        """template icosphere + data[s][v]: one feature row per subject on the template's vertices."""
        txyz, p1 = _d(txyz); ttris, p2 = _i(ttris); data = np.atleast_2d(data); data, p3 = _d(data)
        assert data.shape[1] == len(txyz)
        self.S, self.Vt = data.shape[0], len(txyz)
        self._ck(self._L.msmhost_group_set_data(self.h, p1, len(txyz), p2, len(ttris), self.S, p3))

    def initialize(self):
        self._ck(self._L.msmhost_group_initialize(self.h, self.cpres)); self._sizes()

    def _sizes(self):
        n, l, p, t = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._L.msmhost_group_sizes(self.h, C.byref(n), C.byref(l), C.byref(p), C.byref(t))
        self.N, self.L, self.P, self.T = n.value, l.value, p.value, t.value
        self.N_subj = self.N // max(self.S, 1)

    def reset_meshspace(self, subject, xyz):
        xyz, p = _d(xyz); self._ck(self._L.msmhost_group_reset_meshspace(self.h, int(subject), p))

    def reset_cpgrid(self, subject, xyz):
        xyz, p = _d(xyz); self._ck(self._L.msmhost_group_reset_cpgrid(self.h, int(subject), p))

    def setup(self):
        self._ck(self._L.msmhost_group_setup(self.h)); self._sizes()

    def cpgrids(self):
        out = np.empty((self.S, self.N_subj, 3)); self._ck(self._L.msmhost_group_get(self.h, out.ctypes.data_as(c_dp), None, None, None)); return out

    def labels(self):
        out = np.empty((self.L, 3)); self._ck(self._L.msmhost_group_get(self.h, None, out.ctypes.data_as(c_dp), None, None)); return out

    def pairs(self):
        out = np.empty((self.P, 2), np.int32); self._ck(self._L.msmhost_group_get(self.h, None, None, out.ctypes.data_as(c_ip), None)); return out

    def triplets(self):
        out = np.empty((self.T, 3), np.int32); self._ck(self._L.msmhost_group_get(self.h, None, None, None, out.ctypes.data_as(c_ip))); return out

    def costs(self, kind, q):
        q, pq = _i(q); out = np.empty(len(q))
        self._ck(self._L.msmhost_group_costs(self.h, {"pair": 1, "triplet": 2}[kind], len(q), pq, out.ctypes.data_as(c_dp)))
        return out

    def set_labeling(self, lab):
        lab, p = _i(lab); self._L.msmhost_group_set_labeling(self.h, p)

    def labeling(self):
        out = np.empty(self.N, np.int32); self._L.msmhost_group_get_labeling(self.h, out.ctypes.data_as(c_ip)); return out

    def total(self):
        return float(self._L.msmhost_group_total(self.h))

    def apply_labeling(self):
        out = np.empty((self.S, self.N_subj, 3)); self._ck(self._L.msmhost_group_apply_labeling(self.h, out.ctypes.data_as(c_dp))); return out

    def rotated_values(self, subject, label):
        """rotated_meshes[subject * L + label] (src/DiscreteGroupModel.cpp:92) on the template's vertices (product mirror only)."""
        out = np.empty(self.Vt)
        self._ck(self._L.msmhost_group_rotated(self.h, int(subject), int(label), out.ctypes.data_as(c_dp)))
        return out

    def pair_moves(self, labeling, alpha):
        """The 4 fusion-move costs of every inter-subject pair for proposal label alpha: [P][4] (one GPU launch)."""
        labeling, p = _i(labeling); out = np.empty((self.P, 4))
        self._ck(self._L.msmhost_group_pair_moves(self.h, p, int(alpha), out.ctypes.data_as(c_dp)))
        return out

    # multi-GPU: subjects [s0, s1) are produced by this rank; exchange rotated_all / set_rotated, then finish_setup
    def set_subject_shard(self, s0, s1):
        self._ck(self._L.msmhost_group_set_subject_shard(self.h, int(s0), int(s1)))

    def rotated_all(self, subject):
        out = np.empty((self.L, self.Vt)); self._ck(self._L.msmhost_group_rotated_all(self.h, int(subject), out.ctypes.data_as(c_dp))); return out

    def set_rotated(self, subject, values):
        values, p = _d(values); assert values.shape == (self.L, self.Vt)
        self._ck(self._L.msmhost_group_set_rotated(self.h, int(subject), p))

    def finish_setup(self):
        self._ck(self._L.msmhost_group_finish_setup(self.h))

    def fusion(self, threads=1, tables=False):
        """Fusion::optimize (the reference's, unmodified) or, tables=True, the batched FusionTables::optimize_group — only
        in libraries built with the reference's optimisers (oracle/_ref)."""
        e = C.c_double()
        f = self._L.ref_fusion_tables_group if tables else self._L.ref_fusion_group
        self._ck(f(self.h, int(threads), C.byref(e)))
        return self.labeling(), e.value
