"""ORACLE / TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libref_cost.so: the REFERENCE'S OWN cost-evaluation
classes (src/DiscreteCostFunction.cpp, src/similarities.cc, src/reg_tools.cpp, src/DiscreteModel.cpp, src/DiscreteGroup*.cpp)
and optimisers (include/FastPD, include/ELC, include/Fusion), compiled unmodified where they lie under /root/reference
against the [EXT] stand-ins of oracle/ext/fsl_ext.h (recipe: oracle/Makefile `refcost`), behind the SAME flat entry points
as the product's host library (msm_newmeshreg_b200/host/harness.cpp).  It is the CPU reference that pins both the
oracle port (oracle/oracle.cpp) and the CUDA path.

The .so is built in the dev container only (the reference tree does not exist on the GPU box) and travels with the repo
snapshot.  Only tests, smoke() and bench.py's CPU-baseline legs may import this module."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libref_cost.so")
_lib = None
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(REF_LIB)


def lib():
    global _lib
    if _lib is None:
        from msm_newmeshreg_b200 import host
        L = C.CDLL(REF_LIB)
        L.msmhost_model_triplet_moves.argtypes = None
        _lib = _bind(L, host)
    return _lib


def _bind(L, host):
    # same prototypes as the product's host library, minus the device-context accessor the reference does not have
    L.msmhost_last_error.restype = C.c_char_p
    L.msmhost_model_create.restype = C.c_void_p
    L.msmhost_model_create.argtypes = [C.c_double] * 6 + [C.c_int] * 6 + [C.c_double, C.c_int, C.c_char_p, C.c_int]
    L.msmhost_model_total.restype = C.c_double
    return L


def model(**kw):
    """host.Model API (set_data / initialize / reset_meshspace / reset_cpgrid / setup / unary / costs / labels / pairs /
    triplets / set_labeling / total / apply_labeling / triplet_moves) over the reference's own NonLinearSRegDiscreteModel."""
    from msm_newmeshreg_b200 import host
    return host.Model(_lib_override=lib(), **kw)


def icosphere(level, radius=100.0):
    from msm_newmeshreg_b200 import host
    return host.icosphere(level, radius, _lib_override=lib())


def mesh_interpolate(up, init_xyz, tris, final_xyz):
    from msm_newmeshreg_b200 import host
    return host.mesh_interpolate(up, init_xyz, tris, final_xyz, _lib_override=lib())


def unfold(xyz, tris):
    """newmeshreg::unfold — the reference's own src/reg_tools.cpp:134-181."""
    from msm_newmeshreg_b200 import host
    return host.unfold(xyz, tris, _lib_override=lib())


def patches(m):
    """Per-control-point source patches (`_sourceinrange`) of the reference model `m` after setup(): list of int arrays."""
    sizes = np.zeros(m.N, np.int32)
    if lib().ref_model_patches(m.h, sizes.ctypes.data_as(c_ip), None) != 0:
        raise RuntimeError(lib().msmhost_last_error().decode())
    idx = np.zeros(int(sizes.sum()), np.int32)
    lib().ref_model_patches(m.h, sizes.ctypes.data_as(c_ip), idx.ctypes.data_as(c_ip))
    return np.split(idx, np.cumsum(sizes)[:-1])


class GroupCost:
    """The reference's DiscreteGroupCostFunction (src/DiscreteGroupCostFunction.{h,cpp}); interface of pyorc.GroupCost."""

    def __init__(self):
        lib().ref_gc_create.restype = C.c_void_p
        self.h = C.c_void_p(lib().ref_gc_create())

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_gc_destroy(self.h); self.h = None

    def _ck(self, st):
        if st != 0:
            raise RuntimeError(lib().msmhost_last_error().decode())

    def set(self, lam, rexp, mu, kappa, cp_xyz, cp_tris, orig_xyz, labels, rot, triplets_):
        d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        cp_xyz, cp_tris, orig_xyz, labels, rot, triplets_ = d(cp_xyz), i(cp_tris), d(orig_xyz), d(labels), d(rot), i(triplets_)
        S, N = cp_xyz.shape[0], cp_xyz.shape[1]
        self.L = len(labels)
        self._ck(lib().ref_gc_set(self.h, C.c_double(lam), C.c_double(rexp), C.c_double(mu), C.c_double(kappa), S,
                                  cp_xyz.ctypes.data_as(c_dp), N, cp_tris.ctypes.data_as(c_ip), len(cp_tris), orig_xyz.ctypes.data_as(c_dp),
                                  labels.ctypes.data_as(c_dp), len(labels), rot.ctypes.data_as(c_dp), triplets_.ctypes.data_as(c_ip)))

    def triplet_costs(self, q):
        q = np.ascontiguousarray(q, dtype=np.int32); out = np.empty(len(q))
        self._ck(lib().ref_gc_triplet_costs(self.h, len(q), q.ctypes.data_as(c_ip), out.ctypes.data_as(c_dp)))
        return out

    def set_patches(self, off, keys, vals, pairs_):
        off = np.ascontiguousarray(off, np.int64); keys = np.ascontiguousarray(keys, np.int32)
        vals = np.ascontiguousarray(vals, np.float64); pairs_ = np.ascontiguousarray(pairs_, np.int32)
        self._ck(lib().ref_gc_set_patches(self.h, len(off) - 1, off.ctypes.data_as(C.POINTER(C.c_longlong)), keys.ctypes.data_as(c_ip),
                                          vals.ctypes.data_as(c_dp), pairs_.ctypes.data_as(c_ip), len(pairs_)))

    def pair_costs(self, q):
        q = np.ascontiguousarray(q, dtype=np.int32); out = np.empty(len(q))
        self._ck(lib().ref_gc_pair_costs(self.h, len(q), q.ctypes.data_as(c_ip), self.L, out.ctypes.data_as(c_dp)))
        return out


def fastpd_model(m):
    e = C.c_double()
    if lib().ref_fastpd_model(m.h, C.byref(e)) != 0:
        raise RuntimeError(lib().msmhost_last_error().decode())
    return m.labeling(), e.value


def fusion_model(m, threads=1):
    e = C.c_double()
    if lib().ref_fusion_model(m.h, threads, C.byref(e)) != 0:
        raise RuntimeError(lib().msmhost_last_error().decode())
    return m.labeling(), e.value


def driver_run_discrete_opt(m, source_xyz, iters, threads=1, mciters=0):
    """Mesh_registration::run_discrete_opt ITSELF — the text of src/mesh_registration.cpp:316-397, cut out at build time and
    compiled unmodified (oracle/Makefile, host/harness.cpp WITH_REF_DRIVER_LOOP) — over the reference's own classes, all on the CPU.
    Returns (warped source, final control grid, final labelling)."""
    src = np.array(source_xyz, dtype=np.float64, order="C", copy=True)
    grid = np.zeros((m.N, 3)); lab = np.zeros(m.N, np.int32)
    if lib().ref_driver_run_discrete_opt(m.h, int(iters), int(threads), int(mciters), src.ctypes.data_as(c_dp), grid.ctypes.data_as(c_dp),
                                         lab.ctypes.data_as(c_ip)) != 0:
        raise RuntimeError(lib().msmhost_last_error().decode())
    return src, grid, lab


def has_driver_loop():
    return hasattr(lib(), "ref_driver_run_discrete_opt")


def run_discrete_opt(m, source_xyz, iters, threads=1):
    """Mesh_registration::run_discrete_opt (src/mesh_registration.cpp:316-397) — reference cost functions AND reference
    optimisers, all on the CPU.  Same return convention as oracle.pyref.run_discrete_opt."""
    src = np.array(source_xyz, dtype=np.float64, order="C", copy=True)
    labs = np.zeros((iters, m.N), np.int32)
    en = np.zeros(iters); tc = np.zeros(iters); to = np.zeros(iters); tw = np.zeros(iters)
    n = lib().ref_run_discrete_opt(m.h, iters, threads, src.ctypes.data_as(c_dp), labs.ctypes.data_as(c_ip), en.ctypes.data_as(c_dp),
                                   tc.ctypes.data_as(c_dp), to.ctypes.data_as(c_dp), tw.ctypes.data_as(c_dp))
    if n < 0:
        raise RuntimeError(lib().msmhost_last_error().decode())
    return src, labs[:n], en[:n], {"cost": tc[:n], "optimiser": to[:n], "warp": tw[:n]}
