"""ORACLE / TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libref_optim.so: the reference's own FastPD / ELC-HOCR /
Fusion optimisers compiled unmodified from /root/reference/include (recipe: oracle/Makefile `ref`).  The .so is built in
the dev container only (the reference tree does not exist on the GPU box) and travels with the repo snapshot.
Only tests may import this module."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libref_optim.so")
_lib = None
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(REF_LIB)


def lib():
    global _lib
    if _lib is None:
        from msm_newmeshreg_b200 import _capi, host
        _capi.lib()
        _lib = host.bind(C.CDLL(REF_LIB))
    return _lib


def _d(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


def model(**kw):
    """A GPU-backed host Model living inside the reference-optimiser library (same harness code)."""
    from msm_newmeshreg_b200 import host
    return host.Model(_lib_override=lib(), **kw)


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
    compiled unmodified (oracle/Makefile, host/harness.cpp WITH_REF_DRIVER_LOOP) — over the GPU-backed model `m` (the product's mirror classes).
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
    """Mesh_registration::run_discrete_opt (src/mesh_registration.cpp:316-397) on the GPU-backed model `m` with the
    reference's own optimiser (FastPD or HOCR per the model's dOPT).  Returns (warped source, labelings[it][N],
    energies[it], timings dict of per-iteration seconds)."""
    src = np.array(source_xyz, dtype=np.float64, order="C", copy=True)
    labs = np.zeros((iters, m.N), np.int32)
    en = np.zeros(iters); tc = np.zeros(iters); to = np.zeros(iters); tw = np.zeros(iters)
    n = lib().ref_run_discrete_opt(m.h, iters, threads, src.ctypes.data_as(c_dp), labs.ctypes.data_as(c_ip), en.ctypes.data_as(c_dp),
                                   tc.ctypes.data_as(c_dp), to.ctypes.data_as(c_dp), tw.ctypes.data_as(c_dp))
    if n < 0:
        raise RuntimeError(lib().msmhost_last_error().decode())
    return src, labs[:n], en[:n], {"cost": tc[:n], "optimiser": to[:n], "warp": tw[:n]}


def fastpd_tables(unary, pairs, pairtab):
    """FPD::FastPD (include/FastPD/FastPD.h) on explicit tables: unary [L][N], pairtab [P][L][L] (reference layout)."""
    L, N = unary.shape
    unary, pu = _d(unary); pairs, pp = _i(pairs); pairtab, pt = _d(pairtab)
    lab = np.zeros(N, np.int32); e = C.c_double()
    if lib().ref_fastpd_tables(N, L, len(pairs), pp, pu, pt, lab.ctypes.data_as(c_ip), C.byref(e)) != 0:
        raise RuntimeError(lib().msmhost_last_error().decode())
    return lab, e.value


def fusion_tables(unary, triplets=None, triptab=None, pairs=None, pairtab=None, threads=1):
    """Fusion::optimize (include/Fusion/Fusion.h: HOCR reduction + inner FastPD) on explicit tables."""
    L, N = unary.shape
    unary, pu = _d(unary)
    triplets, ptr = _i(triplets); triptab, ptt = _d(triptab); pairs, pp = _i(pairs); pairtab, ppt = _d(pairtab)
    T = 0 if triplets is None else len(triplets)
    P = 0 if pairs is None else len(pairs)
    lab = np.zeros(N, np.int32); e = C.c_double()
    if lib().ref_fusion_tables(N, L, P, pp, T, ptr, pu, ppt, ptt, threads, lab.ctypes.data_as(c_ip), C.byref(e)) != 0:
        raise RuntimeError(lib().msmhost_last_error().decode())
    return lab, e.value


def fusion_tables_model(m, threads=1):
    """The product's table-aware fusion driver (msm_newmeshreg_b200/host/src/FusionTables.h: batched GPU tables feeding
    the reference's ELC / FastPD cores) on the GPU-backed model `m`; same return convention as fusion_model."""
    e = C.c_double()
    if lib().ref_fusion_tables_model(m.h, threads, C.byref(e)) != 0:
        raise RuntimeError(lib().msmhost_last_error().decode())
    return m.labeling(), e.value
