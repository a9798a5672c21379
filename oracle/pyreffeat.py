"""TEST INFRASTRUCTURE: ctypes access to oracle/_ref/libref_feat.so — the reference's OWN feature pre-processing
(/root/reference/src/featurespace.cc + src/reg_tools.cpp compiled where they lie, oracle/Makefile `reffeat`).
Only tests/ may import this module.  The library is prebuilt in the dev container and travels to the GPU box."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_feat.so")
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.ref_featurespace_initialize.restype = C.c_int
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


def featurespace_initialize(ico, mesh_in, data_in, mesh_ref, data_ref, sigma=(5.0, 5.0), thresholds=(0.0, 0.0), intensitynorm=False,
                            varnorm=False, cut=False, exclude=False, threads=1):
    """newmeshreg::featurespace(datain, dataref) + setters + initialize(ico, IN, exclude) of the reference itself.
    Returns (DATA[0], DATA[1], EXCL[0] or None, EXCL[1] or None), data as [D][V]."""
    x0, p0 = _d(mesh_in[0]); t0, q0 = _i(mesh_in[1]); x1, p1 = _d(mesh_ref[0]); t1, q1 = _i(mesh_ref[1])
    d0, pd0 = _d(np.atleast_2d(data_in)); d1, pd1 = _d(np.atleast_2d(data_ref))
    D = d0.shape[0]
    if ico > 0:
        v0 = v1 = 10 * 4 ** ico + 2
    else:
        v0, v1 = len(x0), len(x1)
    o0, o1 = np.empty((D, v0)), np.empty((D, v1))
    e0, e1 = np.full(len(x0), np.nan), np.full(len(x1), np.nan)
    rc = lib().ref_featurespace_initialize(int(ico), p0, len(x0), q0, len(t0), pd0, p1, len(x1), q1, len(t1), pd1, D, C.c_double(sigma[0]),
                                           C.c_double(sigma[1]), C.c_float(thresholds[0]), C.c_float(thresholds[1]), int(intensitynorm),
                                           int(varnorm), int(cut), int(exclude), int(threads), o0.ctypes.data_as(c_dp),
                                           o1.ctypes.data_as(c_dp), e0.ctypes.data_as(c_dp), e1.ctypes.data_as(c_dp))
    if rc < 0:
        raise RuntimeError("the reference's featurespace::initialize threw")
    has = exclude or cut
    return o0, o1, (e0 if has else None), (e1 if has else None)
