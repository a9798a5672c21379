"""Synthetic registration problems shared by the tests.

Geometry (icospheres, label grid, spacings, cliques) comes from the ORACLE's restatement of the reference model glue
(src/DiscreteModel.cpp) — tests may use oracle/; product code may not.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyorc  # noqa: E402

RAD = 100.0


def f32(x):
    """float32-representable doubles (the reference's parameters are C floats, src/DiscreteCostFunction.h:121-133)."""
    return float(np.float32(x))


def random_field(xyz, seed, nwaves=32, kmax=6.0):
    """Band-limited random field on the sphere: sum of random low-order plane waves, ~unit variance."""
    rng = np.random.default_rng(seed)
    u = xyz / RAD
    k = rng.normal(size=(nwaves, 3))
    k *= (rng.uniform(1.0, kmax, size=(nwaves, 1)) / np.linalg.norm(k, axis=1, keepdims=True))
    ph = rng.uniform(0, 2 * np.pi, size=nwaves)
    amp = rng.normal(size=nwaves)
    f = (amp[None, :] * np.cos(u @ k.T + ph[None, :])).sum(1)
    return (f - f.mean()) / f.std()


def smooth_warp(xyz, seed, amplitude):
    """Smooth tangential displacement of sphere points (amplitude in mm on the R=100 sphere)."""
    rng = np.random.default_rng(seed)
    u = xyz / RAD
    d = np.zeros_like(u)
    for _ in range(6):
        k = rng.normal(size=3) * 2.0
        a = rng.normal(size=3)
        d += a[None, :] * np.sin(u @ k + rng.uniform(0, 2 * np.pi))[:, None]
    d -= (d * u).sum(1, keepdims=True) * u
    d *= amplitude / max(np.abs(d).max(), 1e-12)
    p = xyz + d
    return p * (RAD / np.linalg.norm(p, axis=1, keepdims=True))


def rotate(xyz, axis, deg):
    axis = np.asarray(axis, float); axis /= np.linalg.norm(axis)
    t = np.deg2rad(deg)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * K @ K
    return xyz @ R.T


class Problem:
    pass


def make_problem(data_level=4, cp_level=2, C=1, seed=1, warp=0.3, sg_offset=2, labeldist=0.5, iteration=1, deform_cp=0.0,
                 rough=False):
    """One outer iteration's worth of inputs for the cost function (src/DiscreteModel.cpp:38-89,188-236)."""
    P = Problem()
    P.data_level, P.cp_level, P.C = data_level, cp_level, C
    P.txyz, P.ttris = pyorc.icosphere(data_level)
    P.cxyz0, P.ctris = pyorc.icosphere(cp_level)
    P.maxsep, P.mvdmax, P.mvd = pyorc.spacings(P.cxyz0, P.ctris)
    # source mesh: the data icosphere pushed through a smooth warp (generic position) or left regular (exact ties)
    P.sxyz = smooth_warp(P.txyz, seed + 100, warp * P.mvd) if warp > 0 else P.txyz.copy()
    # control grid (optionally deformed, as after earlier iterations); ORIG = undeformed CP coordinates (nested ids)
    P.cxyz = smooth_warp(P.cxyz0, seed + 200, deform_cp * P.mvd) if deform_cp > 0 else P.cxyz0.copy()
    P.orig = P.cxyz0.copy()
    kmax = 40.0 if rough else 6.0
    P.ref = np.stack([random_field(P.txyz, seed * 10 + c, kmax=kmax) for c in range(C)])
    shared = random_field(P.txyz, seed * 10 + 99)
    if C > 1:
        P.ref = 0.8 * P.ref + 0.6 * shared[None, :]
    # source features: the reference field seen through a known rotation (so labels matter)
    moved = rotate(P.txyz, [0.3, -0.5, 0.8], 2.0)
    P.inp = np.stack([random_field(moved, seed * 10 + c, kmax=kmax) for c in range(C)])
    if C > 1:
        P.inp = 0.8 * P.inp + 0.6 * random_field(moved, seed * 10 + 99)[None, :]
    samples, bary, centre = pyorc.label_grid(cp_level + sg_offset, labeldist * P.mvdmax)
    P.samples, P.barycentres, P.centre = samples, bary, centre
    P.labels = bary if iteration % 2 == 1 else samples  # src/DiscreteModel.cpp:215-220
    P.rot = pyorc.rotations(centre, P.cxyz)
    P.triplets = pyorc.triplets(P.cxyz0, P.ctris)
    P.pairs = pyorc.pairs(P.cxyz0, P.ctris)
    P.N, P.L, P.T = len(P.cxyz), len(P.labels), len(P.triplets)
    return P


DEFAULT_PARAMS = dict(lam=f32(0.1), rexp=f32(2.0), mu=f32(0.4), kappa=f32(1.6), k_exp=f32(2.0), rng=f32(1.0), simmeasure=2, rmode=3)


def oracle_costfunction(P, params=None, cfweight=None, multivariate=None, threads=8, cliques="triplets"):
    pr = dict(DEFAULT_PARAMS); pr.update(params or {})
    cf = pyorc.CostFunction()
    cf.set_params(threads=threads, **pr)
    cf.set_meshes(P.txyz, P.ttris, P.sxyz, P.cxyz0, P.ctris)   # set_meshes sees the INITIAL grid (_oCPgrid)
    cf.reset_cpgrid(P.cxyz)
    cf.set_orig(P.orig)
    cf.set_features(P.ref, P.inp, multivariate)
    if cfweight is not None:
        cf.set_cfweight(cfweight)
    cf.set_spacings(P.maxsep, P.mvdmax)
    cf.set_labels(P.labels, P.rot)
    if cliques == "triplets":
        cf.set_triplets(P.triplets)
    else:
        cf.set_pairs(P.pairs)
    return cf


def gpu_context(P, params=None, cfweight=None, multivariate=None, cliques="triplets", device=0):
    from msm_newmeshreg_b200 import Context
    pr = dict(DEFAULT_PARAMS); pr.update(params or {})
    ctx = Context(device)
    ctx.set_target(P.txyz, P.ttris, P.ref)
    ctx.set_source(P.sxyz, P.inp, cfweight, multivariate)
    ctx.set_cpgrid(P.cxyz0, P.ctris, P.maxsep, P.orig)
    ctx.reset_cpgrid(P.cxyz)
    # MEANANGLE comes from the model (set_initial_angles) — the oracle restates it
    ctx.set_params(meanangle=pyorc.mesh_meanangle(P.cxyz0, P.ctris), mvdmax=P.mvdmax, **pr)
    ctx.set_labels(P.labels, P.centre)
    if cliques == "triplets":
        ctx.set_triplets(P.triplets)
    else:
        ctx.set_pairs(P.pairs)
    return ctx


def singular_nodes(P):
    """Control points (numerically) antipodal to the label-grid centre: ROT[k] = rotation(centre -> CP_k) is a rotation
    by pi about an undefined axis there (src/DiscreteModel.cpp:284-294 + [EXT] estimate_rotation_matrix), so every cost
    touching such a node is noise in the reference itself.  A regular icosphere contains exactly one such CP."""
    c = (P.cxyz @ P.centre) / (np.linalg.norm(P.cxyz, axis=1) * np.linalg.norm(P.centre))
    return np.nonzero(c < -1 + 1e-9)[0]


def rel_err(gpu, ref, floor):
    """|gpu-ref| / max(|ref|, floor): the tolerance form stated in SURVEY.md §7 (hard parts)."""
    gpu = np.asarray(gpu, np.float64); ref = np.asarray(ref, np.float64)
    return np.abs(gpu - ref) / np.maximum(np.abs(ref), floor)
