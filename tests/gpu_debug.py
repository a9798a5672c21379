"""Scratch GPU debugging (not a pytest file)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import problems as pb  # noqa: E402
from oracle import pyorc  # noqa: E402

P = pb.make_problem(4, 2, 1, seed=12, warp=0.3, deform_cp=0.2)
params = dict(rmode=1, rexp=pb.f32(2.0))
ctx = pb.gpu_context(P, params=params, cliques="pairs")
cf = pb.oracle_costfunction(P, params=params, cliques="pairs")
g = ctx.pair_table().astype(np.float64)
o = cf.pair_table()
e = pb.rel_err(g, o, 1e-4)
idx = np.argsort(e.ravel())[::-1][:10]
for i in idx:
    p, lb, la = np.unravel_index(i, e.shape)
    a, b = P.pairs[p]
    R1 = pyorc.rotation_matrix(P.cxyz[a], P.rot[a] @ P.labels[la])
    R2 = pyorc.rotation_matrix(P.cxyz[b], P.rot[b] @ P.labels[lb])
    th = np.arccos((np.trace(R1.T @ R2) - 1) / 2)
    K1, K2 = R1 - np.eye(3), R2 - np.eye(3)
    x = -0.5 * (np.trace(K1) + np.trace(K2) + (K1 * K2).sum())
    th2 = 2 * np.arcsin(np.sqrt(x / 2))
    print(p, la, lb, "gpu", g.ravel()[i], "orc", o.ravel()[i], "rel", e.ravel()[i], "theta acos", th, "theta asin", th2, "ratio^2", (th2 / th) ** 2 - 1,
          "g/o-1", g.ravel()[i] / o.ravel()[i] - 1)

# triplet rmode 2 variant
params = dict(rmode=2, dweight=True, anorm=True, rexp=pb.f32(3.0))
P = pb.make_problem(4, 3, 1, seed=9, warp=0.3, deform_cp=0.25)
ctx = pb.gpu_context(P, params=params)
cf = pb.oracle_costfunction(P, params=params)
rng = np.random.default_rng(0)
n = 30000
q = np.stack([rng.integers(0, P.T, n), rng.integers(0, P.L, n), rng.integers(0, P.L, n), rng.integers(0, P.L, n)], 1).astype(np.int32)
g, o = ctx.triplet_costs(q), cf.triplet_costs(q)
e = pb.rel_err(g, o, 1e-4)
for i in np.argsort(e)[::-1][:8]:
    print(q[i], "gpu", g[i], "orc", o[i], "rel", e[i])
