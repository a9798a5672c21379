"""Scratch GPU debugging / timing breakdown of the end-to-end host path (not a pytest file)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from msm_newmeshreg_b200 import host  # noqa: E402

txyz, ttris = host.icosphere(7)
cxyz0, ctris = host.icosphere(5)
mvd = float(np.linalg.norm(cxyz0[ctris[:, 0]] - cxyz0[ctris[:, 1]], axis=1).mean())
S = bench.Subject(txyz, ttris, 1, 1000, mvd)
m = host.Model(sgres=7, cpres=5, multivariate=False, dopt="HOCR", threads=1, **bench.PARAMS)
m.set_data(txyz, ttris, txyz, ttris, S.inp, S.ref)
m.initialize()
m.reset_meshspace(S.sxyz)
cxyz = bench.smooth_warp(cxyz0, 1200, 0.1 * mvd)
m.reset_cpgrid(cxyz)
m.setup()
ctx = m.context()
ctx.set_timing(1)
lab = np.random.default_rng(7).integers(0, m.L, m.N).astype(np.int32)
out = np.empty((m.L, m.T, 8))
for rep in range(3):
    t0 = time.perf_counter(); m.reset_meshspace(S.sxyz); t1 = time.perf_counter()
    m.reset_cpgrid(cxyz); t2 = time.perf_counter()
    m.setup(); t3 = time.perf_counter()
    m.triplet_moves(lab, -1, out=out); t4 = time.perf_counter()
    print(f"rep {rep}: reset_meshspace {1e3 * (t1 - t0):.2f} ms, reset_cpgrid {1e3 * (t2 - t1):.2f}, setup {1e3 * (t3 - t2):.2f} "
          f"(patches family {ctx.last_kernel_ms('patches'):.2f}, label_rot {ctx.last_kernel_ms('label_rot'):.3f}, unary {ctx.last_kernel_ms('unary'):.3f}), "
          f"triplet_moves {1e3 * (t4 - t3):.2f} (kernel {ctx.last_kernel_ms('triplet_moves'):.3f})")
