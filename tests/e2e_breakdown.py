"""Per-call breakdown of the end-to-end step at C3 (ico7 data, ico5 control grid) through the host classes; with
MSM_HOST_TIMING=1 the host library prints its own breakdown.  Not a pytest file:  python tests/e2e_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import problems as pb
from msm_newmeshreg_b200 import host
from oracle import pyorc

txyz, ttris = pyorc.icosphere(7)
cxyz0, ctris = pyorc.icosphere(5)
mvd = float(np.linalg.norm(cxyz0[ctris[:, 0]] - cxyz0[ctris[:, 1]], axis=1).mean())
inp = pb.random_field(txyz, 3)[None]; ref = pb.random_field(txyz, 4)[None]
sx = pb.smooth_warp(txyz, 5, 0.3 * mvd); cx = pb.smooth_warp(cxyz0, 6, 0.1 * mvd)
m = host.Model(sgres=7, cpres=5, multivariate=False, dopt="HOCR", threads=1, **pb.DEFAULT_PARAMS)
m.set_data(txyz, ttris, txyz, ttris, inp, ref); m.initialize(); m.reset_meshspace(sx); m.reset_cpgrid(cx); m.setup()
c = m.context()
lab = np.random.default_rng(7).integers(0, m.L, m.N).astype(np.int32)
import torch
moves = torch.empty((m.L, m.T, 8), dtype=torch.float64).pin_memory().numpy()
T = {k: [] for k in ("reset_meshspace", "reset_cpgrid", "setup", "moves")}
fam = {}
for it in range(8):
    c.set_timing(1)
    t0 = time.perf_counter(); m.reset_meshspace(sx); t1 = time.perf_counter(); m.reset_cpgrid(cx); t2 = time.perf_counter()
    m.setup(); t3 = time.perf_counter(); m.triplet_moves(lab, -1, out=moves); t4 = time.perf_counter()
    for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)): T[k].append(v * 1e3)
    for f in ("patches", "unary", "label_rot", "triplet_moves"):
        fam.setdefault(f, []).append(c.last_kernel_ms(f))
    c.set_timing(0)
for k, v in T.items(): print(k, "ms: min %.3f med %.3f" % (min(v), sorted(v)[len(v) // 2]))
for k, v in fam.items(): print("  family", k, "ms: min %.3f" % min(v))
# without timing mode (no extra syncs)
ts = []
for it in range(8):
    t0 = time.perf_counter(); m.reset_meshspace(sx); m.reset_cpgrid(cx); m.setup(); m.triplet_moves(lab, -1, out=moves); ts.append((time.perf_counter() - t0) * 1e3)
print("whole e2e step per subject ms: min %.3f med %.3f" % (min(ts), sorted(ts)[4]))
