"""Scratch GPU debugging (not a pytest file)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import problems as pb  # noqa: E402
from oracle import pyorc, pyref  # noqa: E402
from msm_newmeshreg_b200 import host  # noqa: E402

P = pb.make_problem(4, 2, 1, seed=34, warp=0.3, deform_cp=0.1)
pr = dict(pb.DEFAULT_PARAMS); pr.update(rmode=3, dopt="HOCR")
m = host.Model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=False, _lib_override=pyref.lib(), **pr)
m.set_data(P.txyz, P.ttris, P.sxyz, P.ttris, P.inp, P.ref)
m.initialize()
print("labels equal", np.array_equal(m.labels() if m.L else None, P.barycentres) if m.L else "L=0 before setup")
m.reset_cpgrid(P.cxyz)
m.setup()
print("labels equal", np.array_equal(m.labels(), P.barycentres), np.abs(m.labels() - P.barycentres).max())
print("triplets equal", np.array_equal(m.triplets(), P.triplets), "cpgrid equal", np.array_equal(m.cpgrid(), P.cxyz), np.abs(m.cpgrid() - P.cxyz).max())
cf = pb.oracle_costfunction(P, params=dict(rmode=3)); cf.get_source_data()
Uo = cf.unary_costs(); Ug = m.unary()
print("unary max abs diff", np.abs(Ug - Uo).max())
TT = cf.triplet_table()
L = m.L
q = np.array([[t, a, b, c] for t in range(0, P.T, 7) for a in range(0, L, 3) for b in range(0, L, 4) for c in range(0, L, 5)], np.int32)
tg = m.costs("triplet", q)
to = TT[q[:, 0], q[:, 1], q[:, 2], q[:, 3]]
e = pb.rel_err(tg, to, 1e-4)
print("triplet lookups max rel", e.max(), "max abs", np.abs(tg - to).max())
lab_orc, e_orc = pyref.fusion_tables(Uo, P.triplets, TT, threads=4)
lab_orc1, e_orc1 = pyref.fusion_tables(Uo, P.triplets, TT, threads=1)
print("oracle fusion threads 4 vs 1 labels differ:", (lab_orc != lab_orc1).sum(), e_orc, e_orc1)
# GPU tables through the table path
ctx = m.context()
TTg = ctx.triplet_table().astype(np.float64)
lab_gt, e_gt = pyref.fusion_tables(Ug, P.triplets, TTg, threads=1)
print("gpu tables (table path) vs oracle labels differ:", (lab_gt != lab_orc1).sum(), e_gt)
TTo32 = TT.astype(np.float32).astype(np.float64)
lab_o32, e_o32 = pyref.fusion_tables(Uo, P.triplets, TTo32, threads=1)
print("oracle tables rounded to f32 vs oracle labels differ:", (lab_o32 != lab_orc1).sum(), e_o32)
lab_gpu, e_gpu = pyref.fusion_model(m, threads=1)
print("gpu model path vs oracle labels differ:", (lab_gpu != lab_orc1).sum(), e_gpu, " vs gpu table path:", (lab_gpu != lab_gt).sum())
