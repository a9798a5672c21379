"""Timing of the two fusion drivers on the GPU-backed model (not a pytest file): the reference's unmodified Fusion::optimize
(per-element virtuals served from a T*L^3 look-up table) against host/src/FusionTables.h (batched move tables).
  python tests/fusion_driver_timing.py"""
import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import problems as pb
from oracle import pyref
from msm_newmeshreg_b200 import host
for (dl, cl) in ((6, 4), (7, 5)):
    P = pb.make_problem(dl, cl, 1, seed=91, warp=0.3, deform_cp=0.1)
    res = []
    for name, driver in (("Fusion::optimize (virtuals + T*L^3 table)", pyref.fusion_model), ("FusionTables (batched)", pyref.fusion_tables_model)):
        pr = dict(pb.DEFAULT_PARAMS); pr.update(rmode=3, dopt="HOCR")
        m = host.Model(sgres=cl + 2, cpres=cl, multivariate=False, _lib_override=pyref.lib(), **pr)
        m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref); m.initialize(); m.reset_meshspace(P.sxyz); m.reset_cpgrid(P.cxyz); m.setup()
        t = time.perf_counter(); lab, e = driver(m, threads=16); dt = time.perf_counter() - t
        res.append((lab.copy(), e)); print(f"ico{dl}/ico{cl} {name}: {dt:.3f} s, energy {e:.6f}", flush=True)
        m.close()
    print("  labellings identical:", np.array_equal(res[0][0], res[1][0]), "differ:", int((res[0][0] != res[1][0]).sum()))
