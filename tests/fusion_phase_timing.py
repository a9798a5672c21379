"""Where the host optimiser phase goes at HCP scale (not a pytest file): the table-aware fusion driver
(host/src/FusionTables.h) with MSM_FUSION_TIMING=1 prints seconds per phase — unary terms, move tables, AddHigherTerm,
toQuadratic, convert/initialise, FastPD constructor, FastPD::run — summed over the 2 L proposal steps of one outer iteration.
  MSM_FUSION_TIMING=1 python tests/fusion_phase_timing.py [data_level cp_level]"""
import os, sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import problems as pb
from oracle import pyref
from msm_newmeshreg_b200 import host
os.environ.setdefault("MSM_FUSION_TIMING", "1")
dl = int(sys.argv[1]) if len(sys.argv) > 1 else 7
cl = int(sys.argv[2]) if len(sys.argv) > 2 else 5
P = pb.make_problem(dl, cl, 1, seed=91, warp=0.3, deform_cp=0.1)
pr = dict(pb.DEFAULT_PARAMS); pr.update(rmode=3, dopt="HOCR")
m = host.Model(sgres=cl + 2, cpres=cl, multivariate=False, _lib_override=pyref.lib(), **pr)
m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref); m.initialize(); m.reset_meshspace(P.sxyz); m.reset_cpgrid(P.cxyz); m.setup()
t = time.perf_counter(); lab, e = pyref.fusion_tables_model(m, threads=os.cpu_count() or 1); dt = time.perf_counter() - t
print(f"ico{dl}/ico{cl} FusionTables::optimize: {dt:.3f} s, energy {e:.6f}, labels changed {int((lab != 0).sum())} of {m.N}")
m.close()
