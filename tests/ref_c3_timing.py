"""One-off timing of THE REFERENCE ITSELF (oracle/_ref/libref_cost.so, single host thread) on the C3 workload of bench.py
(ico7 data x ico5 control grid, univariate): setupCostFunction (brute-force patch search) + computeUnaryCosts +
the fusion-move triplet costs of a few proposal labels.  Not a pytest file; CPU only.

  python tests/ref_c3_timing.py [--data 7] [--cp 5] [--alphas 2] [--out profiles/x.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import problems as pb  # noqa: E402
from oracle import pyorc, pyrefcost  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--data", type=int, default=7)
ap.add_argument("--cp", type=int, default=5)
ap.add_argument("--alphas", type=int, default=2)
ap.add_argument("--out", default=None)
a = ap.parse_args()

txyz, ttris = pyorc.icosphere(a.data)
cxyz0, ctris = pyorc.icosphere(a.cp)
mvd = float(np.linalg.norm(cxyz0[ctris[:, 0]] - cxyz0[ctris[:, 1]], axis=1).mean())
inp = pb.random_field(txyz, 3)[None]; ref = pb.random_field(txyz, 4)[None]
sx = pb.smooth_warp(txyz, 5, 0.3 * mvd); cx = pb.smooth_warp(cxyz0, 6, 0.1 * mvd)
m = pyrefcost.model(sgres=a.cp + 2, cpres=a.cp, multivariate=False, dopt="HOCR", threads=1, **pb.DEFAULT_PARAMS)
m.set_data(txyz, ttris, txyz, ttris, inp, ref)
t0 = time.perf_counter(); m.initialize(); t_init = time.perf_counter() - t0
m.reset_meshspace(sx); m.reset_cpgrid(cx)
t0 = time.perf_counter(); m.setup(); t_setup = time.perf_counter() - t0
t0 = time.perf_counter(); U = m.unary(); t_unary = time.perf_counter() - t0
lab = np.random.default_rng(7).integers(0, m.L, m.N).astype(np.int32)
t0 = time.perf_counter()
for al in range(a.alphas):
    m.triplet_moves(lab, al)
t_moves = (time.perf_counter() - t0) / max(a.alphas, 1)       # per proposal label, all host threads (re-entrant virtual)
evals = m.N * m.L + m.L * m.T * 8
total = t_setup + t_unary + t_moves * m.L
res = {"workload": f"ico{a.data} data x ico{a.cp} control grid, univariate, L={m.L}, N={m.N}, T={m.T}",
       "arm": "reference: its own classes compiled in place, cost functions on 1 host thread",
       "initialize_s": t_init, "setupCostFunction_s": t_setup, "computeUnaryCosts_s": t_unary,
       "triplet_moves_per_label_s": t_moves, "step_s_extrapolated": total, "evals_per_step": evals,
       "evals_per_s": evals / total, "unary_evals_per_s": m.N * m.L / t_unary, "host_threads_for_moves": os.cpu_count()}
print(json.dumps(res, indent=1))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
m.close()
