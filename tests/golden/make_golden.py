"""Generates tests/golden/small_tables.npz from the ORACLE (oracle/liborc.so).

The reference itself cannot run here (FSL is absent: SURVEY.md §0.2) and ships no golden vectors, so these fixtures
pin the oracle against regressions and give the GPU tests a second, frozen comparison point; they do NOT pin the
oracle to the real reference (PARITY UNPINNED).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import problems as pb  # noqa: E402

CASES = {
    # name: (make_problem kwargs, params)
    "uni_ico4_ico2": (dict(data_level=4, cp_level=2, C=1, seed=21, warp=0.3, deform_cp=0.15), dict(rmode=3)),
    "multi_ico4_ico2": (dict(data_level=4, cp_level=2, C=4, seed=22, warp=0.3, deform_cp=0.15), dict(rmode=3)),
    "uni_ties_ico3_ico1": (dict(data_level=3, cp_level=1, C=1, seed=23, warp=0.0), dict(rmode=2)),
}


def tables(name):
    kw, params = CASES[name]
    P = pb.make_problem(**kw)
    cf = pb.oracle_costfunction(P, params=params)
    cf.get_source_data()
    out = {}
    patches = cf.patches()
    out["patch_sizes"] = np.array([len(p) for p in patches], np.int32)
    out["patch_idx"] = np.concatenate(patches).astype(np.int32)
    out["unary"] = cf.unary_costs()
    rng = np.random.default_rng(99)
    q = np.stack([rng.integers(0, P.T, 400), rng.integers(0, P.L, 400), rng.integers(0, P.L, 400), rng.integers(0, P.L, 400)], 1).astype(np.int32)
    out["triplet_q"] = q
    out["triplet_costs"] = cf.triplet_costs(q)
    cfp = pb.oracle_costfunction(P, params=dict(params, rmode=1), cliques="pairs")
    out["pair_table_first8"] = cfp.pair_table()[:8]
    return P, out


if __name__ == "__main__":
    blob = {}
    for name in CASES:
        _, out = tables(name)
        for k, v in out.items():
            blob[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "small_tables.npz"), **blob)
    print("wrote", os.path.join(HERE, "small_tables.npz"), {k: v.shape for k, v in blob.items()})
