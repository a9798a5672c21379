"""Generates tests/golden/small_tables.npz from THE REFERENCE ITSELF: oracle/_ref/libref_cost.so = the reference's own
src/DiscreteCostFunction.cpp / similarities.cc / reg_tools.cpp / DiscreteModel.cpp compiled unmodified where they lie
(oracle/Makefile `refcost`; the absent FSL libraries stood in for by oracle/ext/fsl_ext.h), run single-threaded (its
multi-threaded unary loop is racy, see tests/test_reference_pin.py).

The committed fixture is what travels: the CPU test checks the oracle port against it, the GPU test the CUDA path.
Re-run (dev container only, needs /root/reference at build time):  python tests/golden/make_golden.py
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


def queries(P):
    rng = np.random.default_rng(99)
    return np.stack([rng.integers(0, P.T, 400), rng.integers(0, P.L, 400), rng.integers(0, P.L, 400), rng.integers(0, P.L, 400)], 1).astype(np.int32)


def tables(name):
    """The ORACLE's version of one case (what test_host_cpu checks against the committed reference output)."""
    kw, params = CASES[name]
    P = pb.make_problem(**kw)
    cf = pb.oracle_costfunction(P, params=params)
    cf.get_source_data()
    out = {}
    patches = cf.patches()
    out["patch_sizes"] = np.array([len(p) for p in patches], np.int32)
    out["patch_idx"] = np.concatenate(patches).astype(np.int32)
    out["unary"] = cf.unary_costs()
    out["triplet_q"] = queries(P)
    out["triplet_costs"] = cf.triplet_costs(out["triplet_q"])
    cfp = pb.oracle_costfunction(P, params=dict(params, rmode=1), cliques="pairs")
    out["pair_table_first8"] = cfp.pair_table()[:8]
    return P, out


def reference_tables(name):
    """The same case through the reference's own classes (src/mesh_registration.cpp:66,86,336-337,347 call order)."""
    from oracle import pyrefcost
    kw, params = CASES[name]
    P = pb.make_problem(**kw)

    def model(**over):
        pr = dict(pb.DEFAULT_PARAMS); pr.update(params); pr.update(over)
        m = pyrefcost.model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=P.C > 1, threads=1, **pr)
        m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref)
        m.initialize()
        m.reset_meshspace(P.sxyz)
        m.reset_cpgrid(P.cxyz)
        m.setup()
        return m

    m = model()
    assert np.array_equal(m.labels(), P.labels) and np.array_equal(m.triplets(), P.triplets)
    out = {}
    patches = pyrefcost.patches(m)
    out["patch_sizes"] = np.array([len(p) for p in patches], np.int32)
    out["patch_idx"] = np.concatenate(patches).astype(np.int32)
    out["unary"] = m.unary()
    out["triplet_q"] = queries(P)
    out["triplet_costs"] = m.costs("triplet", out["triplet_q"])
    m.close()
    mp = model(rmode=1, dopt="FastPD")
    assert np.array_equal(mp.pairs(), P.pairs)
    L = mp.L
    q = np.array([[p, la, lb] for p in range(8) for lb in range(L) for la in range(L)], np.int32)   # [P][lb][la] (:303)
    out["pair_table_first8"] = mp.costs("pair", q).reshape(8, L, L)
    mp.close()
    return P, out


if __name__ == "__main__":
    blob = {}
    for name in CASES:
        _, out = reference_tables(name)
        for k, v in out.items():
            blob[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "small_tables.npz"), **blob)
    print("wrote", os.path.join(HERE, "small_tables.npz"), {k: v.shape for k, v in blob.items()})
