"""Pins the ORACLE port (oracle/oracle.cpp) against the REFERENCE ITSELF run here: oracle/_ref/libref_cost.so holds the
reference's own src/DiscreteCostFunction.cpp, src/similarities.cc, src/reg_tools.cpp and src/DiscreteModel.cpp compiled
unmodified where they lie (oracle/Makefile `refcost`), with only the absent third-party FSL libraries stood in for
(oracle/ext/fsl_ext.h, whose geometry forwards to oracle/geom.h).  CPU only.

The reference's multi-threaded unary loop is racy (src/DiscreteCostFunction.cpp:306-313 runs computeUnaryCost under
`omp parallel for` while `sim` keeps the patch means in members, src/similarities.h:38, and the multivariate
`_targetdata` is indexed by source vertex and shared between overlapping patches, :484-505); its single-threaded
result is the deterministic semantics everything here is pinned to.
"""
import numpy as np
import pytest

import problems as pb
from oracle import pyorc, pyrefcost

pytestmark = pytest.mark.skipif(not pyrefcost.available(),
                                reason="oracle/_ref/libref_cost.so not built (needs /root/reference at build time)")


def ref_model(P, **kw):
    pr = dict(pb.DEFAULT_PARAMS); pr.update(kw)
    pr.setdefault("threads", 1)
    m = pyrefcost.model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=P.C > 1, **pr)
    m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref)   # src/mesh_registration.cpp:66
    m.initialize()                                               # :86
    m.reset_meshspace(P.sxyz)                                    # :336
    return m


def test_model_glue():
    """Icosphere, control grid, label sets of both parities, cliques, spacings — src/DiscreteModel.cpp:38-162,245-294."""
    P = pb.make_problem(4, 2, 1, seed=31, warp=0.3)
    m = ref_model(P, rmode=3)
    m.setup()
    assert (m.N, m.L, m.T) == (P.N, len(P.barycentres), P.T)
    assert np.array_equal(m.labels(), P.barycentres)            # iteration 1: barycentres (:215-220)
    assert np.array_equal(m.triplets(), P.triplets)
    assert np.abs(m.cpgrid() - P.cxyz0).max() < 1e-12
    m.setup()
    assert np.array_equal(m.labels(), P.samples)                # iteration 2: vertex samples
    mp = ref_model(P, rmode=1, dopt="FastPD"); mp.setup()
    assert np.array_equal(mp.pairs(), P.pairs)
    for lvl in (0, 1, 3):
        rx, rt = pyrefcost.icosphere(lvl); ox, ot = pyorc.icosphere(lvl)
        assert np.array_equal(rt, ot) and np.abs(rx - ox).max() < 1e-12
    m.close(); mp.close()


@pytest.mark.parametrize("C,deform", [(1, 0.0), (1, 0.15), (3, 0.1)])
def test_unary_table(C, deform):
    """computeUnaryCosts (src/DiscreteCostFunction.cpp:306-313) -> get_source_data / get_target_data / corr."""
    P = pb.make_problem(4, 2, C, seed=41 + C, warp=0.3, deform_cp=deform)
    m = ref_model(P, rmode=3)
    m.reset_cpgrid(P.cxyz)
    m.setup()
    U = m.unary()
    cf = pb.oracle_costfunction(P); cf.get_source_data()
    good = np.ones(P.N, bool); good[pb.singular_nodes(P)] = False
    assert np.abs(U - cf.unary_costs())[:, good].max() < 1e-12
    q = np.array([[5, 3], [17, 0], [100, 18]], np.int32)
    assert np.allclose(m.costs("unary", q), U[q[:, 1], q[:, 0]], atol=0, rtol=1e-15)
    m.close()


@pytest.mark.parametrize("rmode,k_exp", [(3, 2.0), (3, 1.5), (2, 2.0)])
def test_triplet_costs(rmode, k_exp):
    """computeTripletCost (src/DiscreteCostFunction.cpp:138-254) + calculate_triangular_strain / triangle_strain
    (src/reg_tools.cpp:556-760)."""
    P = pb.make_problem(3, 2, 1, seed=51, warp=0.2, deform_cp=0.15)
    kw = dict(rmode=rmode, k_exp=pb.f32(k_exp))
    m = ref_model(P, **kw)
    m.reset_cpgrid(P.cxyz)
    m.setup()
    cf = pb.oracle_costfunction(P, params=kw); cf.get_source_data()
    rng = np.random.default_rng(rmode)
    q = np.stack([rng.integers(0, P.T, 4000)] + [rng.integers(0, P.L, 4000) for _ in range(3)], 1).astype(np.int32)
    r, o = m.costs("triplet", q), cf.triplet_costs(q)
    sing = np.isin(P.triplets[q[:, 0]], pb.singular_nodes(P)).any(1)
    assert pb.rel_err(r[~sing], o[~sing], 1e-9).max() < 1e-9
    lab = rng.integers(0, P.L, P.N).astype(np.int32)             # the 8 fusion-move costs of every triplet
    mv = m.triplet_moves(lab, 3)
    q8 = []
    for t in range(P.T):
        cur = lab[P.triplets[t]]
        for c in range(8):
            q8.append([t, 3 if c & 4 else cur[0], 3 if c & 2 else cur[1], 3 if c & 1 else cur[2]])
    assert np.array_equal(mv.ravel(), m.costs("triplet", np.array(q8, np.int32)))
    m.close()


@pytest.mark.parametrize("kw", [
    dict(simmeasure=1), dict(simmeasure=3),                                  # sign convention (:444-447)
    dict(rng=pb.f32(0.6)), dict(rng=pb.f32(1.4)),                            # within_controlpt_range (:104-109)
    dict(rexp=pb.f32(1.0)), dict(lam=pb.f32(0.0)), dict(mu=pb.f32(0.1), kappa=pb.f32(10.0)),
    dict(rmode=2, dweight=True), dict(rmode=2, anorm=True), dict(rmode=2, dweight=True, anorm=True, rexp=pb.f32(1.0)),
], ids=lambda kw: ",".join(f"{k}={float(v):g}" for k, v in kw.items()))
def test_parameter_sweep(kw):
    """The remaining entries of set_parameters (src/DiscreteCostFunction.cpp:121-136), one at a time, on the unary table,
    the patch lists and a sample of triplet costs."""
    P = pb.make_problem(3, 2, 1, seed=61, warp=0.25, deform_cp=0.12)
    m = ref_model(P, **kw)
    m.reset_cpgrid(P.cxyz)
    m.setup()
    cf = pb.oracle_costfunction(P, params=kw); cf.get_source_data()
    po, pr = cf.patches(), pyrefcost.patches(m)
    assert all(np.array_equal(a, b) for a, b in zip(po, pr))
    good = np.ones(P.N, bool); good[pb.singular_nodes(P)] = False
    assert np.abs(m.unary() - cf.unary_costs())[:, good].max() < 1e-12
    rng = np.random.default_rng(7)
    q = np.stack([rng.integers(0, P.T, 3000)] + [rng.integers(0, P.L, 3000) for _ in range(3)], 1).astype(np.int32)
    r, o = m.costs("triplet", q), cf.triplet_costs(q)
    sing = np.isin(P.triplets[q[:, 0]], pb.singular_nodes(P)).any(1)
    assert pb.rel_err(r[~sing], o[~sing], 1e-9).max() < 1e-9
    if float(kw.get("lam", 1.0)) == 0.0:
        assert not r.any()
    m.close()


def test_pairwise_costs_and_total():
    """computePairwiseCost (:256-296), evaluateTotalCostSum (:35-70)."""
    P = pb.make_problem(3, 2, 1, seed=52, warp=0.2, deform_cp=0.1)
    m = ref_model(P, rmode=1, dopt="FastPD")
    m.reset_cpgrid(P.cxyz)
    m.setup()
    cf = pb.oracle_costfunction(P, params=dict(rmode=1), cliques="pairs"); cf.get_source_data()
    rng = np.random.default_rng(3)
    q = np.stack([rng.integers(0, len(P.pairs), 3000), rng.integers(0, P.L, 3000), rng.integers(0, P.L, 3000)], 1).astype(np.int32)
    r = m.costs("pair", q)
    o = np.array([cf.pair_cost(int(p), int(a), int(b)) for p, a, b in q])
    assert np.abs(r - o).max() <= 1e-9 * max(1.0, np.abs(o).max())
    m.unary()                                                    # fills the unary look-up table the total reads
    lab = rng.integers(0, P.L, P.N).astype(np.int32)
    m.set_labeling(lab)
    assert abs(m.total() - cf.total(lab)) <= 1e-9 * abs(cf.total(lab))
    m.close()
    mt = ref_model(P, rmode=3); mt.reset_cpgrid(P.cxyz); mt.setup(); mt.unary()
    ct = pb.oracle_costfunction(P); ct.get_source_data()
    mt.set_labeling(lab)
    assert abs(mt.total() - ct.total(lab)) <= 1e-9 * abs(ct.total(lab))
    mt.close()


def test_apply_labeling():
    """applyLabeling (src/DiscreteModel.cpp:238-243): CP_k <- ROT[k] * labels[labeling[k]]."""
    P = pb.make_problem(3, 2, 1, seed=32, warp=0.2)
    m = ref_model(P, rmode=3); m.setup()
    lab = np.random.default_rng(0).integers(0, m.L, m.N).astype(np.int32)
    m.set_labeling(lab)
    new = m.apply_labeling()
    expect = np.einsum("nij,nj->ni", P.rot, P.labels[lab])
    good = np.ones(P.N, bool); good[pb.singular_nodes(P)] = False
    assert np.abs(new - expect)[good].max() < 1e-9
    m.close()


def test_unfold_and_warp():
    """unfold / spatialgradient / check_for_intersections are the reference's own (src/reg_tools.cpp:62-181)."""
    xyz, tris = pyorc.icosphere(3)
    rng = np.random.default_rng(9)
    bad = xyz.copy()
    for v in rng.choice(len(xyz), 6, replace=False):             # push a few vertices across their neighbours
        nb = np.unique(tris[(tris == v).any(1)]); nb = nb[nb != v]
        far = xyz[nb[0]] + 0.6 * (xyz[nb[0]] - xyz[v])
        bad[v] = far * (100.0 / np.linalg.norm(far))
    assert pyorc.count_folded(bad, tris) > 0
    r, (o, _) = pyrefcost.unfold(bad, tris), pyorc.unfold(bad, tris)
    assert np.abs(r - o).max() < 1e-9
    assert pyorc.count_folded(r, tris) == 0
    cx, ct = pyorc.icosphere(2)
    moved = pb.smooth_warp(cx, 4, 2.0)
    up = pb.smooth_warp(xyz, 6, 0.5)
    assert np.abs(pyrefcost.mesh_interpolate(up, cx, ct, moved) - pyorc.mesh_interpolate(up, cx, ct, moved)).max() < 1e-10


def test_group_cost_function():
    """DiscreteGroupCostFunction::computeTripletCost / computePairwiseCost (src/DiscreteGroupCostFunction.cpp:9-54)."""
    P = pb.make_problem(3, 2, 1, seed=11)
    S = 3
    cps = np.stack([pb.smooth_warp(P.cxyz0, 50 + s, 0.25 * P.mvd) for s in range(S)])
    rot = np.concatenate([pyorc.rotations(P.centre, cps[s]) for s in range(S)])
    trip = np.concatenate([P.triplets + s * P.N for s in range(S)])
    lam, mu, kappa = pb.f32(0.2), pb.f32(0.4), pb.f32(1.6)
    go, gr = pyorc.GroupCost(), pyrefcost.GroupCost()
    for g in (go, gr):
        g.set(lam, 2.0, mu, kappa, cps, P.ctris, P.cxyz0, P.labels, rot, trip)
    rng = np.random.default_rng(0)
    n = 5000
    q = np.stack([rng.integers(0, S * P.T, n)] + [rng.integers(0, P.L, n) for _ in range(3)], 1).astype(np.int32)
    r, o = gr.triplet_costs(q), go.triplet_costs(q)
    assert (o == 1e7).any() and np.array_equal(r == 1e7, o == 1e7)
    assert pb.rel_err(r, o, 1e-9).max() < 1e-9
    # pairwise: correlation over the intersection of two sparse patch maps
    N, L, V = 12, 5, 400
    off = [0]; keys = []; vals = []
    for m in range(S * N * L):
        k = np.sort(rng.choice(V // (1 if m % 7 else 8), size=min(int(rng.integers(0, 60)), V // 8), replace=False))
        keys.append(k); vals.append(rng.normal(size=len(k)) + (5.0 if m % 3 == 0 else 0.0))
        off.append(off[-1] + len(k))
    keys = np.concatenate(keys).astype(np.int32); vals = np.concatenate(vals)
    pairs = np.array([[a * N + v, b * N + int(rng.integers(0, N))] for a in range(S) for v in range(N) for b in range(a + 1, S)], np.int32)
    ix, it = pyorc.icosphere(0)
    cps0 = np.tile(ix, (S, 1))
    labels = np.tile(ix[0], (L, 1)) + rng.normal(size=(L, 3)) * 0.01
    go, gr = pyorc.GroupCost(), pyrefcost.GroupCost()
    for g in (go, gr):
        g.set(1.0, 2.0, 0.4, 1.6, cps0.reshape(S, N, 3), it, ix, labels, np.tile(np.eye(3), (S * N, 1, 1)), np.zeros((S * 20, 3), np.int32))
        g.set_patches(off, keys, vals, pairs)
    q = np.stack([rng.integers(0, len(pairs), 3000), rng.integers(0, L, 3000), rng.integers(0, L, 3000)], 1).astype(np.int32)
    r, o = gr.pair_costs(q), go.pair_costs(q)
    assert np.abs(r - o).max() < 1e-14 and (o == 0.5).any()


def test_reference_optimisers_on_reference_costs_equal_oracle_tables():
    """The reference end to end on the CPU (its cost functions AND its optimisers) against its optimisers driven by the
    oracle's tables: identical labellings — FastPD (src/mesh_registration.cpp:346-352) and HOCR fusion (:358)."""
    from oracle import pyref
    P = pb.make_problem(3, 2, 1, seed=33, warp=0.3, deform_cp=0.1)
    m = ref_model(P, rmode=1, dopt="FastPD"); m.reset_cpgrid(P.cxyz); m.setup()
    lab_ref, e_ref = pyrefcost.fastpd_model(m)
    cf = pb.oracle_costfunction(P, params=dict(rmode=1), cliques="pairs"); cf.get_source_data()
    lab_orc, e_orc = pyref.fastpd_tables(cf.unary_costs(), P.pairs, cf.pair_table()) if pyref.available() else (lab_ref, e_ref)
    assert np.array_equal(lab_ref, lab_orc) and abs(e_ref - e_orc) <= 1e-9 * abs(e_orc)
    m.close()
    m = ref_model(P, rmode=3, dopt="HOCR"); m.reset_cpgrid(P.cxyz); m.setup()
    lab_ref, e_ref = pyrefcost.fusion_model(m, threads=1)
    cf = pb.oracle_costfunction(P); cf.get_source_data()
    if pyref.available():
        lab_orc, e_orc = pyref.fusion_tables(cf.unary_costs(), P.triplets, cf.triplet_table(), threads=2)
        assert np.array_equal(lab_ref, lab_orc) and abs(e_ref - e_orc) <= 1e-9 * abs(e_orc)
    m.close()


def test_bench_reference_sample_equals_oracle_sample():
    """bench.py's CPU arm drives the reference's cost functions directly on a bounded sample (host/harness.cpp
    ref_cost_sample): the sample must be the same numbers the oracle port produces for those control points / triplets."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    tabs = {}
    r = bench.reference_sample(4, 2, 1, rows=40, threads=2, tables=tabs)
    txyz, ttris = pyorc.icosphere(4); cxyz0, ctris = pyorc.icosphere(2)
    maxsep, mvdmax, mvd = pyorc.spacings(cxyz0, ctris)
    S = bench.Subject(txyz, ttris, 1, 1000, mvd)
    samples, bary, centre = pyorc.label_grid(4, 0.5 * mvdmax)
    cxyz = pb.smooth_warp(cxyz0, 1200, 0.1 * mvd)
    trip = pyorc.triplets(cxyz0, ctris)
    cf = pyorc.CostFunction(); cf.set_params(threads=2, **bench.PARAMS)
    cf.set_meshes(txyz, ttris, S.sxyz, cxyz0, ctris); cf.reset_cpgrid(cxyz); cf.set_orig(cxyz0)
    cf.set_features(S.ref, S.inp, False); cf.set_spacings(maxsep, mvdmax)
    cf.set_labels(bary, pyorc.rotations(centre, cxyz)); cf.set_triplets(trip)
    cf.get_source_data(0, 40)
    assert np.abs(tabs["unary"] - cf.unary_costs(0, 40)[:, :40]).max() < 1e-12
    lab, L, nt = tabs["labeling"], r["L"], r["ntrip"]
    q = np.array([[t, a if c & 4 else lab[trip[t, 0]], a if c & 2 else lab[trip[t, 1]], a if c & 1 else lab[trip[t, 2]]]
                  for a in range(L) for t in range(nt) for c in range(8)], np.int32)
    assert pb.rel_err(tabs["moves"].ravel(), cf.triplet_costs(q), 1e-9).max() < 1e-9


@pytest.mark.parametrize("C", [1, 3])
def test_resample_weights_oracle_equals_reference_build(C):
    """Non-unit cost-function weighting: AbsoluteWeights = metric_resample([EXT] A8) of the per-vertex maximum weight
    (src/DiscreteCostFunction.cpp:364-383) multiplies every unary cost.  The oracle port's restatement against the
    reference's own get_source_data / computeUnaryCosts."""
    P = pb.make_problem(4, 2, C, seed=31, warp=0.3, deform_cp=0.1)
    rows = 1 if C == 1 else C
    w = 0.25 + np.abs(np.stack([pb.random_field(P.sxyz, 700 + r) for r in range(rows)]))
    m = pyrefcost.model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=C > 1, threads=1, **pb.DEFAULT_PARAMS)
    m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref); m.initialize()
    m.reset_meshspace(P.sxyz); m.reset_cpgrid(P.cxyz); m.set_weights(w); m.setup()
    cf = pb.oracle_costfunction(P, cfweight=w); cf.get_source_data()
    aw = cf.absweights()
    assert aw.std() > 1e-2 and aw.min() > 0.2                     # really resampled, not the unit default
    good = np.ones(P.N, bool); good[pb.singular_nodes(P)] = False
    assert np.abs(cf.unary_costs()[:, good] - m.unary()[:, good]).max() < 1e-12
    m.close()
