"""Known-answer tests that pin the ORACLE (the reference has no tests of its own: SURVEY.md §4, §8c).

Each test names the reference lines whose behaviour it checks.  CPU only.
"""
import numpy as np
import pytest

import problems as pb
from oracle import pyorc

RAD = 100.0


# ---------------------------------------------------------------- A7 icosphere
@pytest.mark.parametrize("level", [0, 1, 2, 3, 4])
def test_icosphere_counts_and_nesting(level):
    xyz, tris = pyorc.icosphere(level)
    assert len(xyz) == 10 * 4 ** level + 2 and len(tris) == 20 * 4 ** level
    assert np.allclose(np.linalg.norm(xyz, axis=1), RAD, atol=1e-10)
    if level > 0:  # nested numbering: coarse vertices keep their ids (src/DiscreteCostFunction.cpp:175-177 relies on it)
        cx, _ = pyorc.icosphere(level - 1)
        assert np.array_equal(xyz[: len(cx)], cx)
    # closed, consistently oriented surface: every face normal points outward
    a, b, c = xyz[tris[:, 0]], xyz[tris[:, 1]], xyz[tris[:, 2]]
    assert (np.einsum("ij,ij->i", np.cross(b - a, c - a), a + b + c) > 0).all()


# ---------------------------------------------------------------- A2 rotation
def test_rotation_matrix_maps_a_to_b_and_is_minimal():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.normal(size=3); b = rng.normal(size=3)
        R = pyorc.rotation_matrix(a, b)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12
        assert np.allclose(R @ (a / np.linalg.norm(a)), b / np.linalg.norm(b), atol=1e-12)
        axis = np.cross(a, b)
        assert np.allclose(R @ axis, axis, atol=1e-10)  # rotation about a x b: the minimal one
    assert np.array_equal(pyorc.rotation_matrix([1, 2, 3], [2, 4, 6]), np.eye(3))  # identity below 1e-8 rad


# ---------------------------------------------------------------- A3/A4 locate (vii)
@pytest.mark.parametrize("level", [2, 4])
def test_grid_locator_equals_brute_force_definition(level):
    xyz, tris = pyorc.icosphere(level)
    rng = np.random.default_rng(level)
    pts = rng.normal(size=(3000, 3)); pts *= RAD / np.linalg.norm(pts, axis=1, keepdims=True)
    tb, wb = pyorc.locate(xyz, tris, pts, brute=True)
    tg, wg = pyorc.locate(xyz, tris, pts)
    assert np.array_equal(tb, tg)
    assert np.allclose(wg.sum(1), 1.0) and (wg > -1e-12).all()
    rec = (wg[:, :, None] * xyz[tris[tg]]).sum(1)       # radial projection: the weights reproduce the direction of p
    rec *= RAD / np.linalg.norm(rec, axis=1, keepdims=True)
    assert np.abs(rec - pts).max() < 1e-9


def test_locator_on_deformed_mesh():
    xyz, tris = pyorc.icosphere(3)
    warped = pb.smooth_warp(xyz, 5, 3.0)
    rng = np.random.default_rng(1)
    pts = rng.normal(size=(1000, 3)); pts *= RAD / np.linalg.norm(pts, axis=1, keepdims=True)
    tb, _ = pyorc.locate(warped, tris, pts, brute=True)
    tg, _ = pyorc.locate(warped, tris, pts)
    assert np.array_equal(tb, tg)


def test_vertex_gets_unit_weight():
    xyz, tris = pyorc.icosphere(3)
    t, w = pyorc.locate(xyz, tris, xyz[:50])
    for i in range(50):
        j = list(tris[t[i]]).index(i)
        assert abs(w[i, j] - 1) < 1e-9


# ---------------------------------------------------------------- similarity (src/similarities.cc:122-173)
def test_corr_matches_numpy_weighted_pearson():
    rng = np.random.default_rng(2)
    a, b, w = rng.normal(size=200), rng.normal(size=200), rng.uniform(0.1, 2, size=200)
    ma, mb = np.average(a, weights=w), np.average(b, weights=w)
    ref = np.sum(w * (a - ma) * (b - mb)) / np.sqrt(np.sum(w * (a - ma) ** 2) * np.sum(w * (b - mb) ** 2))
    assert abs(pyorc.corr(a, b, w) - ref) < 1e-13
    assert abs(pyorc.corr(a, b) - np.corrcoef(a, b)[0, 1]) < 1e-13
    assert pyorc.corr(np.ones(10), b[:10]) == 0.0                      # :171 zero variance -> 0
    assert pyorc.corr(np.zeros(0), np.zeros(0)) == 0.0                 # empty patch


# ---------------------------------------------------------------- strain (iv)  src/reg_tools.cpp:556-601,703-748
def test_strain_closed_forms():
    rng = np.random.default_rng(3)
    T0 = rng.normal(size=(3, 3)) * 5 + np.array([50.0, 20, 70])
    mu, kappa = pb.f32(0.4), pb.f32(1.6)
    assert abs(pyorc.strain(T0, T0, mu, kappa)) < 1e-10                      # identical triangles
    for s in (0.5, 1.3, 2.0):                                                 # uniform scale: R = 1, J = s^2
        for k in (2.0, 1.0, 3.0):
            W = pyorc.strain(T0, s * T0, mu, kappa, k)
            assert abs(W - 0.5 * kappa * (s ** (2 * k) + s ** (-2 * k) - 2)) < 1e-9 * max(1, W)
    # rigid motion leaves the energy at zero
    R = pyorc.rotation_matrix([1, 0, 0], [0.3, 0.5, -0.2])
    assert abs(pyorc.strain(T0, T0 @ R.T + 3.0, mu, kappa)) < 1e-10
    # pure shear of a triangle in the xy-plane: principal stretches (lam, 1/lam) -> J = 1, R = lam^2
    P0 = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0]])
    lam = 1.25
    W = pyorc.strain(P0, P0 * np.array([lam, 1 / lam, 1]), mu, kappa)
    assert abs(W - 0.5 * mu * (lam ** 4 + lam ** -4 - 2)) < 1e-12


def test_strain_literal_route_equals_gram_form():
    rng = np.random.default_rng(4)
    for _ in range(200):
        T0 = rng.normal(size=(3, 3)) * 3 + rng.normal(size=3) * 60
        T1 = T0 + rng.normal(size=(3, 3)) * 0.8
        for k in (2.0, 1.5):
            a = pyorc.strain(T0, T1, 0.4, 1.6, k, which=0)
            b = pyorc.strain(T0, T1, 0.4, 1.6, k, which=1)
            assert abs(a - b) <= 1e-9 * max(1.0, abs(a))


# ---------------------------------------------------------------- model glue (src/DiscreteModel.cpp)
def test_label_grid():
    for cp_level in (2, 3):
        cxyz, ctris = pyorc.icosphere(cp_level)
        _, mvdmax, _ = pyorc.spacings(cxyz, ctris)
        samples, bary, centre = pyorc.label_grid(cp_level + 2, 0.5 * mvdmax)
        assert len(samples) == 19 and len(bary) == 19                 # L = 19 at labeldist 0.5, SG = CP+2
        assert np.array_equal(samples[0], centre) and np.array_equal(bary[0], centre)   # Q2: label 0 = identity
        for lab in (samples, bary):
            assert np.allclose(np.linalg.norm(lab, axis=1), RAD)
            assert (np.linalg.norm(lab[1:] - centre, axis=1) <= 0.5 * mvdmax + 1e-9).all()
            assert len(np.unique(np.round(lab, 6), axis=0)) == len(lab)


def test_cliques_and_spacings():
    cxyz, ctris = pyorc.icosphere(2)
    trip = pyorc.triplets(cxyz, ctris)
    assert trip.shape == (320, 3) and (np.diff(trip, axis=1) > 0).all()            # sorted ids (:267-282)
    assert np.array_equal(np.sort(trip, axis=1), np.sort(ctris, axis=1))
    pairs = pyorc.pairs(cxyz, ctris)
    assert pairs.shape == (480, 2) and (pairs[:, 0] < pairs[:, 1]).all()            # 3(N-2) edges, i<j (:245-265)
    maxsep, mvdmax, mvd = pyorc.spacings(cxyz, ctris)
    chord = np.linalg.norm(cxyz[pairs[:, 0]] - cxyz[pairs[:, 1]], axis=1)
    assert abs(mvdmax - chord.max()) < 1e-12 and abs(mvd - chord.mean()) < 1e-9
    geo = 2 * RAD * np.arcsin(chord / (2 * RAD))
    for k in (0, 5, 100):
        m = (pairs[:, 0] == k) | (pairs[:, 1] == k)
        assert abs(maxsep[k] - geo[m].max()) < 1e-12                                # :53-62


# ---------------------------------------------------------------- unary term (i)-(iii)
def _self_problem(C=1):
    P = pb.make_problem(3, 1, C, seed=7, warp=0.0)
    P.inp = P.ref.copy()  # source field == reference field on the same (undeformed) mesh
    return P


def test_unary_identity_label_self_correlation_is_zero_cost():
    P = _self_problem()
    cf = pb.oracle_costfunction(P)
    cf.get_source_data()
    U = cf.unary_costs()
    assert np.abs(U[0]).max() < 1e-12           # (i) label 0 = identity, src == ref  =>  corr = 1  =>  cost 0
    assert (U[1:] > 1e-6).all()                  # any displacement decorrelates
    assert abs(U[3, 5] - cf.unary_cost(5, 3)) == 0.0


def test_unary_constant_patch_and_sign_and_weights():
    P = _self_problem()
    P.inp = np.ones_like(P.inp)                                  # (ii) constant source patch: variance 0 -> corr 0
    cf = pb.oracle_costfunction(P); cf.get_source_data()
    assert np.allclose(cf.unary_costs(), 0.5)
    cf = pb.oracle_costfunction(P, params=dict(simmeasure=3)); cf.get_source_data()
    assert np.allclose(cf.unary_costs(), -0.5)                   # Q1: sign flips unless simmeasure in {1,2}
    aw = np.linspace(0.5, 2.0, P.N)
    cf = pb.oracle_costfunction(P); cf.set_absweights(aw); cf.get_source_data()
    assert np.allclose(cf.unary_costs(), 0.5 * aw[None, :])      # AbsoluteWeights scale the cost (:441)


def test_multivariate_single_channel_is_half():
    P = pb.make_problem(3, 1, 1, seed=8)
    cf = pb.oracle_costfunction(P, multivariate=True); cf.get_source_data()
    assert np.allclose(cf.unary_costs(), 0.5)                    # (iii) one channel: per-vertex variance is 0


def test_multivariate_self_correlation():
    P = _self_problem(C=4)
    cf = pb.oracle_costfunction(P); cf.get_source_data()
    U = cf.unary_costs()
    assert np.abs(U[0]).max() < 1e-12 and (U[1:] > 0).all()


def test_patch_membership_definition():
    P = pb.make_problem(3, 1, 1, seed=9, warp=0.3)
    cf = pb.oracle_costfunction(P); cf.get_source_data()
    patches = cf.patches()
    for k in (0, 7, 41):
        d = 2 * RAD * np.arcsin(np.linalg.norm(P.sxyz - P.cxyz[k], axis=1) / (2 * RAD))
        assert np.array_equal(patches[k], np.nonzero(d < P.maxsep[k])[0])           # :104-109, ascending order


# ---------------------------------------------------------------- triplet / pairwise (v)-(vi)
def test_triplet_identity_fold_and_threshold():
    P = pb.make_problem(3, 2, 1, seed=10)
    lam = pb.f32(0.1)
    cf = pb.oracle_costfunction(P, params=dict(lam=lam, rmode=3))
    ok = ~np.isin(P.triplets, pb.singular_nodes(P)).any(1)          # the antipode of the label-grid centre is singular
    q = np.array([[t, 0, 0, 0] for t in range(P.T)], np.int32)[ok]
    assert np.array_equal(cf.triplet_costs(q), np.zeros(len(q)))  # (iv)+(vi): undeformed grid, identity labels -> exactly 0
    # (v) a folded proposal costs (1 + 1e6) * lambda * 1^rexp: make triangle 3 very thin in the CURRENT grid so that
    # moving its apex by a label carries it across the opposite edge
    t = int(np.nonzero(ok)[0][3])
    i0, i1, i2 = P.triplets[t]
    cp2 = P.cxyz.copy()
    mid = 0.5 * (cp2[i1] + cp2[i2])
    cp2[i0] = mid + 0.02 * (cp2[i0] - mid)
    cp2[i0] *= RAD / np.linalg.norm(cp2[i0])
    cf.reset_cpgrid(cp2)
    cf.set_labels(P.labels, pyorc.rotations(P.centre, cp2))
    costs = cf.triplet_costs(np.array([[t, a, 0, 0] for a in range(P.L)], np.int32))
    folded = np.isclose(costs, (1 + 1e6) * lam, rtol=1e-12, atol=0)   # cost = 1, weight = 1 + 1e6 (:163-167)
    assert folded.any() and not folded.all() and not folded[0]        # label 0 reproduces the current grid: never a fold
    assert (costs >= 0).all()


def test_group_triplet_rules():
    P = pb.make_problem(3, 2, 1, seed=11, deform_cp=0.2)
    S = 3
    cps = np.stack([pb.smooth_warp(P.cxyz0, 50 + s, 0.2 * P.mvd) for s in range(S)])
    rot = np.concatenate([pyorc.rotations(P.centre, cps[s]) for s in range(S)])
    trip = np.concatenate([P.triplets + s * P.N for s in range(S)])
    lam, mu, kappa = pb.f32(0.2), pb.f32(0.4), pb.f32(1.6)
    gc = pyorc.GroupCost()
    gc.set(lam, 2.0, mu, kappa, cps, P.ctris, P.cxyz0, P.labels, rot, trip)
    rng = np.random.default_rng(0)
    q = np.stack([rng.integers(0, S * P.T, 500), rng.integers(0, P.L, 500), rng.integers(0, P.L, 500), rng.integers(0, P.L, 500)], 1).astype(np.int32)
    g = gc.triplet_costs(q)
    # same numbers from the pairwise-subject cost function configured per subject (k_exp = 2, rexp = 2) except folds
    for s in range(S):
        Ps = pb.make_problem(3, 2, 1, seed=11)
        Ps.cxyz = cps[s]; Ps.rot = pyorc.rotations(P.centre, cps[s])
        cf = pb.oracle_costfunction(Ps, params=dict(lam=lam, rmode=3, k_exp=2.0, rexp=2.0))
        m = q[:, 0] // P.T == s
        qs = q[m].copy(); qs[:, 0] -= s * P.T
        c = cf.triplet_costs(qs)
        fold = np.isclose(c, (1 + 1e6) * lam, rtol=1e-12, atol=0)
        assert np.allclose(g[m][~fold], c[~fold], rtol=1e-12, atol=1e-15)
        assert (g[m][fold] == 1e7).all()                          # Q8: bare 1e7 on folds


def test_pairwise_rules():
    P = pb.make_problem(3, 2, 1, seed=12)
    lam = pb.f32(0.1)
    cf = pb.oracle_costfunction(P, params=dict(lam=lam, rmode=1), cliques="pairs")
    assert cf.pair_cost(0, 0, 0) == 0.0                                             # :272 identical rotations
    theta_mvd = 2 * np.arcsin(P.mvdmax / (2 * RAD))
    a, b = P.pairs[4]
    for la, lb in ((1, 0), (0, 2), (3, 7)):
        R1 = pyorc.rotation_matrix(P.cxyz[a], P.rot[a] @ P.labels[la])
        R2 = pyorc.rotation_matrix(P.cxyz[b], P.rot[b] @ P.labels[lb])
        th = np.arccos((np.trace(R1.T @ R2) - 1) / 2)
        expect = lam * (np.sqrt(2) * th / theta_mvd) ** 2
        got = cf.pair_cost(4, la, lb)
        assert abs(got - expect) < 1e-12 or abs(got - (1 + 1e6) * expect) < 1e-6     # fold weight :279
    tab = cf.pair_table()
    assert tab.shape == (len(P.pairs), P.L, P.L)
    assert tab[4, 7, 3] == cf.pair_cost(4, 3, 7)                                     # :303 layout [p][labelB][labelA]


def test_total_energy_is_sum_of_terms():
    P = pb.make_problem(3, 1, 1, seed=13)
    cf = pb.oracle_costfunction(P, params=dict(rmode=3)); cf.get_source_data()
    rng = np.random.default_rng(0)
    lab = rng.integers(0, P.L, P.N).astype(np.int32)
    U = cf.unary_costs()
    q = np.array([[t, lab[P.triplets[t, 0]], lab[P.triplets[t, 1]], lab[P.triplets[t, 2]]] for t in range(P.T)], np.int32)
    expect = U[lab, np.arange(P.N)].sum() + cf.triplet_costs(q).sum()
    assert abs(cf.total(lab) - expect) < 1e-9 * abs(expect)                          # :41-70


# ------------------------------------------------------------------------------------------------ [EXT] A8 metric_resample
def test_adaptive_barycentric_resampling_known_answers():
    """oracle/geom.h metric_resample (adaptive barycentric, restated from the published ADAP_BARY_AREA algorithm):
    a mesh resampled onto itself is the identity; constants are reproduced exactly on any pair of meshes; a smooth field
    survives a small warp; going from a FINE to a COARSE sphere every coarse vertex gathers many sources (the reverse sets
    win), from coarse to fine exactly the 3 corners of the enclosing triangle (the forward sets win)."""
    x5, t5 = pyorc.icosphere(5)
    x3, t3 = pyorc.icosphere(3)
    f = pb.random_field(x5, 5)
    assert np.abs(pyorc.metric_resample(x5, t5, f, x5, t5)[0] - f).max() < 1e-12
    w5 = pb.smooth_warp(x5, 9, 1.0)
    for (a, ta, b, tb) in ((w5, t5, x5, t5), (x5, t5, x3, t3), (x3, t3, x5, t5)):
        c = pyorc.metric_resample(a, ta, np.full(len(a), 3.25), b, tb)[0]
        assert np.abs(c - 3.25).max() < 1e-12
    g = pyorc.metric_resample(w5, t5, f, x5, t5)[0]
    assert np.abs(g - f).max() < 0.15 and np.corrcoef(g, f)[0, 1] > 0.995
    # fine -> coarse: an impulse at one fine vertex reaches at most the 3 coarse corners around it, and every coarse vertex
    # averages over a whole neighbourhood (smoother than point sampling)
    down = pyorc.metric_resample(x5, t5, f, x3, t3)[0]
    point = f[:len(x3)]                                   # nested numbering: coarse vertices are the first fine vertices
    assert np.abs(down - point).max() > 1e-3              # not plain barycentric interpolation (which would return `point` exactly)
    imp = np.zeros(len(x5)); imp[4000] = 1.0
    hit = np.nonzero(pyorc.metric_resample(x5, t5, imp, x3, t3)[0])[0]
    assert 1 <= len(hit) <= 3
    # coarse -> fine: forward sets — at the coarse vertices themselves the value is the coarse value
    up = pyorc.metric_resample(x3, t3, point, x5, t5)[0]
    assert np.abs(up[:len(x3)] - point).max() < 1e-9
