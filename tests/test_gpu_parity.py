"""GPU parity tests: the CUDA path, called through the C ABI (include/msm_b200.h), against the oracle on the same
seeded inputs, against the committed golden fixtures, and — at BASELINE.json's full sizes — through size-independent
properties.

Tolerance (north_star: cost tables within 1e-5 relative, fp32):  |gpu - ref| <= 1e-5 * max(|ref|, floor) with
floor = 1e-2 for the unary term (costs live in [0,1]; SURVEY.md §7 "hard parts") and 1e-3 * lambda-scale for the
regularisers.  Patch membership and point location are index work: bit-exact.
"""
import os

import numpy as np
import pytest

import problems as pb
from oracle import pyorc

pytestmark = pytest.mark.gpu

TOL = 1e-5
UFLOOR = 1e-2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(g, o, floor, what, tol=TOL):
    e = pb.rel_err(g, o, floor)
    assert e.max() <= tol, f"{what}: max rel err {e.max():.3e} at {np.unravel_index(e.argmax(), e.shape)} (ref {np.asarray(o).flat[e.argmax()]:.6e})"


def _good_nodes(P):
    m = np.ones(P.N, bool); m[pb.singular_nodes(P)] = False
    return m


# ------------------------------------------------------------------------------------------------ patches
@pytest.mark.parametrize("kw", [dict(data_level=4, cp_level=2, warp=0.3), dict(data_level=4, cp_level=2, warp=0.0),
                                dict(data_level=5, cp_level=3, warp=0.3, deform_cp=0.2), dict(data_level=3, cp_level=0, warp=0.0)])
def test_patch_membership_bit_exact(built, kw):
    """get_source_data: identical lists, including the exact ties of an undeformed icosphere (warp=0)."""
    P = pb.make_problem(C=1, seed=5, **kw)
    ctx = pb.gpu_context(P); ctx.build_patches()
    off, idx = ctx.patches()
    cf = pb.oracle_costfunction(P); cf.get_source_data()
    po = cf.patches()
    assert off[-1] == sum(len(p) for p in po)
    for k in range(P.N):
        assert np.array_equal(po[k], idx[off[k]:off[k + 1]]), k
    ctx.close()


def test_patch_range_parameter_and_empty_patches(built):
    P = pb.make_problem(4, 2, 1, seed=6, warp=0.3)
    for rng_ in (pb.f32(0.7), pb.f32(1.5), pb.f32(0.01)):     # 0.01: every patch holds at most the CP's own vertex
        ctx = pb.gpu_context(P, params=dict(rng=rng_)); ctx.build_patches()
        off, idx = ctx.patches()
        cf = pb.oracle_costfunction(P, params=dict(rng=rng_)); cf.get_source_data()
        po = cf.patches()
        assert all(np.array_equal(po[k], idx[off[k]:off[k + 1]]) for k in range(P.N))
        U, Uo = ctx.unary_table(), cf.unary_costs()
        _check(U, Uo, UFLOOR, f"unary range={rng_}")
        if rng_ < 0.1:
            assert (np.diff(off) <= 1).all() and np.allclose(U, 0.5)   # empty / single-sample patch: rho = 0 -> 0.5*AW
        ctx.close()


def _set_range(ctx, P, r):
    pr = dict(pb.DEFAULT_PARAMS); pr.update(rng=pb.f32(r))
    ctx.set_params(meanangle=pyorc.mesh_meanangle(P.cxyz0, P.ctris), mvdmax=P.mvdmax, **pr)


@pytest.mark.parametrize("warp", [0.3, 0.0])
def test_patch_rebuild_paths_give_identical_lists(built, warp):
    """From the second build on the count pass stashes its members and the fill pass only sorts them; a patch that outgrew
    that slab sends the build through the two-pass search again.  Every path gives the lists of a fresh context (which the
    tests above hold to the oracle), boundary ties of the undeformed icosphere (warp=0) included."""
    P = pb.make_problem(5, 3, 1, seed=9, warp=warp, deform_cp=0.0 if warp == 0.0 else 0.1)
    Q = pb.make_problem(5, 3, 1, seed=10, warp=0.3, deform_cp=0.1)

    def fresh(sxyz, cxyz, r):
        c = pb.gpu_context(P, params=dict(rng=pb.f32(r)))
        c.reset_source(sxyz); c.reset_cpgrid(cxyz); c.build_patches()
        out = c.patches() + (c.unary_table(),)
        c.close()
        return out

    def same(ctx, ref):
        off, idx = ctx.patches()
        assert np.array_equal(off, ref[0]) and np.array_equal(idx, ref[1])
        assert np.array_equal(ctx.unary_table(), ref[2])

    ctx = pb.gpu_context(P, params=dict(rng=pb.f32(0.7)))
    ref_small = fresh(P.sxyz, P.cxyz, 0.7)
    ctx.build_patches(); same(ctx, ref_small)                 # first build: two passes
    assert ctx.patch_build_stats() == (0, 1)
    ctx.build_patches(); same(ctx, ref_small)                 # second build: one pass
    assert ctx.patch_build_stats() == (1, 1)
    ctx.reset_source(Q.sxyz); ctx.reset_cpgrid(Q.cxyz)        # other geometry, still the slab
    ctx.build_patches(); same(ctx, fresh(Q.sxyz, Q.cxyz, 0.7))
    _set_range(ctx, P, 1.5)                                   # patches 4-5x larger than the slab: falls back
    ctx.build_patches(); same(ctx, fresh(Q.sxyz, Q.cxyz, 1.5))
    ctx.reset_source(P.sxyz); ctx.reset_cpgrid(P.cxyz)        # and the slab again, sized from the large patches
    ctx.build_patches(); same(ctx, fresh(P.sxyz, P.cxyz, 1.5))
    _set_range(ctx, P, 0.01)                                  # (nearly) empty patches through the slab
    ctx.build_patches(); same(ctx, fresh(P.sxyz, P.cxyz, 0.01))
    one, two = ctx.patch_build_stats()
    assert one + two == 6 and two >= 2 and one >= 3           # the range 1.5 build overflowed its rows: exact path
    ctx.close()
    if warp == 0.0:
        # range 1 on the undeformed icosphere: neighbours sit EXACTLY on the patch boundary — the one-pass build sees the
        # ties, gives up, and the exact path resolves them with the reference's own formula, every time
        c1 = pb.gpu_context(P)
        ref1 = fresh(P.sxyz, P.cxyz, 1.0)
        for _ in range(3):
            c1.build_patches(); same(c1, ref1)
        assert c1.patch_build_stats() == (0, 3)
        c1.close()


def test_deferred_patch_build_matches_synchronous_build(built):
    """msm_build_patches_async + device-resident table + msm_patches_resolve: same table bits as the synchronous build, and a
    flagged build (boundary ties of the undeformed icosphere, a row that outgrew its capacity) reports `redone`, after which the
    repeated launch gives the exact path's table.  Host-output entry points resolve on their own."""
    import torch
    P = pb.make_problem(5, 3, 1, seed=9, warp=0.3, deform_cp=0.1)
    Q = pb.make_problem(5, 3, 1, seed=9, warp=0.0, deform_cp=0.0)     # undeformed: exact ties at range 1

    def fresh(sxyz, cxyz, r):
        c = pb.gpu_context(P, params=dict(rng=pb.f32(r)))
        c.reset_source(sxyz); c.reset_cpgrid(cxyz); c.build_patches()
        out = c.patches() + (c.unary_table(),)
        c.close()
        return out

    ctx = pb.gpu_context(P)
    d_out = torch.empty((ctx.L, ctx.N), dtype=torch.float32, device="cuda")

    def dev_table():
        ctx.unary_table_dev(d_out.data_ptr()); ctx.synchronize()
        return d_out.cpu().numpy().astype(np.float64)

    ref = fresh(P.sxyz, P.cxyz, 1.0)
    ctx.build_patches_async()                                    # first build of the context: exact path, nothing deferred
    assert not ctx.patches_resolve()
    assert np.array_equal(dev_table(), ref[2])
    ctx.build_patches_async()                                    # one-pass build, status word still in flight
    t = dev_table()
    assert not ctx.patches_resolve()
    assert np.array_equal(t, ref[2]) and ctx.patch_build_stats() == (1, 1)
    off, idx = ctx.patches()
    assert np.array_equal(off, ref[0]) and np.array_equal(idx, ref[1])
    # ties: the one-pass build is flagged, resolve repeats it through the exact path and says so
    ctx.reset_source(Q.sxyz); ctx.reset_cpgrid(Q.cxyz)
    refq = fresh(Q.sxyz, Q.cxyz, 1.0)
    ctx.build_patches_async()
    dev_table()                                                  # computed from a flagged build: not to be trusted
    assert ctx.patches_resolve()
    assert np.array_equal(dev_table(), refq[2])
    off, idx = ctx.patches()
    assert np.array_equal(off, refq[0]) and np.array_equal(idx, refq[1])
    # the host-output table resolves on its own (and recomputes)
    ctx.build_patches_async()
    assert np.array_equal(ctx.unary_table(), refq[2])
    assert not ctx.patches_resolve()
    # a deferred build whose inputs change before anybody looked is simply retired
    ctx.reset_source(P.sxyz); ctx.reset_cpgrid(P.cxyz)
    ctx.build_patches_async()
    ctx.reset_source(Q.sxyz); ctx.reset_cpgrid(Q.cxyz)
    ctx.build_patches_async()
    assert np.array_equal(ctx.unary_table(), refq[2])
    # row overflow: patches 4-5x larger than the rows sized by the previous build
    ctx.reset_source(P.sxyz); ctx.reset_cpgrid(P.cxyz)
    ctx.build_patches()
    _set_range(ctx, P, 1.5)
    ctx.build_patches_async()
    dev_table()
    assert ctx.patches_resolve()
    assert np.array_equal(dev_table(), fresh(P.sxyz, P.cxyz, 1.5)[2])
    ctx.close()


# ------------------------------------------------------------------------------------------------ unary
@pytest.mark.parametrize("kw", [
    dict(data_level=4, cp_level=2, C=1, warp=0.3),
    dict(data_level=4, cp_level=2, C=1, warp=0.0),                      # regular source mesh (iteration 1 of a level)
    dict(data_level=5, cp_level=3, C=1, warp=0.3, deform_cp=0.2, rough=True),
    dict(data_level=5, cp_level=3, C=4, warp=0.3, deform_cp=0.2),
    dict(data_level=5, cp_level=2, C=3, warp=0.2),                      # C not a multiple of 4: padded channels
    dict(data_level=5, cp_level=2, C=7, warp=0.2),
    dict(data_level=4, cp_level=2, C=16, warp=0.2),                     # largest register-resident variant
    dict(data_level=4, cp_level=2, C=21, warp=0.2),                     # generic variant (> 16 channels)
    dict(data_level=6, cp_level=2, C=1, warp=0.3),                      # C1 of BASELINE.json (P ~ 1036: label groups)
    dict(data_level=6, cp_level=4, C=4, warp=0.3, deform_cp=0.2),       # C2 level 3
])
def test_unary_table_parity(built, kw):
    P = pb.make_problem(seed=1, **kw)
    ctx = pb.gpu_context(P); ctx.build_patches()
    U = ctx.unary_table()
    cf = pb.oracle_costfunction(P); cf.get_source_data()
    Uo = cf.unary_costs()
    good = _good_nodes(P)
    _check(U[:, good], Uo[:, good], UFLOOR, f"unary {kw}")
    ctx.close()


def test_unary_samples_label_set_and_sign(built):
    P = pb.make_problem(4, 2, 1, seed=2, warp=0.3, iteration=2)          # even iteration: vertex samples as labels
    for sim in (1, 2, 3):
        ctx = pb.gpu_context(P, params=dict(simmeasure=sim)); ctx.build_patches()
        cf = pb.oracle_costfunction(P, params=dict(simmeasure=sim)); cf.get_source_data()
        good = _good_nodes(P)
        U, Uo = ctx.unary_table(), cf.unary_costs()
        _check(U[:, good], Uo[:, good], UFLOOR, f"unary simmeasure={sim}")
        assert (np.sign(U[:, good]) == (1 if sim in (1, 2) else -1)).all()
        ctx.close()


@pytest.mark.parametrize("C,wrows", [(1, 1), (4, 1), (4, 4)])
def test_unary_costfunction_weights(built, C, wrows):
    P = pb.make_problem(4, 2, C, seed=3, warp=0.3)
    rng = np.random.default_rng(0)
    w = rng.uniform(0.0, 2.0, size=(wrows, len(P.sxyz)))
    w[:, rng.integers(0, len(P.sxyz), 200)] = 0.0                        # masked vertices
    aw = rng.uniform(0.5, 1.5, P.N)
    ctx = pb.gpu_context(P, cfweight=w)
    ctx.set_cpgrid(P.cxyz0, P.ctris, P.maxsep, P.orig, aw); ctx.reset_cpgrid(P.cxyz)
    ctx.build_patches()
    cf = pb.oracle_costfunction(P, cfweight=w); cf.set_absweights(aw); cf.get_source_data()
    good = _good_nodes(P)
    _check(ctx.unary_table()[:, good], cf.unary_costs()[:, good], UFLOOR, "weighted unary")
    ctx.close()


def test_unary_known_answers_on_gpu(built):
    P = pb.make_problem(3, 1, 1, seed=7, warp=0.0)
    P.inp = P.ref.copy()
    ctx = pb.gpu_context(P); ctx.build_patches()
    U = ctx.unary_table()
    assert np.abs(U[0]).max() < 1e-6                                   # identity label, same field: cost 0
    ctx.close()
    P.inp = np.ones_like(P.inp)
    ctx = pb.gpu_context(P); ctx.build_patches()
    assert np.allclose(ctx.unary_table(), 0.5)                          # constant patch: rho = 0
    ctx.close()
    P.inp = np.zeros_like(P.inp); P.ref = np.zeros_like(P.ref)           # medial-wall style all-zero data
    ctx = pb.gpu_context(P); ctx.build_patches()
    assert np.allclose(ctx.unary_table(), 0.5)
    ctx.close()
    P1 = pb.make_problem(3, 1, 1, seed=8)
    ctx = pb.gpu_context(P1, multivariate=True); ctx.build_patches()
    assert np.allclose(ctx.unary_table(), 0.5)                          # multivariate with one channel
    ctx.close()


def test_target_numbering_and_orientation_independent(built):
    """Any vertex/face numbering and any rotation of the target icosphere gives the same table."""
    P = pb.make_problem(4, 2, 1, seed=4, warp=0.3)
    ctx = pb.gpu_context(P); ctx.build_patches(); U0 = ctx.unary_table(); ctx.close()
    rng = np.random.default_rng(1)
    vp = rng.permutation(len(P.txyz)); inv = np.argsort(vp)
    Q = pb.make_problem(4, 2, 1, seed=4, warp=0.3)
    Q.txyz = P.txyz[vp]; Q.ref = P.ref[:, vp]
    Q.ttris = inv[P.ttris][rng.permutation(len(P.ttris))]
    Q.ttris = np.stack([np.roll(t, rng.integers(0, 3)) for t in Q.ttris]).astype(np.int32)
    ctx = pb.gpu_context(Q); ctx.build_patches(); U1 = ctx.unary_table(); ctx.close()
    assert np.abs(U1 - U0).max() < 2e-7
    with pytest.raises(Exception, match="icos|sphere"):
        bad = pb.make_problem(3, 1, 1, seed=4)
        bad.txyz = pb.smooth_warp(bad.txyz, 1, 2.0)
        pb.gpu_context(bad)


def test_contexts_over_differently_oriented_targets_coexist(built):
    """The base-icosahedron frames are per-context state: a second live context over a rotated target sphere does not
    disturb the first one (and vice versa)."""
    P = pb.make_problem(4, 2, 1, seed=4, warp=0.3)
    a = pb.gpu_context(P); a.build_patches(); U0 = a.unary_table()
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]]) @ \
        np.array([[1.0, 0.0, 0.0], [0.0, np.cos(0.4), -np.sin(0.4)], [0.0, np.sin(0.4), np.cos(0.4)]])
    Q = pb.make_problem(4, 2, 1, seed=4, warp=0.3)
    Q.txyz = P.txyz @ R.T
    b = pb.gpu_context(Q); b.build_patches(); Ub = b.unary_table()
    assert np.array_equal(a.unary_table(), U0)
    xyz = P.sxyz[:500]
    ta, wa = a.locate(xyz)
    a.close()
    assert np.array_equal(b.unary_table(), Ub)
    b.close()
    c = pb.gpu_context(Q); c.build_patches()
    assert np.array_equal(c.unary_table(), Ub)
    c.close()
    d = pb.gpu_context(P)
    td, wd = d.locate(xyz)
    assert np.array_equal(ta, td) and np.array_equal(wa, wd)
    d.close()


def test_contexts_driven_from_concurrent_host_threads(built):
    """Independent subjects, one context and one host thread each (bench.py's end-to-end leg): the tables are
    bit-identical to the ones a single thread gets one subject at a time."""
    from concurrent.futures import ThreadPoolExecutor
    probs = [pb.make_problem(5, 3, 1, seed=20 + s, warp=0.3, deform_cp=0.1) for s in range(4)]
    q = np.array([[t, (t * 7) % probs[0].L, (t * 3 + 1) % probs[0].L, (t * 5 + 2) % probs[0].L] for t in range(probs[0].T)], np.int32)

    def run(P, reps):
        ctx = pb.gpu_context(P)
        out = None
        for _ in range(reps):
            ctx.reset_source(P.sxyz); ctx.reset_cpgrid(P.cxyz)
            ctx.build_patches()
            cur = (ctx.unary_table(), ctx.triplet_costs(q), ctx.patches()[1])
            assert out is None or all(np.array_equal(a, b) for a, b in zip(out, cur))
            out = cur
        ctx.close()
        return out

    serial = [run(P, 1) for P in probs]
    with ThreadPoolExecutor(4) as pool:
        threaded = list(pool.map(lambda P: run(P, 5), probs))
    for a, b in zip(serial, threaded):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_shards_equal_single_context(built):
    """CP sharding: rows of the sharded tables are bit-identical to the single-context table."""
    from msm_newmeshreg_b200 import sharding as sh
    P = pb.make_problem(5, 3, 1, seed=5, warp=0.3, deform_cp=0.1)
    ctx = pb.gpu_context(P); ctx.build_patches(); U = ctx.unary_table(); ctx.close()
    world = 3
    out = np.zeros_like(U)
    for r in range(world):
        c = pb.gpu_context(P)
        c.set_shard(*sh.shard_range(P.N, r, world))
        c.build_patches(); c.unary_table(out); c.close()
    assert np.array_equal(out, U)


# ------------------------------------------------------------------------------------------------ point location
@pytest.mark.parametrize("level", [0, 2, 5])
def test_locate_bit_exact_triangles(built, level):
    from msm_newmeshreg_b200 import Context
    xyz, tris = pyorc.icosphere(level)
    ctx = Context(); ctx.set_target(xyz, tris)
    rng = np.random.default_rng(level)
    pts = rng.normal(size=(20000, 3)); pts *= 100.0 / np.linalg.norm(pts, axis=1, keepdims=True)
    pts = np.concatenate([pts, xyz[:min(len(xyz), 500)], 0.5 * (xyz[tris[:200, 0]] + xyz[tris[:200, 1]])])  # vertices, edge points
    pts *= 100.0 / np.linalg.norm(pts, axis=1, keepdims=True)
    tg, wg = ctx.locate(pts)
    to, wo = pyorc.locate(xyz, tris, pts)
    rec = (wg[:, :, None] * xyz[tris[tg]]).sum(1); rec *= 100.0 / np.linalg.norm(rec, axis=1, keepdims=True)
    assert np.abs(rec - pts).max() < 1e-9                               # weights reproduce the point (also on ties)
    interior = wo.min(1) > 1e-9
    assert np.array_equal(tg[interior], to[interior])                   # same triangle away from edges
    assert np.abs(wg[interior] - wo[interior]).max() < 1e-10
    ctx.close()


# ------------------------------------------------------------------------------------------------ triplets
@pytest.mark.parametrize("params", [dict(rmode=3), dict(rmode=3, k_exp=pb.f32(1.5), rexp=pb.f32(1.0)), dict(rmode=2),
                                    dict(rmode=2, dweight=True, anorm=True, rexp=pb.f32(3.0))])
def test_triplet_costs_parity(built, params):
    P = pb.make_problem(4, 3, 1, seed=9, warp=0.3, deform_cp=0.25)
    ctx = pb.gpu_context(P, params=params)
    cf = pb.oracle_costfunction(P, params=params)
    rng = np.random.default_rng(0)
    n = 30000
    q = np.stack([rng.integers(0, P.T, n), rng.integers(0, P.L, n), rng.integers(0, P.L, n), rng.integers(0, P.L, n)], 1).astype(np.int32)
    q[:P.T, 0] = np.arange(P.T); q[:P.T, 1:] = 0
    g, o = ctx.triplet_costs(q), cf.triplet_costs(q)
    _check(g, o, 1e-3 * pb.DEFAULT_PARAMS["lam"], f"triplet {params}", tol=1e-9)     # fp64 kernel
    folds = np.isclose(o, (1 + 1e6) * pb.DEFAULT_PARAMS["lam"], rtol=1e-9)  # anorm is not applied on the fold branch (:163-167)
    assert folds.any()                                                  # the deformed grid exercises the fold branch
    ctx.close()


def test_triplet_table_and_moves(built):
    P = pb.make_problem(4, 2, 1, seed=10, warp=0.3, deform_cp=0.2)
    ctx = pb.gpu_context(P); cf = pb.oracle_costfunction(P)
    tt = ctx.triplet_table(0, 40)
    _check(tt, cf.triplet_table(0, 40), 1e-3 * pb.DEFAULT_PARAMS["lam"], "triplet table", tol=2e-7)   # fp32 storage
    rng = np.random.default_rng(1)
    lab = rng.integers(0, P.L, P.N).astype(np.int32)
    allmv = ctx.triplet_moves(lab, -1)
    for alpha in (0, 5, P.L - 1):
        mv = ctx.triplet_moves(lab, alpha)
        assert np.array_equal(mv, allmv[alpha])
        q = []
        for t in range(P.T):                                            # include/Fusion/Fusion.h:187-194 bit order
            a, b, c = lab[P.triplets[t]]
            for combo in range(8):
                q.append([t, alpha if combo & 4 else a, alpha if combo & 2 else b, alpha if combo & 1 else c])
        o = cf.triplet_costs(np.array(q, np.int32)).reshape(P.T, 8)
        _check(mv, o, 1e-3 * pb.DEFAULT_PARAMS["lam"], f"moves alpha={alpha}", tol=1e-9)
    with pytest.raises(Exception, match="no unary table"):           # the energy is never silently partial
        ctx.total_cost(lab)
    ctx.build_patches(); U = ctx.unary_table(); cf.get_source_data()
    s = ctx.total_cost(lab)
    assert abs(s[2] - cf.total(lab, use_unary=False)) <= 1e-9 * abs(s[2])  # evaluateTotalCostSum, triplet part
    assert abs(s[0] - U[lab, np.arange(P.N)].sum()) <= 1e-6 * abs(s[0])     # unary part = the table at the labelling
    assert abs(sum(s) - cf.total(lab, use_unary=True)) <= 1e-5 * abs(sum(s))
    ctx.set_shard(0, P.N // 2); ctx.build_patches(); ctx.unary_table()
    with pytest.raises(Exception, match="shard"):
        ctx.total_cost(lab)
    ctx.close()


def test_group_triplets(built):
    from msm_newmeshreg_b200 import Context
    P = pb.make_problem(3, 2, 1, seed=11)
    S = 4
    cps = np.stack([pb.smooth_warp(P.cxyz0, 50 + s, 0.25 * P.mvd) for s in range(S)])
    rot = np.concatenate([pyorc.rotations(P.centre, cps[s]) for s in range(S)])
    trip = np.concatenate([P.triplets + s * P.N for s in range(S)])
    lam, mu, kappa = pb.f32(0.2), pb.f32(0.4), pb.f32(1.6)
    gc = pyorc.GroupCost(); gc.set(lam, 2.0, mu, kappa, cps, P.ctris, P.cxyz0, P.labels, rot, trip)
    ctx = Context()
    ctx.set_params(lam=lam, rexp=2.0, mu=mu, kappa=kappa, k_exp=pb.f32(3.0), rmode=3, group=True)   # k_exp must be ignored (Q8)
    ctx.set_cpgrid(cps.reshape(-1, 3), np.zeros((0, 3), np.int32), np.ones(S * P.N), np.tile(P.cxyz0, (S, 1)))
    ctx.set_group(S, P.N, P.T)
    ctx.set_labels(P.labels, P.centre)
    ctx.set_triplets(trip)
    rng = np.random.default_rng(0)
    n = 20000
    q = np.stack([rng.integers(0, S * P.T, n), rng.integers(0, P.L, n), rng.integers(0, P.L, n), rng.integers(0, P.L, n)], 1).astype(np.int32)
    g, o = ctx.triplet_costs(q), gc.triplet_costs(q)
    assert (o == 1e7).any()
    _check(g, o, 1e-3 * lam, "group triplets", tol=1e-9)
    ctx.close()


def test_group_pairwise_patch_correlation(built):
    """DiscreteGroupCostFunction::computePairwiseCost (src/DiscreteGroupCostFunction.cpp:32-54): correlation over the
    intersection of two (template vertex -> value) patch maps, incl. empty and single-element intersections."""
    from msm_newmeshreg_b200 import Context
    rng = np.random.default_rng(5)
    S, N, L, V = 3, 12, 5, 400
    nmaps = S * N * L
    off = [0]; keys = []; vals = []
    for m in range(nmaps):
        n = int(rng.integers(0, 60))
        k = np.sort(rng.choice(V // (1 if m % 7 else 8), size=min(n, V // 8), replace=False))   # some maps sparse / disjoint
        keys.append(k); vals.append(rng.normal(size=len(k)) + (5.0 if m % 3 == 0 else 0.0))
        off.append(off[-1] + len(k))
    keys = np.concatenate(keys).astype(np.int32); vals = np.concatenate(vals)
    pairs = np.array([[a * N + v, b * N + int(rng.integers(0, N))] for a in range(S) for v in range(N) for b in range(a + 1, S)], np.int32)
    cps = np.tile(pyorc.icosphere(0)[0], (S, 1))                      # 12 control points per subject (geometry unused here)
    ctx = Context()
    ctx.set_params(group=True, rmode=3)
    ctx.set_cpgrid(cps, np.zeros((0, 3), np.int32), np.ones(S * N), cps)
    ctx.set_group(S, N, 20)
    labels = np.tile(cps[0], (L, 1)) + rng.normal(size=(L, 3)) * 0.01
    ctx.set_labels(labels, labels[0])
    ctx.set_pairs(pairs)
    ctx.set_group_patches(off, keys, vals)
    gc = pyorc.GroupCost()
    gc.set(1.0, 2.0, 0.4, 1.6, cps.reshape(S, N, 3), pyorc.icosphere(0)[1], cps[:N], labels, np.tile(np.eye(3), (S * N, 1, 1)),
           np.zeros((S * 20, 3), np.int32))
    gc.set_patches(off, keys, vals, pairs)
    q = np.stack([rng.integers(0, len(pairs), 3000), rng.integers(0, L, 3000), rng.integers(0, L, 3000)], 1).astype(np.int32)
    g, o = ctx.group_pair_costs(q), gc.pair_costs(q)
    assert np.abs(g - o).max() < 1e-12
    assert (o == 0.5).any()                                           # empty / degenerate intersections give rho = 0
    lab = rng.integers(0, L, S * N).astype(np.int32)
    mv = ctx.group_pair_moves(lab, 2)
    qm = np.array([[p, 2 if c & 2 else lab[pairs[p, 0]], 2 if c & 1 else lab[pairs[p, 1]]] for p in range(len(pairs)) for c in range(4)], np.int32)
    assert np.abs(mv.ravel() - gc.pair_costs(qm)).max() < 1e-12      # include/Fusion/Fusion.h:169-172 order
    ctx.close()


def test_antipodal_control_point_is_well_defined(built):
    """The control point antipodal to the label-grid centre: rotation by pi about the agreed axis on both sides."""
    P = pb.make_problem(4, 2, 1, seed=14, warp=0.3)
    sing = pb.singular_nodes(P)
    assert len(sing) == 1
    ctx = pb.gpu_context(P); ctx.build_patches()
    cf = pb.oracle_costfunction(P); cf.get_source_data()
    _check(ctx.unary_table()[:, sing], cf.unary_costs()[:, sing], UFLOOR, "unary at the antipodal CP")
    ts = np.nonzero((P.triplets == sing[0]).any(1))[0]
    q = np.array([[t, a, b, c] for t in ts for a in (0, 3, 7) for b in (0, 5) for c in (0, 2, 11)], np.int32)
    _check(ctx.triplet_costs(q), cf.triplet_costs(q), 1e-3 * pb.DEFAULT_PARAMS["lam"], "triplets at the antipodal CP", tol=1e-9)
    ctx.close()


# ------------------------------------------------------------------------------------------------ pairwise
def test_pair_table_parity(built):
    P = pb.make_problem(4, 2, 1, seed=12, warp=0.3, deform_cp=0.2)
    for rexp in (pb.f32(2.0), pb.f32(1.0)):
        params = dict(rmode=1, rexp=rexp)
        ctx = pb.gpu_context(P, params=params, cliques="pairs")
        cf = pb.oracle_costfunction(P, params=params, cliques="pairs")
        # tolerance 1e-5 (north_star).  The reference's rotation estimate takes acos of a dot product; for the identity
        # label that yields a spurious ~2e-8 rad rotation about a noise axis whenever the dot rounds below 1, which
        # moves small pairwise costs by ~1e-6 relative.  The device uses atan2(|a x b|, a.b) and returns the identity.
        _check(ctx.pair_table(), cf.pair_table(), 1e-3 * pb.DEFAULT_PARAMS["lam"], f"pair table rexp={rexp}", tol=TOL)
        lab = np.random.default_rng(2).integers(0, P.L, P.N).astype(np.int32)
        ctx.build_patches(); ctx.unary_table()
        s = ctx.total_cost(lab)
        assert abs(s[1] - cf.total(lab, use_unary=False)) <= TOL * max(abs(s[1]), 1e-12)
        ctx.close()


# ------------------------------------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("name", ["uni_ico4_ico2", "multi_ico4_ico2", "uni_ties_ico3_ico1"])
def test_against_golden_fixtures(built, name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "small_tables.npz"))
    kw, params = mg.CASES[name]
    P = pb.make_problem(**kw)
    ctx = pb.gpu_context(P, params=params); ctx.build_patches()
    off, idx = ctx.patches()
    assert np.array_equal(np.diff(off), gold[f"{name}/patch_sizes"]) and np.array_equal(idx, gold[f"{name}/patch_idx"])
    good = _good_nodes(P)
    _check(ctx.unary_table()[:, good], gold[f"{name}/unary"][:, good], UFLOOR, "golden unary")
    q = gold[f"{name}/triplet_q"]
    ok = ~np.isin(P.triplets[q[:, 0]], pb.singular_nodes(P)).any(1)
    _check(ctx.triplet_costs(q)[ok], gold[f"{name}/triplet_costs"][ok], 1e-3 * pb.DEFAULT_PARAMS["lam"], "golden triplets", tol=1e-9)
    ctx.close()
    ctx = pb.gpu_context(P, params=dict(params, rmode=1), cliques="pairs")
    okp = ~np.isin(P.pairs[:8], pb.singular_nodes(P)).any(1)
    _check(ctx.pair_table()[:8][okp], gold[f"{name}/pair_table_first8"][okp], 1e-3 * pb.DEFAULT_PARAMS["lam"], "golden pairs", tol=TOL)
    ctx.close()


# ------------------------------------------------------------------------------------------------ full size (C3)
def test_full_size_c3_properties(built):
    """ico7 data / ico5 control grid (BASELINE.json configs[2]): oracle on a sample of rows + size-independent properties."""
    P = pb.make_problem(7, 5, 1, seed=1, warp=0.3, deform_cp=0.1)
    ctx = pb.gpu_context(P); ctx.build_patches()
    U = ctx.unary_table()
    assert U.shape == (19, 10242) and np.isfinite(U).all() and (U >= -1e-6).all() and (U <= 1 + 1e-6).all()
    U2 = ctx.unary_table()
    assert np.array_equal(U, U2)                                        # deterministic (fixed reduction order)
    cf = pb.oracle_costfunction(P)
    rows = 400
    cf.get_source_data(0, rows)
    off, idx = ctx.patches()
    po = cf.patches()
    assert all(np.array_equal(po[k], idx[off[k]:off[k + 1]]) for k in range(rows))
    _check(U[:, :rows], cf.unary_costs(0, rows), UFLOOR, "C3 unary rows")
    # a rigid rotation of EVERYTHING (target stays: it is the frame) changes nothing when source==target geometry is
    # rotated consistently: here the cheap invariant is shift/scale invariance of Pearson in the source features
    Q = pb.make_problem(7, 5, 1, seed=1, warp=0.3, deform_cp=0.1)
    Q.inp = 3.0 * Q.inp + 11.0
    c2 = pb.gpu_context(Q); c2.build_patches()
    assert np.abs(c2.unary_table() - U).max() < 5e-6
    c2.close()
    # triplet fusion-move table at full size vs oracle on a sample
    lab = np.random.default_rng(3).integers(0, P.L, P.N).astype(np.int32)
    mv = ctx.triplet_moves(lab, 4)
    ts = np.random.default_rng(4).integers(0, P.T, 300)
    q = []
    for t in ts:
        a, b, c = lab[P.triplets[t]]
        for combo in range(8):
            q.append([t, 4 if combo & 4 else a, 4 if combo & 2 else b, 4 if combo & 1 else c])
    _check(mv[ts].ravel(), cf.triplet_costs(np.array(q, np.int32)), 1e-3 * pb.DEFAULT_PARAMS["lam"], "C3 moves", tol=1e-9)
    ctx.close()


def test_unary_bits_do_not_depend_on_label_group_size(built, monkeypatch):
    """The per-item arithmetic is written so that an item gives the same bits wherever it is processed (two-in-flight body,
    single-item tail, any label-group size): forcing the group size must not change one bit of the table."""
    P = pb.make_problem(5, 3, 1, seed=5, warp=0.3, deform_cp=0.1)
    ctx = pb.gpu_context(P); ctx.build_patches()
    monkeypatch.delenv("MSM_UNARY_LG", raising=False)
    U = ctx.unary_table()
    for lg in (19, 10, 7, 3, 1):
        monkeypatch.setenv("MSM_UNARY_LG", str(lg))
        assert np.array_equal(ctx.unary_table(), U), f"label-group size {lg}"
    ctx.close()


@pytest.mark.parametrize("C", [1, 4])
def test_spatial_packing_order_only_reorders_sums(built, monkeypatch, C):
    """k_patch_pack packs samples in target-leaf order (cache locality); the public patch lists stay ascending and the
    table changes only by the rounding of re-ordered sums."""
    P = pb.make_problem(5, 3, C, seed=6, warp=0.3, deform_cp=0.1)
    tabs, lists = [], []
    for nosort in ("1", "0"):
        monkeypatch.setenv("MSM_PATCH_NOSORT", nosort)
        ctx = pb.gpu_context(P); ctx.build_patches()
        tabs.append(ctx.unary_table()); lists.append(ctx.patches()); ctx.close()
    assert np.array_equal(lists[0][0], lists[1][0]) and np.array_equal(lists[0][1], lists[1][1])
    assert pb.rel_err(tabs[1], tabs[0], 1e-2).max() < 2e-6
    cf = pb.oracle_costfunction(P); cf.get_source_data()
    assert pb.rel_err(tabs[1], cf.unary_costs(), 1e-2).max() < 1e-5


def test_fused_exchange_kernels_fill_every_peer_table(built):
    """msm_unary_table_peers / msm_triplet_moves_range_peers (control-point / triplet sharding): two shard contexts on one
    GPU each store into BOTH "rank" tables; afterwards every table must be the complete single-context table, bit for bit."""
    import torch
    from msm_newmeshreg_b200 import sharding as sh
    P = pb.make_problem(5, 3, 1, seed=7, warp=0.3, deform_cp=0.1)
    ctx = pb.gpu_context(P); ctx.build_patches()
    U = ctx.unary_table().astype(np.float32)
    lab = np.random.default_rng(1).integers(0, P.L, P.N).astype(np.int32)
    M = ctx.triplet_moves(lab, -1).astype(np.float32)
    ctx.close()
    world = 2
    tabs_u = [torch.full((P.L, P.N), float("nan"), dtype=torch.float32, device="cuda") for _ in range(world)]
    tabs_m = [torch.full((P.L, P.T, 8), float("nan"), dtype=torch.float32, device="cuda") for _ in range(world)]
    lab_dev = torch.from_numpy(lab).cuda()
    for r in range(world):
        c = pb.gpu_context(P)
        c.set_shard(*sh.shard_range(P.N, r, world)); c.build_patches()
        c.unary_table_peers([t.data_ptr() for t in tabs_u])
        t0, t1 = sh.shard_range(P.T, r, world)
        c.triplet_moves_range_peers(lab_dev.data_ptr(), -1, t0, t1, [t.data_ptr() for t in tabs_m])
        c.synchronize(); c.close()
    for r in range(world):
        assert np.array_equal(tabs_u[r].cpu().numpy(), U)
        assert np.array_equal(tabs_m[r].cpu().numpy(), M)


def test_group_triplets_full_size(built):
    """BASELINE configs[4] at full size (16 subjects x ico4 control grids, 81 920 triangles): a sample of the all-label
    fusion-move table against the oracle port and, when built, the reference's own DiscreteGroupCostFunction."""
    import torch
    from msm_newmeshreg_b200 import Context
    from oracle import pyrefcost
    S, lvl = 16, 4
    cxyz0, ctris = pyorc.icosphere(lvl)
    maxsep, mvdmax, mvd = pyorc.spacings(cxyz0, ctris)
    samples, bary, centre = pyorc.label_grid(lvl + 2, 0.5 * mvdmax)
    trip1 = pyorc.triplets(cxyz0, ctris)
    N, T, L = len(cxyz0), len(trip1), len(bary)
    cps = np.stack([pb.smooth_warp(cxyz0, 500 + s, 0.25 * mvd) for s in range(S)])
    trip = np.concatenate([trip1 + s * N for s in range(S)]).astype(np.int32)
    lam, mu, kappa = pb.f32(0.2), pb.f32(0.4), pb.f32(1.6)
    ctx = Context()
    ctx.set_params(lam=lam, rexp=2.0, mu=mu, kappa=kappa, rmode=3, group=True)
    ctx.set_cpgrid(cps.reshape(-1, 3), np.zeros((0, 3), np.int32), np.ones(S * N), np.tile(cxyz0, (S, 1)))
    ctx.set_group(S, N, T); ctx.set_labels(bary, centre); ctx.set_triplets(trip)
    lab = np.random.default_rng(7).integers(0, L, S * N).astype(np.int32)
    full = torch.full((L, S * T, 8), float("nan"), dtype=torch.float32, device="cuda")
    ctx.triplet_moves_range_peers(torch.from_numpy(lab).cuda().data_ptr(), -1, 0, S * T, [full.data_ptr()])
    ctx.synchronize()
    M = full.cpu().numpy()
    assert np.isfinite(M).all()
    rng = np.random.default_rng(3)
    qa, qt, qc = rng.integers(0, L, 4000), rng.integers(0, S * T, 4000), rng.integers(0, 8, 4000)
    q = np.stack([qt, np.where(qc & 4, qa, lab[trip[qt, 0]]), np.where(qc & 2, qa, lab[trip[qt, 1]]), np.where(qc & 1, qa, lab[trip[qt, 2]])], 1).astype(np.int32)
    rot = np.concatenate([pyorc.rotations(centre, cps[s]) for s in range(S)])
    got = M[qa, qt, qc].astype(np.float64)
    for G in (pyorc.GroupCost, pyrefcost.GroupCost if pyrefcost.available() else None):
        if G is None:
            continue
        g = G(); g.set(lam, 2.0, mu, kappa, cps, ctris, cxyz0, bary, rot, trip)
        want = g.triplet_costs(q)
        assert (np.abs(got - want) / np.maximum(np.abs(want), 1e-3 * lam)).max() < 1e-6     # fp32 table
    ctx.close()


def test_fused_exchange_argument_errors(built):
    import torch
    from msm_newmeshreg_b200 import MsmError
    P = pb.make_problem(3, 2, 1, seed=8, warp=0.2)
    ctx = pb.gpu_context(P); ctx.build_patches()
    buf = torch.zeros((P.L, P.N), dtype=torch.float32, device="cuda")
    with pytest.raises(MsmError, match="world size"):
        ctx.unary_table_peers([])
    with pytest.raises(MsmError, match="world size"):
        ctx.unary_table_peers([buf.data_ptr()] * 17)
    with pytest.raises(MsmError, match="null peer"):
        ctx.unary_table_peers([buf.data_ptr(), 0])
    lab = torch.zeros(P.N, dtype=torch.int32, device="cuda")
    with pytest.raises(MsmError, match="bad triplet range"):
        ctx.triplet_moves_range_peers(lab.data_ptr(), -1, 5, P.T + 1, [buf.data_ptr()])
    ctx.unary_table_peers([buf.data_ptr()]); ctx.synchronize()          # still usable after the errors
    assert np.array_equal(buf.cpu().numpy().astype(np.float64), ctx.unary_table())
    ctx.close()
