"""BASELINE.json's own configurations at FULL size against THE REFERENCE ITSELF (oracle/_ref/libref_cost.so = its
cost-function / model / optimiser sources compiled in place), plus the named parity holes of the round-1 review:

  * configs[2] (ico7 data x ico5 control grid): ALL 10 242 rows of the unary table, C = 1 and C = 4, every patch list, and
    fusion-move triplet costs of whole proposal labels;
  * configs[1] (ico6, 4 channels, control grids ico2 -> ico3 -> ico4, strain regulariser, HOCR): the complete 3-level
    registration — labellings of every outer iteration identical, warped sphere within 1e-4;
  * a near-constant patch (variance ~1e-10 of the mean square) against src/similarities.cc:171;
  * point location ON edges and vertices: the chosen triangle is one of the incident ones, the weights reproduce the point.

Slow on the CPU side (the reference's single-threaded brute-force patch search: ~25 s per subject), hence one file."""
import numpy as np
import pytest

import problems as pb
from oracle import pyorc, pyref, pyrefcost

pytestmark = pytest.mark.gpu
UFLOOR = 1e-2           # |d| <= 1e-5 max(|ref|, 1e-2) for the unary term
TOL = 1e-5              # north_star: cost tables within 1e-5 relative (fp32)
need_ref = pytest.mark.skipif(not pyrefcost.available(), reason="oracle/_ref/libref_cost.so not built (needs /root/reference at build time)")


def _ref_model(P, **kw):
    pr = dict(pb.DEFAULT_PARAMS); pr.update(kw)
    m = pyrefcost.model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=P.C > 1, threads=1, **pr)
    m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref)
    m.initialize(); m.reset_meshspace(P.sxyz); m.reset_cpgrid(P.cxyz); m.setup()
    return m


@need_ref
@pytest.mark.parametrize("C", [1, 4])
def test_c3_every_row_against_the_reference_build(built, C):
    P = pb.make_problem(7, 5, C, seed=1, warp=0.3, deform_cp=0.1)
    ctx = pb.gpu_context(P); ctx.build_patches()
    U = ctx.unary_table()
    off, idx = ctx.patches()
    m = _ref_model(P)                                                   # the reference: ~25 s of brute-force patch search
    assert (m.N, m.L, m.T) == (10242, 19, 20480) and np.array_equal(m.triplets(), P.triplets)
    # patch lists: every control point, bit exact
    rp = pyrefcost.patches(m)
    assert sum(len(p) for p in rp) == len(idx)
    assert all(np.array_equal(rp[k], idx[off[k]:off[k + 1]]) for k in range(P.N))
    # unary table: every row.  The control point antipodal to the label-grid centre is compared separately below: its
    # rotation is by pi about an axis the published formula leaves undefined, fixed by the same convention
    # on all three sides.
    Ur = m.unary()
    err = pb.rel_err(U, Ur, UFLOOR)
    sing = pb.singular_nodes(P)
    good = np.ones(P.N, bool); good[sing] = False
    assert err[:, good].max() < TOL, (C, float(err[:, good].max()), np.unravel_index(err[:, good].argmax(), err[:, good].shape))
    if len(sing):                                                       # (a deformed grid has no exactly antipodal control point)
        assert err[:, sing].max() < 1e-3                                # same convention: still close, not bit-for-bit geometry
    # fusion-move triplet costs of three whole proposal labels against the reference's computeTripletCost
    lab = np.random.default_rng(3).integers(0, P.L, P.N).astype(np.int32)
    okt = ~np.isin(P.triplets, sing).any(1)
    for alpha in (0, 7, P.L - 1):
        mv = ctx.triplet_moves(lab, alpha)
        mr = m.triplet_moves(lab, alpha)
        e = pb.rel_err(mv[okt], mr[okt], 1e-3 * pb.DEFAULT_PARAMS["lam"])
        assert e.max() < 1e-9, (alpha, float(e.max()))
    m.close(); ctx.close()


@need_ref
@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libref_optim.so not built (needs /root/reference at build time)")
def test_configs1_three_level_registration_equals_the_reference(built):
    """BASELINE.json configs[1] as written: pairwise ico6, 4-channel features, control grids ico2 -> ico3 -> ico4, strain
    regulariser, HOCR, 3 outer iterations per level.  GPU-backed classes + the reference's optimisers
    against the reference's own classes + optimisers: labellings identical at every iteration of every level, final
    warped sphere within 1e-4 (north_star)."""
    import os
    import level_walltime as lw
    levels, iters, C = [2, 3, 4], 3, 4
    P = pb.make_problem(6, levels[0], C, seed=51, warp=0.0)
    P.seed = 51
    lam = pb.f32(0.1)
    src_gpu, labs_gpu, rep = lw.run_gpu(P, levels, iters, C, os.cpu_count() or 1, lam)
    src_ref, labs_ref, rep_ref = lw.run_reference(P, levels, iters, C, 1, lam)
    for lvl, g, r in zip(levels, labs_gpu, labs_ref):
        assert g.shape == r.shape, (lvl, g.shape, r.shape)
        assert np.array_equal(g, r), (lvl, int((g != r).sum()))
    assert np.abs(src_gpu - src_ref).max() < 1e-4
    assert np.linalg.norm(src_gpu - P.txyz, axis=1).max() > 0.5        # the registration did move the sphere


def test_near_constant_patch_against_similarities_cc_171(built):
    """src/similarities.cc:171 returns rho = 0 only when a variance is EXACTLY zero.  The device works from single-pass
    moments of shifted fp32 data and treats a variance below 1e-14 of the raw second moment as zero (rounding noise of
    that formula).  Patches whose variance is tiny but real — 1e-10 and 1e-6 of the mean square — must still give the
    reference's correlation, and an exactly constant patch the reference's 0.5."""
    P = pb.make_problem(5, 2, 1, seed=21, warp=0.3)
    base = pb.random_field(P.txyz, 77)
    moved = pb.rotate(P.txyz, [0.3, -0.5, 0.8], 2.0)
    base_s = pb.random_field(moved, 77)
    for eps in (1e-5, 1e-3, 0.0):                                      # relative variance eps^2 = 1e-10, 1e-6, 0
        Q = pb.make_problem(5, 2, 1, seed=21, warp=0.3)
        Q.ref = (3.0 + eps * base)[None]                               # nearly constant target field around a large mean
        Q.inp = (3.0 + eps * base_s)[None]
        ctx = pb.gpu_context(Q); ctx.build_patches()
        U = ctx.unary_table()
        cf = pb.oracle_costfunction(Q); cf.get_source_data()
        Uo = cf.unary_costs()
        good = np.ones(Q.N, bool); good[pb.singular_nodes(Q)] = False
        err = pb.rel_err(U[:, good], Uo[:, good], UFLOOR).max()
        if eps == 0.0:
            assert np.all(Uo == 0.5) and np.all(U == 0.5)              # both variances exactly zero -> rho = 0 -> 0.5 AW
        else:
            assert Uo[:, good].std() > 1e-3                            # the correlation is real, not the zero-variance branch
            assert err < TOL, (eps, float(err))
        if pyrefcost.available():                                      # the reference's own classes say the same
            m = _ref_model(Q)
            assert pb.rel_err(U[:, good], m.unary()[:, good], UFLOOR).max() < TOL
            m.close()
        ctx.close()


@pytest.mark.parametrize("level", [1, 3, 6])
def test_locate_on_edges_and_vertices_picks_an_incident_triangle(built, level):
    """Index work must be exact: away from edges the triangle id is bit-identical to the oracle's
    (test_locate_bit_exact_triangles); ON an edge or a vertex several triangles are equally right, and the GPU must return
    one of them — with weights that reproduce the point — never a neighbour's neighbour."""
    from msm_newmeshreg_b200 import Context
    xyz, tris = pyorc.icosphere(level)
    ctx = Context()
    ctx.set_target(xyz, tris, None)
    rng = np.random.default_rng(level)
    # (1) every vertex itself
    t, w = ctx.locate(xyz)
    assert all(v in tris[t[v]] for v in range(len(xyz)))
    assert np.abs(np.einsum("nj,njk->nk", w, xyz[tris[t]]) - xyz).max() < 1e-9
    # (2) points exactly on edges (radially projected edge points, several positions along each sampled edge)
    E = np.unique(np.sort(np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]]), 1), axis=0)
    E = E[rng.choice(len(E), min(len(E), 4000), replace=False)]
    for s in (0.5, 0.25, 1e-3, 1.0 - 1e-6):
        p = (1 - s) * xyz[E[:, 0]] + s * xyz[E[:, 1]]
        p *= 100.0 / np.linalg.norm(p, axis=1, keepdims=True)
        t, w = ctx.locate(p)
        tv = tris[t]
        has_a = (tv == E[:, :1]).any(1); has_b = (tv == E[:, 1:]).any(1)
        assert (has_a & has_b).all(), (level, s, int((~(has_a & has_b)).sum()))     # a triangle containing the WHOLE edge
        assert (w > -1e-9).all() and np.allclose(w.sum(1), 1.0)
        # the weight of the vertex opposite the edge vanishes; the other two reproduce the radial projection of the point
        third = np.where((tv != E[:, :1]) & (tv != E[:, 1:]), w, 0.0).sum(1)
        assert np.abs(third).max() < 1e-9
        q = np.einsum("nj,njk->nk", w, xyz[tv])
        q *= 100.0 / np.linalg.norm(q, axis=1, keepdims=True)
        assert np.abs(q - p).max() < 1e-8
    ctx.close()
