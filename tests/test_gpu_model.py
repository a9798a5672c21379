"""GPU tests of the drop-in boundary: the host-side mirror classes (NonLinearSRegDiscreteModel + GPU-backed cost
functions, msm_newmeshreg_b200/host/src) against the oracle's restatement of the same model glue, and the
optimiser-equivalence tests: the reference's own, unmodified FastPD / Fusion(HOCR) optimisers (oracle/_ref) must return
IDENTICAL labellings whether they are driven by the GPU-backed model or by the oracle's cost tables."""
import numpy as np
import pytest

import problems as pb
from oracle import pyorc, pyref

pytestmark = pytest.mark.gpu


def _model(P, lib=None, **kw):
    from msm_newmeshreg_b200 import host
    pr = dict(pb.DEFAULT_PARAMS); pr.update(kw)
    m = host.Model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=P.C > 1, _lib_override=lib, **pr)
    # as in the reference driver: the level starts with source == target == the undeformed ico template
    # (model->set_meshspace(SPH_orig, SPH_orig), src/mesh_registration.cpp:66; _ORIG is that mesh), the deformed source
    # arrives through reset_meshspace (:336)
    m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref)
    m.initialize()
    m.reset_meshspace(P.sxyz)
    return m


def test_model_glue_matches_oracle(built):
    """Label grid, cliques, control grid and the unary table through the class interface."""
    P = pb.make_problem(4, 2, 1, seed=31, warp=0.3)
    m = _model(P, rmode=3)
    m.setup()                                   # iteration 1: barycentre labels
    assert (m.N, m.L, m.T) == (P.N, len(P.barycentres), P.T)
    assert np.array_equal(m.labels(), P.barycentres) and np.array_equal(m.triplets(), P.triplets)
    assert np.abs(m.cpgrid() - P.cxyz0).max() < 1e-12     # recentre() of the driver shifts coordinates by ~1e-15
    cf = pb.oracle_costfunction(P); cf.get_source_data()
    good = np.ones(P.N, bool); good[pb.singular_nodes(P)] = False
    assert pb.rel_err(m.unary()[:, good], cf.unary_costs()[:, good], 1e-2).max() < 1e-5
    q = np.array([[5, 3], [17, 0], [100, 18]], np.int32)
    assert np.array_equal(m.costs("unary", q), m.unary()[q[:, 1], q[:, 0]])   # per-element virtual = table look-up
    m.setup()                                   # iteration 2: vertex samples (src/DiscreteModel.cpp:215-220)
    assert np.array_equal(m.labels(), P.samples)
    mp = _model(P, rmode=1, dopt="FastPD")
    mp.setup()
    assert np.array_equal(mp.pairs(), P.pairs)
    m.close(); mp.close()


def test_apply_labeling(built):
    P = pb.make_problem(3, 2, 1, seed=32, warp=0.2)
    m = _model(P, rmode=3); m.setup()
    lab = np.random.default_rng(0).integers(0, m.L, m.N).astype(np.int32)
    m.set_labeling(lab)
    new = m.apply_labeling()
    expect = np.einsum("nij,nj->ni", P.rot, P.labels[lab])          # CP_k <- ROT[k] * labels[lab[k]]  (:238-243)
    good = np.ones(P.N, bool); good[pb.singular_nodes(P)] = False
    assert np.abs(new - expect)[good].max() < 1e-9
    m.close()


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libref_optim.so not built (needs /root/reference at build time)")
def test_fastpd_labelling_identical(built):
    """--dopt FastPD (src/mesh_registration.cpp:346-352): GPU-backed model vs oracle tables, same unmodified optimiser."""
    P = pb.make_problem(4, 2, 1, seed=33, warp=0.3, deform_cp=0.1)
    m = _model(P, lib=pyref.lib(), rmode=1, dopt="FastPD")
    m.reset_cpgrid(P.cxyz)
    m.setup()
    lab_gpu, e_gpu = pyref.fastpd_model(m)
    cf = pb.oracle_costfunction(P, params=dict(rmode=1), cliques="pairs"); cf.get_source_data()
    lab_orc, e_orc = pyref.fastpd_tables(cf.unary_costs(), P.pairs, cf.pair_table())
    assert np.array_equal(lab_gpu, lab_orc), f"{(lab_gpu != lab_orc).sum()} of {P.N} labels differ"
    assert abs(e_gpu - e_orc) <= 1e-5 * abs(e_orc)
    m.close()


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libref_optim.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("C", [1, 4])
def test_fusion_hocr_labelling_identical(built, C):
    """--dopt HOCR with the strain regulariser (--regoption 3): Fusion::optimize -> ELC/HOCR -> inner FastPD."""
    P = pb.make_problem(4, 2, C, seed=34, warp=0.3, deform_cp=0.1)
    m = _model(P, lib=pyref.lib(), rmode=3, dopt="HOCR")
    m.reset_cpgrid(P.cxyz)
    m.setup()
    lab_gpu, e_gpu = pyref.fusion_model(m, threads=4)
    cf = pb.oracle_costfunction(P, params=dict(rmode=3)); cf.get_source_data()
    lab_orc, e_orc = pyref.fusion_tables(cf.unary_costs(), P.triplets, cf.triplet_table(), threads=4)
    assert np.array_equal(lab_gpu, lab_orc), f"{(lab_gpu != lab_orc).sum()} of {P.N} labels differ"
    assert abs(e_gpu - e_orc) <= 1e-5 * abs(e_orc)
    m.close()


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libref_optim.so not built (needs /root/reference at build time)")
def test_outer_iterations_labellings_identical(built):
    """Two outer iterations of run_discrete_opt (src/mesh_registration.cpp:320-396) restated with the source mesh held
    fixed (the source warp is a 'next' row): setupCostFunction -> HOCR -> applyLabeling -> reset_CPgrid, with
    the label set alternating between barycentres and samples; labellings must agree at every iteration."""
    P = pb.make_problem(4, 2, 1, seed=35, warp=0.3)
    m = _model(P, lib=pyref.lib(), rmode=3, dopt="HOCR")
    cp = P.cxyz0.copy()
    for it in (1, 2):
        m.setup()
        lab_gpu, _ = pyref.fusion_model(m, threads=4)
        Q = pb.make_problem(4, 2, 1, seed=35, warp=0.3, iteration=it)
        Q.cxyz = cp; Q.rot = pyorc.rotations(Q.centre, cp)
        cf = pb.oracle_costfunction(Q, params=dict(rmode=3)); cf.get_source_data()
        lab_orc, _ = pyref.fusion_tables(cf.unary_costs(), Q.triplets, cf.triplet_table(), threads=4)
        assert np.array_equal(lab_gpu, lab_orc), f"iteration {it}: {(lab_gpu != lab_orc).sum()} labels differ"
        new_gpu = m.apply_labeling()
        cp = np.einsum("nij,nj->ni", Q.rot, Q.labels[lab_orc])
        cp *= 100.0 / np.linalg.norm(cp, axis=1, keepdims=True)
        sing = pb.singular_nodes(Q)
        keep = np.ones(P.N, bool); keep[sing] = False
        assert np.abs(new_gpu - cp)[keep].max() < 1e-9
        cp[sing] = new_gpu[sing]
        m.reset_cpgrid(cp)
    m.close()


# ---------------------------------------------------------------------------------------------------------------------
# GPU path against THE REFERENCE ITSELF (oracle/_ref/libref_cost.so: the reference's own sources compiled in place)
# ---------------------------------------------------------------------------------------------------------------------
def _ref_model(P, **kw):
    from oracle import pyrefcost
    pr = dict(pb.DEFAULT_PARAMS); pr.update(kw)
    m = pyrefcost.model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=P.C > 1, threads=1, **pr)
    m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref)
    m.initialize()
    m.reset_meshspace(P.sxyz)
    return m


def _have_refcost():
    from oracle import pyrefcost
    return pyrefcost.available()


@pytest.mark.skipif(not _have_refcost(), reason="oracle/_ref/libref_cost.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("C", [1, 4])
def test_gpu_model_equals_reference_model(built, C):
    """Same driver call sequence (src/mesh_registration.cpp:66,86,336-337,347) on the GPU-backed mirror classes and on the
    reference's own classes: unary table 1e-5 rel (fp32 path), strain / total 1e-9 rel (fp64 paths), move tables."""
    P = pb.make_problem(4, 2, C, seed=61 + C, warp=0.3, deform_cp=0.12)
    g, r = _model(P, rmode=3), _ref_model(P, rmode=3)
    for m in (g, r):
        m.reset_cpgrid(P.cxyz); m.setup()
    assert (g.N, g.L, g.T) == (r.N, r.L, r.T)
    assert np.array_equal(g.labels(), r.labels()) and np.array_equal(g.triplets(), r.triplets())
    Ug, Ur = g.unary(), r.unary()
    assert pb.rel_err(Ug, Ur, 1e-2).max() < 1e-5
    rng = np.random.default_rng(C)
    q = np.stack([rng.integers(0, P.T, 3000)] + [rng.integers(0, P.L, 3000) for _ in range(3)], 1).astype(np.int32)
    assert pb.rel_err(g.costs("triplet", q), r.costs("triplet", q), 1e-9).max() < 1e-5   # look-up table is fp32
    lab = rng.integers(0, P.L, P.N).astype(np.int32)
    assert pb.rel_err(g.triplet_moves(lab, 5), r.triplet_moves(lab, 5), 1e-9).max() < 1e-9
    for m in (g, r):
        m.set_labeling(lab)
    assert abs(g.total() - r.total()) <= 1e-5 * abs(r.total())
    assert np.abs(g.apply_labeling() - r.apply_labeling()).max() < 1e-9
    g.close(); r.close()


@pytest.mark.skipif(not _have_refcost(), reason="oracle/_ref/libref_cost.so not built (needs /root/reference at build time)")
def test_gpu_pairwise_model_equals_reference_model(built):
    P = pb.make_problem(3, 2, 1, seed=66, warp=0.2, deform_cp=0.1)
    g, r = _model(P, rmode=1, dopt="FastPD"), _ref_model(P, rmode=1, dopt="FastPD")
    for m in (g, r):
        m.reset_cpgrid(P.cxyz); m.setup()
    assert np.array_equal(g.pairs(), r.pairs())
    rng = np.random.default_rng(1)
    q = np.stack([rng.integers(0, g.P, 3000), rng.integers(0, P.L, 3000), rng.integers(0, P.L, 3000)], 1).astype(np.int32)
    a, b = g.costs("pair", q), r.costs("pair", q)
    assert np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1e-3)       # identity-label acos noise, see test_gpu_parity
    g.close(); r.close()


@pytest.mark.skipif(not (_have_refcost() and pyref.available()), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("dopt,rmode", [("HOCR", 3), ("FastPD", 1)])
def test_run_discrete_opt_gpu_equals_reference_end_to_end(built, dopt, rmode):
    """Mesh_registration::run_discrete_opt (src/mesh_registration.cpp:316-397), three outer iterations: the reference all on
    the CPU (its cost functions, its optimisers, its unfold) against the same driver over the GPU-backed cost functions.
    Labellings identical at every iteration; warped source vertices within 1e-4 (north_star)."""
    from msm_newmeshreg_b200 import host
    from oracle import pyrefcost
    P = pb.make_problem(4, 2, 1, seed=71, warp=0.0)
    pr = dict(pb.DEFAULT_PARAMS); pr.update(rmode=rmode, dopt=dopt)
    g = host.Model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=False, _lib_override=pyref.lib(), **pr)
    r = pyrefcost.model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=False, threads=1, **pr)
    for m in (g, r):
        m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref); m.initialize()
    start = pb.smooth_warp(P.txyz, 171, 0.2 * P.mvd)
    sg, lg, eg, _ = pyref.run_discrete_opt(g, start, 3, threads=4)
    sr, lr, er, _ = pyrefcost.run_discrete_opt(r, start, 3, threads=1)
    assert len(lg) == len(lr)
    sing = pb.singular_nodes(P)
    keep = np.ones(P.N, bool); keep[sing] = False
    for it, (a, b) in enumerate(zip(lg, lr)):
        assert np.array_equal(a[keep], b[keep]), f"iteration {it + 1}: {(a != b).sum()} labels differ"
    assert np.allclose(eg, er, rtol=1e-5)
    d_sing = np.linalg.norm(P.txyz - P.cxyz0[sing[0]], axis=1) if len(sing) else np.full(len(P.txyz), 1e9)
    assert np.abs(sg - sr)[d_sing > 2.5 * P.mvdmax].max() < 1e-4
    g.close(); r.close()


@pytest.mark.skipif(not (_have_refcost() and pyref.available()), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("dopt,rmode", [("HOCR", 3), ("FastPD", 1)])
def test_the_references_own_driver_loop_runs_unchanged_over_the_mirror(built, dopt, rmode):
    """Drop-in at source level: the TEXT of Mesh_registration::run_discrete_opt (src/mesh_registration.cpp:316-397; cut out at
    build time, oracle/Makefile) compiled unmodified against the product's mirror classes drives the GPU, and — compiled against
    the reference's own classes — is the reference.  Same labelling, control grid and warped source after three iterations."""
    from msm_newmeshreg_b200 import host
    from oracle import pyrefcost
    if not (pyref.has_driver_loop() and pyrefcost.has_driver_loop()):
        pytest.skip("oracle/_ref built without the driver loop")
    P = pb.make_problem(4, 2, 1, seed=73, warp=0.0)
    pr = dict(pb.DEFAULT_PARAMS); pr.update(rmode=rmode, dopt=dopt)
    g = host.Model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=False, _lib_override=pyref.lib(), **pr)
    r = pyrefcost.model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=False, threads=1, **pr)
    for m in (g, r):
        m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref); m.initialize()
    start = pb.smooth_warp(P.txyz, 173, 0.2 * P.mvd)
    sg, gg, lg = pyref.driver_run_discrete_opt(g, start, 3, threads=4)
    sr, gr, lr = pyrefcost.driver_run_discrete_opt(r, start, 3, threads=1)
    sing = pb.singular_nodes(P)
    keep = np.ones(P.N, bool); keep[sing] = False
    assert np.array_equal(lg[keep], lr[keep])
    assert np.abs(gg - gr)[keep].max() < 1e-4
    d_sing = np.linalg.norm(P.txyz - P.cxyz0[sing[0]], axis=1) if len(sing) else np.full(len(P.txyz), 1e9)
    assert np.abs(sg - sr)[d_sing > 2.5 * P.mvdmax].max() < 1e-4
    # and it is the loop the harness restates (tests above): same warped source
    g2 = host.Model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=False, _lib_override=pyref.lib(), **pr)
    g2.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref); g2.initialize()
    s2, _, _, _ = pyref.run_discrete_opt(g2, start, 3, threads=4)
    assert np.abs(s2 - sg).max() < 1e-9
    g.close(); r.close(); g2.close()


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libref_optim.so not built (needs /root/reference at build time)")
def test_the_references_mcmc_optimiser_runs_over_the_mirror(built):
    """MCMC::optimise (src/mcmc_opt.h, included where it lies) compiled against the mirror: reads the GPU-built unary table and
    the re-entrant triplet virtual.  Randomised (std::random_device): only the energy contract is checked — it never accepts
    a move that raises unary + triplet of the clique it looks at, so the total stays finite and the labelling valid."""
    from msm_newmeshreg_b200 import host
    if not pyref.has_driver_loop():
        pytest.skip("oracle/_ref built without the driver loop")
    P = pb.make_problem(4, 2, 1, seed=74, warp=0.0)
    pr = dict(pb.DEFAULT_PARAMS); pr.update(rmode=3, dopt="MCMC")
    g = host.Model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=False, _lib_override=pyref.lib(), **pr)
    g.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref); g.initialize()
    start = pb.smooth_warp(P.txyz, 174, 0.2 * P.mvd)
    s, grid, lab = pyref.driver_run_discrete_opt(g, start, 1, threads=2, mciters=20)
    assert np.isfinite(s).all() and np.isfinite(grid).all()
    assert lab.min() >= 0 and lab.max() < len(P.labels) and (lab != 0).any()
    assert np.abs(np.linalg.norm(s, axis=1) - 100.0).max() < 1e-6
    g.close()


def test_triplet_moves_fp32_variant_is_the_rounded_fp64_table(built):
    P = pb.make_problem(4, 2, 1, seed=81, warp=0.3, deform_cp=0.1)
    m = _model(P, rmode=3); m.reset_cpgrid(P.cxyz); m.setup()
    lab = np.random.default_rng(2).integers(0, m.L, m.N).astype(np.int32)
    d = m.triplet_moves(lab, -1)
    f = m.triplet_moves(lab, -1, out=np.empty((m.L, m.T, 8), np.float32))
    assert f.dtype == np.float32 and np.array_equal(f, d.astype(np.float32))
    f1 = m.triplet_moves(lab, 4, out=np.empty((m.T, 8), np.float32))
    assert np.array_equal(f1, d[4].astype(np.float32))
    m.close()


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libref_optim.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("C", [1, 4])
def test_table_aware_fusion_driver_equals_fusion_optimize(built, C):
    """host/src/FusionTables.h (batched move tables, no T*L^3 table, no per-element virtuals) against the reference's
    unmodified Fusion::optimize on the same GPU-backed model: same labelling, same energy."""
    P = pb.make_problem(4, 2, C, seed=91, warp=0.3, deform_cp=0.1)
    labs, ens = [], []
    for driver in (pyref.fusion_model, pyref.fusion_tables_model):
        m = _model(P, lib=pyref.lib(), rmode=3, dopt="HOCR")
        m.reset_cpgrid(P.cxyz); m.setup()
        lab, e = driver(m, threads=4)
        labs.append(lab.copy()); ens.append(e)
        m.close()
    assert np.array_equal(labs[0], labs[1]), f"{(labs[0] != labs[1]).sum()} of {P.N} labels differ"
    assert abs(ens[0] - ens[1]) <= 1e-6 * abs(ens[0])
    assert len(np.unique(labs[0])) > 1


@pytest.mark.skipif(not _have_refcost(), reason="oracle/_ref/libref_cost.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("C", [1, 4])
def test_non_unit_costfunction_weights_give_the_reference_absolute_weights(built, C):
    """resample_weights (src/DiscreteCostFunction.cpp:364-383): with a non-constant cost-function weighting the per-control-
    point AbsoluteWeights are the adaptive-barycentric resampling ([EXT] A8) of the per-vertex maximum weight — computed by
    the product (GPU point location + host weight assembly), not taken as an input.  The whole unary table against the
    reference's own classes."""
    P = pb.make_problem(4, 2, C, seed=23, warp=0.3, deform_cp=0.1)
    rows = 1 if C == 1 else C
    w = 0.25 + np.abs(np.stack([pb.random_field(P.sxyz, 900 + r) for r in range(rows)]))
    tabs = []
    for lib in (None, "ref"):
        m = _model(P) if lib is None else _ref_model(P)
        m.reset_meshspace(P.sxyz); m.reset_cpgrid(P.cxyz)
        m.set_weights(w)
        m.setup()
        tabs.append(m.unary()); m.close()
    good = np.ones(P.N, bool); good[pb.singular_nodes(P)] = False
    e = pb.rel_err(tabs[0][:, good], tabs[1][:, good], 1e-2)
    assert e.max() < 1e-5, float(e.max())
    # and the weights do matter: the same table with unit weights is different
    m = _model(P); m.reset_meshspace(P.sxyz); m.reset_cpgrid(P.cxyz); m.setup()
    assert np.abs(m.unary() - tabs[0]).max() > 1e-2
    m.close()
