"""Mesh-warp row: barycentric_mesh_interpolation on the GPU and unfold on the host against the
oracle's restatement, and the complete outer loop of run_discrete_opt (src/mesh_registration.cpp:316-397) — cost
evaluation, the reference's own HOCR optimiser, applyLabeling, source warp, unfold — against the same loop assembled from
oracle pieces: identical labellings per iteration and warped sphere vertices within 1e-4 (north_star)."""
import numpy as np
import pytest

import problems as pb
from oracle import pyorc, pyref

RAD = 100.0


def test_oracle_mesh_interpolation_known_answers():
    """CPU: identity warp leaves points in place; a rigid rotation of the control grid rotates every point."""
    cxyz, ctris = pyorc.icosphere(2)
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(500, 3)); pts *= RAD / np.linalg.norm(pts, axis=1, keepdims=True)
    assert np.abs(pyorc.mesh_interpolate(pts, cxyz, ctris, cxyz) - pts).max() < 1e-10
    R = pyorc.rotation_matrix([1, 0, 0], [0.9, 0.3, -0.2])
    out = pyorc.mesh_interpolate(pts, cxyz, ctris, cxyz @ R.T)
    assert np.abs(out - pts @ R.T).max() < 1e-9                   # barycentric combination commutes with a linear map
    assert np.allclose(np.linalg.norm(out, axis=1), RAD)


def test_oracle_unfold():
    """CPU: an unfolded mesh is untouched; a vertex pushed across its 1-ring is detected and repaired."""
    xyz, tris = pyorc.icosphere(3)
    same, it = pyorc.unfold(xyz, tris)
    assert it == 0 and np.array_equal(same, xyz) and pyorc.count_folded(xyz, tris) == 0
    bad = xyz.copy()
    v = 100
    nb = np.unique(tris[(tris == v).any(1)]); nb = nb[nb != v]
    far = nb[0]
    bad[v] = xyz[far] + 1.2 * (xyz[far] - xyz[v]); bad[v] *= RAD / np.linalg.norm(bad[v])   # jump over a neighbour
    assert pyorc.count_folded(bad, tris) > 0
    fixed, it = pyorc.unfold(bad, tris)
    assert it >= 1 and pyorc.count_folded(fixed, tris) == 0
    assert np.allclose(np.linalg.norm(fixed, axis=1), RAD)


@pytest.mark.gpu
@pytest.mark.parametrize("cp_level,deform", [(2, 0.0), (3, 0.3), (4, 0.25)])
def test_mesh_interpolate_parity(built, cp_level, deform):
    from msm_newmeshreg_b200 import Context
    cxyz0, ctris = pyorc.icosphere(cp_level)
    _, _, mvd = pyorc.spacings(cxyz0, ctris)
    start = pb.smooth_warp(cxyz0, 3, deform * mvd) if deform else cxyz0          # deformed START grid (not an icosphere)
    end = pb.smooth_warp(cxyz0, 4, 0.4 * mvd)
    pts, _ = pyorc.icosphere(5)
    pts = pb.smooth_warp(pts, 5, 0.5)
    ctx = Context()
    out, tri, w = ctx.mesh_interpolate(pts, start, ctris, end, details=True)
    ref = pyorc.mesh_interpolate(pts, start, ctris, end)
    assert np.abs(out - ref).max() < 1e-9
    assert np.allclose(w.sum(1), 1.0) and (w > -1e-12).all()
    to, _ = pyorc.locate(start, ctris, pts)
    interior = w.min(1) > 1e-9
    assert np.array_equal(tri[interior], to[interior])
    ctx.close()


@pytest.mark.gpu
def test_mesh_interpolate_accepts_either_winding(built):
    """msm_mesh_interpolate promises "any closed spherical triangle mesh": a mesh whose faces are wound inward as a whole (two
    corners of every face swapped) locates the same triangles and gives the same result as the outward-wound mesh.  (Single
    flipped faces of a FOLDED mesh are a different matter: they keep losing to their neighbours, as in the oracle's locator —
    the groupwise producers rely on that, tests/test_gpu_group.py.)"""
    from msm_newmeshreg_b200 import Context
    cxyz0, ctris = pyorc.icosphere(3)
    _, _, mvd = pyorc.spacings(cxyz0, ctris)
    start = pb.smooth_warp(cxyz0, 3, 0.2 * mvd)
    end = pb.smooth_warp(cxyz0, 4, 0.4 * mvd)
    pts, _ = pyorc.icosphere(4)
    pts = pb.smooth_warp(pts, 5, 0.5)
    ctx = Context()
    out, tri, w = ctx.mesh_interpolate(pts, start, ctris, end, details=True)
    flipped = ctris[:, [0, 2, 1]].copy()
    o2, t2, w2 = ctx.mesh_interpolate(pts, start, flipped, end, details=True)
    interior = w.min(1) > 1e-9
    assert np.array_equal(t2[interior], tri[interior])
    assert np.abs(o2 - out).max() < 1e-9
    assert np.abs(w2[:, [0, 2, 1]] - w).max() < 1e-9          # weights follow the face's own vertex order
    ctx.close()


@pytest.mark.gpu
def test_host_warp_functions(built):
    from msm_newmeshreg_b200 import host
    cxyz0, ctris = pyorc.icosphere(3)
    end = pb.smooth_warp(cxyz0, 4, 2.0)
    pts, ptris = pyorc.icosphere(4)
    moved = host.mesh_interpolate(pts, cxyz0, ctris, end)
    assert np.abs(moved - pyorc.mesh_interpolate(pts, cxyz0, ctris, end)).max() < 1e-9
    bad = pts.copy(); v = 300
    nb = np.unique(ptris[(ptris == v).any(1)]); nb = nb[nb != v]
    bad[v] = pts[nb[0]] + 1.2 * (pts[nb[0]] - pts[v]); bad[v] *= RAD / np.linalg.norm(bad[v])
    fixed_o, _ = pyorc.unfold(bad, ptris)
    assert np.abs(host.unfold(bad, ptris) - fixed_o).max() < 1e-12           # same sequential algorithm on the host


@pytest.mark.gpu
@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libref_optim.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("C", [1, 4])
def test_multiresolution_registration_matches_oracle(built, C):
    """run_multiresolutions restated (tests/level_walltime.py): two levels (control grids ico2 -> ico3 on ico5 data), the
    registration of level 1 carried to level 2 through project_CPgrid; labellings identical at every iteration of every
    level and final warped sphere within 1e-4.  The full BASELINE configs[1] run (ico6, 3 levels, C=4) is recorded in
    profiles/r01_levels_C2.json (labellings identical, warp difference 1.2e-12)."""
    import level_walltime as lw
    P = pb.make_problem(5, 2, C, seed=51, warp=0.0)
    P.seed = 51
    lam = pb.f32(0.1)
    src_gpu, labs_gpu, report = lw.run_gpu(P, [2, 3], 3, C, 4, lam)
    src_orc, labs_orc = lw.run_oracle(P, [2, 3], 3, C, 4, lam)
    for lvl, (g, o) in enumerate(zip(labs_gpu, labs_orc)):
        assert g.shape == o.shape and np.array_equal(g, o), f"level {lvl + 1}: {(g != o).sum()} labels differ"
    assert np.abs(src_gpu - src_orc).max() < 1e-4
    assert np.linalg.norm(src_gpu - P.txyz, axis=1).max() > 1.0       # the registration actually moved the sphere
    assert all(r["iterations"] >= 2 for r in report)


@pytest.mark.gpu
@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libref_optim.so not built (needs /root/reference at build time)")
def test_run_discrete_opt_matches_oracle_loop(built):
    """Whole outer loop, 3 iterations, HOCR + strain regulariser: labellings identical per iteration, final warped
    source vertices within 1e-4 of the oracle-assembled loop."""
    from msm_newmeshreg_b200 import host
    P = pb.make_problem(4, 2, 1, seed=41, warp=0.0)               # level start: source mesh == template (iteration 1)
    pr = dict(pb.DEFAULT_PARAMS); pr.update(rmode=3, dopt="HOCR")
    m = host.Model(sgres=P.cp_level + 2, cpres=P.cp_level, multivariate=False, _lib_override=pyref.lib(), **pr)
    m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref)
    m.initialize()
    iters = 3
    src_gpu, labs_gpu, en_gpu, _ = pyref.run_discrete_opt(m, P.txyz, iters, threads=4)

    # the same loop from oracle pieces (src/mesh_registration.cpp:320-396)
    src = P.txyz.copy(); cp = P.cxyz0.copy()
    energy = 0.0
    labs_orc = []
    for it in range(1, iters + 1):
        Q = pb.make_problem(4, 2, 1, seed=41, warp=0.0, iteration=it)
        Q.sxyz = src; Q.cxyz = cp; Q.rot = pyorc.rotations(Q.centre, cp)
        cf = pb.oracle_costfunction(Q, params=dict(rmode=3)); cf.get_source_data()
        lab, newenergy = pyref.fusion_tables(cf.unary_costs(), Q.triplets, cf.triplet_table(), threads=4)
        labs_orc.append(lab)
        if it > 1 and (it - 1) % 2 == 0 and energy - newenergy < 0.001:
            break
        newcp = np.einsum("nij,nj->ni", Q.rot, Q.labels[lab])
        src = pyorc.mesh_interpolate(src, cp, Q.ctris, newcp)
        newcp, _ = pyorc.unfold(newcp, Q.ctris)
        cp = newcp
        src, _ = pyorc.unfold(src, P.ttris)
        energy = newenergy
    assert len(labs_gpu) == len(labs_orc)
    sing = pb.singular_nodes(P)
    for it, (a, b) in enumerate(zip(labs_gpu, labs_orc)):
        keep = np.ones(P.N, bool); keep[sing] = False
        assert np.array_equal(a[keep], b[keep]), f"iteration {it + 1}: {(a != b).sum()} labels differ"
    # vertices influenced by the singular control point (antipode of the label-grid centre) are excluded
    d_sing = np.linalg.norm(P.txyz - P.cxyz0[sing[0]], axis=1) if len(sing) else np.full(len(P.txyz), 1e9)
    far = d_sing > 2.5 * P.mvdmax
    assert np.abs(src_gpu - src)[far].max() < 1e-4
    m.close()


@pytest.mark.gpu
def test_large_patches_from_concurrent_host_threads(built):
    """Patches of ~1000 samples need more than 48 KB of dynamic shared memory, an attribute CUDA keeps per function and
    device: contexts driven from different threads serialise set-attribute + launch (csrc/msm_b200.cu BigSmemLock).
    Last test of the GPU suite on purpose — it was written after the round's GPU budget was spent."""
    from concurrent.futures import ThreadPoolExecutor
    # ico6 data / ico2 control grid (configs[0]): 162 control points x 5 label groups of 4 labels, ~57 KB per CTA
    probs = [pb.make_problem(6, 2, 1, seed=60 + s, warp=0.3) for s in range(3)]

    def run(P):
        ctx = pb.gpu_context(P); ctx.build_patches()
        out = (ctx.unary_table(), ctx.patches()[1])
        ctx.close()
        return out

    serial = [run(P) for P in probs]
    assert max(len(s[1]) for s in serial) // probs[0].N > 700      # large patches: the > 48 KB path
    for _ in range(3):
        with ThreadPoolExecutor(3) as pool:
            threaded = list(pool.map(run, probs))
        for a, b in zip(serial, threaded):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.gpu
@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libref_optim.so not built (needs /root/reference at build time)")
def test_run_discrete_opt_with_the_table_aware_driver_gives_the_same_registration(built, monkeypatch):
    """MSM_HOCR_DRIVER=tables swaps Fusion::optimize for host/src/FusionTables.h (batched GPU move tables, flat binary MRF
    in front of the reference's unchanged ELC / FastPD cores) inside the run_discrete_opt loop: same labellings at every
    outer iteration, same warped sphere."""
    from msm_newmeshreg_b200 import host
    P = pb.make_problem(5, 3, 1, seed=33, warp=0.0)
    out = []
    for drv in ("fusion", "tables"):
        monkeypatch.setenv("MSM_HOCR_DRIVER", drv)
        pr = dict(pb.DEFAULT_PARAMS); pr.update(rmode=3, dopt="HOCR")
        m = host.Model(sgres=5, cpres=3, multivariate=False, _lib_override=pyref.lib(), threads=4, **pr)
        m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref); m.initialize()
        src, labs, en, tm = pyref.run_discrete_opt(m, P.txyz, 3, threads=4)
        out.append((src, labs, en)); m.close()
    assert out[0][1].shape == out[1][1].shape and np.array_equal(out[0][1], out[1][1])
    assert np.abs(out[0][0] - out[1][0]).max() < 1e-9 and np.allclose(out[0][2], out[1][2], rtol=1e-6)
