"""CPU-side tests: the C ABI library loads and exports every symbol declared in include/msm_b200.h, the product fails
loudly without a GPU, the oracle reproduces the committed golden fixtures, and the multi-rank sharding logic works
over gloo with world_size 2."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest

import problems as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_capi_exports_every_declared_symbol(built):
    import msm_newmeshreg_b200 as m
    hdr = open(os.path.join(ROOT, "include", "msm_b200.h")).read()
    declared = set(re.findall(r"\b(msm_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(m.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"symbols declared in include/msm_b200.h but not exported: {missing}"
    assert declared == set(m.EXPORTS), (declared ^ set(m.EXPORTS))


def test_struct_layout_matches_header(built):
    import msm_newmeshreg_b200 as m
    p = m.MsmParams()
    m.lib().msm_default_params(ctypes.byref(p))
    assert (p.lambda_, p.rexp, p.k_exp, p.range, p.simmeasure, p.rmode, p.group) == (1.0, 2.0, 2.0, 1.0, 2, 1, 0)
    assert abs(p.mu - 0.4) < 1e-7 and abs(p.kappa - 1.6) < 1e-7   # C floats in the reference (DiscreteCostFunction.h:121-122)


@pytest.mark.skipif(_have_cuda(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(built):
    import msm_newmeshreg_b200 as m
    with pytest.raises(m.MsmError, match="no CPU fallback"):
        m.Context()


def test_missing_library_fails_loudly(built, monkeypatch):
    import msm_newmeshreg_b200._capi as capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/libmsm_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.lib()


def test_oracle_reproduces_golden_fixtures():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "small_tables.npz"))
    for name in mg.CASES:
        _, out = mg.tables(name)
        for k, v in out.items():
            g = gold[f"{name}/{k}"]
            if v.dtype.kind == "i":
                assert np.array_equal(v, g), (name, k)
            else:
                assert np.allclose(v, g, rtol=1e-10, atol=1e-13), (name, k, np.abs(v - g).max())


def test_shard_ranges_cover_and_balance():
    from msm_newmeshreg_b200 import sharding as sh
    for n in (0, 1, 7, 162, 10242):
        for w in (1, 2, 3, 4, 8):
            r = [sh.shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert sh.subjects_of_rank(64, 3, 8) == list(range(3, 64, 8))
    with pytest.raises(ValueError):
        sh.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _gloo_worker(rank, world, port, n, L, q):
    import torch
    import torch.distributed as dist
    from msm_newmeshreg_b200 import sharding as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(L * n, dtype=torch.float32).reshape(L, n)   # the table a single GPU would produce
        b, e = sh.shard_range(n, rank, world)
        local = full[:, b:e].contiguous()                               # this rank's slice (label-major)
        got = sh.allgather_label_major(local, n)
        ok = bool(torch.equal(got, full))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ok, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_allgather_label_major_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n, L, world = 163, 19, 2            # odd N: unequal shards exercise the padding
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, L, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(t == float(world) for _, _, t in res)


def _gloo_subject_worker(rank, world, port, S, shape, q):
    import torch
    import torch.distributed as dist
    from msm_newmeshreg_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s0, s1 = sharding.shard_range(S, rank, world)
    full = torch.arange(S * int(np.prod(shape)), dtype=torch.float64).reshape((S,) + shape)
    got = sharding.allgather_subject_rows(full[s0:s1].clone(), S)
    q.put((rank, bool(torch.equal(got, full)), tuple(got.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("S", [4, 3, 1])
def test_allgather_subject_rows_gloo_world2(S):
    """Groupwise data term over 2 ranks: contiguous subject shards, uneven (S = 3) and with an empty shard (S = 1),
    come back as [S][L][V] in subject order from ONE all-gather."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29621 + S
    procs = [ctx.Process(target=_gloo_subject_worker, args=(r, world, port, S, (3, 5), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res) and all(shape == (S, 3, 5) for _, _, shape in res)


def test_assemble_label_major_numpy():
    from msm_newmeshreg_b200 import sharding as sh
    full = np.arange(19 * 10).reshape(19, 10)
    slices = [full[:, slice(*sh.shard_range(10, r, 4))] for r in range(4)]
    assert np.array_equal(sh.assemble_label_major(slices, 10), full)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: stdout carries ONE JSON line (library banners go to stderr); the reference arm's line has the
    keys the driver reads.  Tiny sample so it runs in seconds."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--data-level", "5", "--cp-level", "3", "--cpu-rows", "32"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True
    for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert "workload" in d["config"]
    assert d["value_tables_only"] >= d["value"]                        # like for like with the GPU arm: value includes the patch search
    if d["cpu_baseline"]["kind"] == "reference":                        # one whole resolution level through the reference itself
        lw = d["level_wall"]
        assert "error" not in lw, lw
        assert lw["arm"] == "reference" and lw["level_wall_s"] > 0 and lw["N"] == 642


# ------------------------------------------------------------------------------------------------ hierarchy builder (CPU)
@pytest.fixture(scope="module")
def ich():
    """tests/cpu_hierarchy.cpp: the product's host-side hierarchy builder (csrc/ico_hierarchy.h) behind a small C ABI."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    out_dir = os.path.join(here, "_build"); os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcpu_hierarchy.so")
    src = os.path.join(here, "cpu_hierarchy.cpp")
    hdr = os.path.join(here, "..", "msm_newmeshreg_b200", "csrc", "ico_hierarchy.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so])
    L = ctypes.CDLL(so)
    L.ich_error.restype = ctypes.c_char_p
    return L


def _ich_build(L, xyz, tris):
    xyz = np.ascontiguousarray(xyz, np.float64); tris = np.ascontiguousarray(tris, np.int32)
    return L.ich_build(xyz.ctypes.data_as(ctypes.c_void_p), len(xyz), tris.ctypes.data_as(ctypes.c_void_p), len(tris))


def _sphere_points(n, seed):
    p = np.random.default_rng(seed).normal(size=(n, 3))
    return 100.0 * p / np.linalg.norm(p, axis=1, keepdims=True)


@pytest.mark.parametrize("level", [0, 1, 3, 5])
def test_hierarchy_descent_equals_brute_force_closest_triangle(ich, level):
    """Base face + `level` scalar four-way steps over the tables ico_hierarchy.h builds find the triangle the oracle's
    brute-force search finds, with its barycentric weights — for any vertex / face numbering and orientation."""
    from oracle import pyorc
    xyz, tris = pyorc.icosphere(level)
    rng = np.random.default_rng(level)
    th = 0.9
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]]) @ \
        np.array([[np.cos(0.3), -np.sin(0.3), 0], [np.sin(0.3), np.cos(0.3), 0], [0, 0, 1]])
    vp = rng.permutation(len(xyz)); inv = np.argsort(vp)
    xyz2 = (xyz @ R.T)[vp]
    tris2 = inv[tris][rng.permutation(len(tris))]
    tris2 = np.stack([np.roll(t, rng.integers(0, 3)) for t in tris2]).astype(np.int32)
    assert _ich_build(ich, xyz2, tris2) == 0, ich.ich_error()
    assert ich.ich_depth() == level
    pts = _sphere_points(20000, 100 + level)
    tri = np.empty(len(pts), np.int32); w = np.empty((len(pts), 3))
    ich.ich_locate(pts.ctypes.data_as(ctypes.c_void_p), len(pts), tri.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p))
    bt, bw = pyorc.locate(xyz2, tris2, pts, brute=True)
    inside = w.min(1) > 1e-9                                  # a point on an edge belongs to either triangle
    assert inside.mean() > 0.99
    assert np.array_equal(tri[inside], bt[inside])
    assert np.abs(w - bw)[inside].max() < 1e-12
    assert np.abs(w.sum(1) - 1).max() < 1e-14 and w.min() >= 0


@pytest.mark.parametrize("h", [0, 1, 2, 3])
def test_hint_frame_walk_reaches_the_same_triangle(ich, h):
    """k_unary's route: coordinates in the frame of the level-h face under ANOTHER point (the control point), edge walk
    with the hT tables, descent of the remaining levels — same triangle and weights as the direct descent."""
    from oracle import pyorc
    level = 4
    xyz, tris = pyorc.icosphere(level)
    assert _ich_build(ich, xyz, tris) == 0, ich.ich_error()
    pts = _sphere_points(20000, 7)
    rng = np.random.default_rng(8)
    edge_h = 100.0 * 1.1071487177940904 / 2 ** h             # arc length of a level-h edge
    d = rng.normal(size=pts.shape) * (0.8 * edge_h / np.sqrt(3))
    starts = pts + d; starts *= 100.0 / np.linalg.norm(starts, axis=1, keepdims=True)
    tri = np.empty(len(pts), np.int32); w = np.empty((len(pts), 3))
    tri2 = np.empty(len(pts), np.int32); w2 = np.empty((len(pts), 3)); steps = np.empty(len(pts), np.int32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ich.ich_locate(P(pts), len(pts), P(tri), P(w))
    assert ich.ich_locate_via_hint(h, P(starts), P(pts), len(pts), P(tri2), P(w2), P(steps)) == 0, ich.ich_error()
    inside = (w.min(1) > 1e-9) & (w2.min(1) > 1e-9)
    assert inside.mean() > 0.99
    assert np.array_equal(tri[inside], tri2[inside])
    assert np.abs(w - w2)[inside].max() < 1e-11
    assert steps.max() < 16 and steps.max() > 0               # some points did cross edges


def test_hierarchy_builder_rejects_what_is_not_a_regular_icosphere(ich):
    from oracle import pyorc
    xyz, tris = pyorc.icosphere(3)
    assert _ich_build(ich, pb.smooth_warp(xyz, 1, 2.0), tris) != 0 and re.search(b"icos|sphere", ich.ich_error())
    assert _ich_build(ich, xyz[:-1], tris[:-2]) != 0
    bad = tris.copy(); bad[5] = bad[6]                         # a face twice, another missing
    assert _ich_build(ich, xyz, bad) != 0
    assert _ich_build(ich, xyz * np.array([1.0, 1.0, 1.02]), tris) != 0 and re.search(b"sphere", ich.ich_error())


def test_reference_driver_loop_compiles_against_the_mirror():
    """Compile-level drop-in proof (no GPU needed): oracle/_ref/libref_optim.so is host/harness.cpp built over the MIRROR headers
    together with the unmodified text of Mesh_registration::run_discrete_opt / combine_costfunction_weighting
    (src/mesh_registration.cpp:316-397,888-909), src/mcmc_opt.h and the vendored Fusion / FastPD / ELC headers — if any name,
    signature or type of the mirror's model surface drifted from the reference's, that library would not build."""
    import ctypes
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_optim.so")
    if not os.path.exists(so):
        import pytest
        pytest.skip("oracle/_ref/libref_optim.so not built (needs /root/reference at build time)")
    lib = ctypes.CDLL(so)
    for name in ("ref_driver_run_discrete_opt", "ref_run_discrete_opt", "ref_fusion_model", "ref_fastpd_model", "msmhost_featurespace_initialize"):
        assert hasattr(lib, name), name


def test_bench_numa_binding_degrades_gracefully():
    """bench.py binds multi-GPU ranks to the cores of their GPU's NUMA node; without NVML / without a second node it must say
    so and leave the process unbound."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    before = os.sched_getaffinity(0)
    r = bench.bind_to_gpu_numa_node(0, 2)
    assert r["bound"] is False and "why" in r
    assert os.sched_getaffinity(0) == before
    os.environ["MSM_BENCH_BIND"] = "0"
    try:
        assert bench.bind_to_gpu_numa_node(0, 2) == {"bound": False, "why": "MSM_BENCH_BIND=0"}
    finally:
        os.environ.pop("MSM_BENCH_BIND")


# ------------------------------------------------------------------------------------------------ mesh container of the mirror (CPU)
@pytest.fixture(scope="module")
def shim():
    """tests/cpu_shim.cpp: host/src/newresampler_shim.h behind a small C ABI (header-only, no GPU library needed)."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    out_dir = os.path.join(here, "_build"); os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcpu_shim.so")
    src = os.path.join(here, "cpu_shim.cpp")
    hdr = os.path.join(here, "..", "msm_newmeshreg_b200", "host", "src", "newresampler_shim.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so])
    return ctypes.CDLL(so)


@pytest.mark.parametrize("level", [0, 2, 4])
def test_mirror_icosphere_and_lazy_adjacency_match_the_oracle(shim, level):
    """make_mesh_from_icosa of the mirror (cached per level) = the oracle's ([EXT] A7 numbering); the adjacency the mirror builds on
    first use has the oracle's order — faces ascending, neighbours by first appearance — which label_sampling_grid's
    breadth-first sweep (src/DiscreteModel.cpp:105-162) depends on."""
    from oracle import pyorc
    ox, ot = pyorc.icosphere(level, 1.0)
    nv, nf = ctypes.c_int(), ctypes.c_int()
    shim.shim_icosphere(level, ctypes.byref(nv), ctypes.byref(nf), None, None)
    assert (nv.value, nf.value) == (len(ox), len(ot))
    xyz = np.empty((nv.value, 3)); tris = np.empty((nf.value, 3), np.int32)
    shim.shim_icosphere(level, ctypes.byref(nv), ctypes.byref(nf), xyz.ctypes.data_as(ctypes.c_void_p), tris.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(tris, ot) and np.abs(xyz - ox).max() < 1e-15
    noff = np.empty(len(xyz) + 1, np.int32); toff = np.empty(len(xyz) + 1, np.int32)
    nbr = np.empty(6 * len(xyz), np.int32); tri = np.empty(6 * len(xyz), np.int32)
    shim.shim_adjacency(xyz.ctypes.data_as(ctypes.c_void_p), len(xyz), tris.ctypes.data_as(ctypes.c_void_p), len(tris),
                        noff.ctypes.data_as(ctypes.c_void_p), nbr.ctypes.data_as(ctypes.c_void_p), toff.ctypes.data_as(ctypes.c_void_p),
                        tri.ctypes.data_as(ctypes.c_void_p))
    for v in range(len(xyz)):
        inc = [t for t in range(len(tris)) if v in tris[t]] if level <= 2 else None
        if inc is not None:
            assert list(tri[toff[v]:toff[v + 1]]) == inc
            seen = []
            for t in inc:
                for u in tris[t]:
                    if u != v and u not in seen:
                        seen.append(int(u))
            assert list(nbr[noff[v]:noff[v + 1]]) == seen
    deg = np.diff(noff)
    assert (deg == 5).sum() == 12 and set(np.unique(deg)) <= {5, 6} and np.array_equal(deg, np.diff(toff))


def test_mirror_mesh_copies_are_independent(shim):
    assert shim.shim_semantics(3) == 0
