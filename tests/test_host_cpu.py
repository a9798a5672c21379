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
