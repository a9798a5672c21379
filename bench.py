#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native newMSM cost-evaluation path.

Metric (BASELINE.json): unary + triplet cost evaluations per second (one evaluation = one control point x label unary
cost, or one triplet x label-triple cost).

Workload (config.workload): C3 of BASELINE.json — ico7 data mesh (163 842 vertices) x ico5 control grid (10 242 control
points, 20 480 triplets), L = 19 labels, range 1, univariate NCC/Pearson.  One "step" is one outer iteration's cost
evaluation for B independent subject->template registrations per GPU: the whole unary table (N*L evaluations) plus the
fusion-move triplet costs of every label (L*T*8 evaluations) per subject.  Subjects are independent, so N GPUs process
N*B subjects with no data-path collective ("weak" scaling, SURVEY.md §8e row "batched").  B subjects x ~80 MB of device
state exceed the 126 MB L2, so successive timed iterations do not find their inputs in L2.

  value : evaluations / second, inputs resident in HBM, timed with CUDA events on the launching stream, max over ranks.
          The step is LIKE FOR LIKE with the reference arm: patch construction (get_source_data) + the whole unary table +
          the all-label fusion-move triplet costs.  value_tables_only (both arms) leaves the patch construction out.
  e2e   : the same step through the reference-facing host classes (NonLinearSRegDiscreteModel::reset_meshspace ->
          setupCostFunction [patch construction + unary table] -> triplet moves) with pinned HOST buffers in and out;
          the B independent subjects of a GPU are driven by --e2e-threads host threads, one private stream each
          (e2e.value_one_host_thread = the same with a serial loop).
  --impl reference : THE REFERENCE'S OWN cost-function classes (oracle/_ref/libref_cost.so: its sources compiled in place
          over stand-ins for the absent FSL libraries) on the host cores, bounded sample of the same workload; falls back
          to the oracle port (kind "port") only if that library was not built.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0
RAD = 100.0


def f32(x):
    return float(np.float32(x))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- the one stdout line
_RESULT_FD = None


def capture_stdout():
    """stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner, for one) are sent to
    stderr for the whole run, the result line goes to the original descriptor."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit_line(text):
    sys.stdout.flush()
    if _RESULT_FD is None:
        print(text, flush=True)
    else:
        os.write(_RESULT_FD, (text + "\n").encode())


# ---------------------------------------------------------------------------------------------- synthetic data
def random_field(xyz, seed, nwaves=32, kmax=6.0):
    rng = np.random.default_rng(seed)
    u = xyz / RAD
    k = rng.normal(size=(nwaves, 3))
    k *= (rng.uniform(1.0, kmax, size=(nwaves, 1)) / np.linalg.norm(k, axis=1, keepdims=True))
    f = (rng.normal(size=nwaves)[None, :] * np.cos(u @ k.T + rng.uniform(0, 2 * np.pi, size=nwaves)[None, :])).sum(1)
    return (f - f.mean()) / f.std()


def smooth_warp(xyz, seed, amplitude):
    rng = np.random.default_rng(seed)
    u = xyz / RAD
    d = np.zeros_like(u)
    for _ in range(6):
        k = rng.normal(size=3) * 2.0
        d += rng.normal(size=3)[None, :] * np.sin(u @ k + rng.uniform(0, 2 * np.pi))[:, None]
    d -= (d * u).sum(1, keepdims=True) * u
    d *= amplitude / max(np.abs(d).max(), 1e-12)
    p = xyz + d
    return p * (RAD / np.linalg.norm(p, axis=1, keepdims=True))


def rotate(xyz, axis, deg):
    axis = np.asarray(axis, float); axis /= np.linalg.norm(axis)
    t = np.deg2rad(deg)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return xyz @ (np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * K @ K).T


PARAMS = dict(lam=f32(0.1), rexp=f32(2.0), mu=f32(0.4), kappa=f32(1.6), k_exp=f32(2.0), rng=f32(1.0), simmeasure=2, rmode=3)


class Subject:
    """Synthetic subject: reference field on the template icosphere, source field = the same random process seen through
    a known 2-degree rotation, source mesh = the icosphere pushed through a smooth warp."""

    def __init__(self, txyz, ttris, C, seed, cp_spacing):
        self.ref = np.stack([random_field(txyz, seed * 10 + c) for c in range(C)])
        moved = rotate(txyz, [0.3, -0.5, 0.8], 2.0)
        self.inp = np.stack([random_field(moved, seed * 10 + c) for c in range(C)])
        self.sxyz = smooth_warp(txyz, seed + 100, 0.3 * cp_spacing)


# ---------------------------------------------------------------------------------------------- NUMA placement of a rank
def bind_to_gpu_numa_node(index, world):
    """Multi-GPU runs: pin this rank's host threads to the CPU cores of its GPU's NUMA node (NVML's ideal CPU affinity) BEFORE any
    pinned buffer is allocated, so that staging buffers are first-touched on the node the GPU hangs off and the end-to-end leg's
    copies do not cross the socket interconnect.  MSM_BENCH_BIND=0 switches it off.  Returns a description for the bench line."""
    if os.environ.get("MSM_BENCH_BIND", "1") == "0":
        return {"bound": False, "why": "MSM_BENCH_BIND=0"}
    try:
        import pynvml
        pynvml.nvmlInit()
        ncpu = os.cpu_count() or 1
        allowed = os.sched_getaffinity(0)

        def cpus_of(i):
            words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(i), (ncpu + 63) // 64)
            return {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1} & allowed
        cpus = cpus_of(index)
        if not cpus or cpus == allowed:
            return {"bound": False, "why": "one NUMA node / no affinity information", "cpus": len(allowed)}
        sharing = sum(1 for i in range(world) if cpus_of(i) == cpus)     # ranks whose GPUs hang off the same node share its cores
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "cpus": len(cpus), "of": len(allowed), "ranks_on_node": max(1, sharing)}
    except Exception as e:   # no NVML, no permission: run unbound
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


# ---------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True); self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU (oracle) arm
def oracle_sample(data_level, cp_level, C, rows, threads, seed=1000, steps=1, warmup=0):
    """The oracle's restatement of the reference path on a bounded sample: patch construction (single thread, as in the
    reference: src/DiscreteCostFunction.cpp:394-410 has no OpenMP), unary table (labels outer, omp over control points,
    :306-313) for `rows` control points, and the fusion-move triplet costs of every label for the proportional share of
    triplets (omp over triplets, include/Fusion/Fusion.h:180)."""
    from oracle import pyorc
    txyz, ttris = pyorc.icosphere(data_level)
    cxyz0, ctris = pyorc.icosphere(cp_level)
    maxsep, mvdmax, mvd = pyorc.spacings(cxyz0, ctris)
    S = Subject(txyz, ttris, C, seed, mvd)
    samples, bary, centre = pyorc.label_grid(cp_level + 2, 0.5 * mvdmax)
    cxyz = smooth_warp(cxyz0, seed + 200, 0.1 * mvd)
    trip = pyorc.triplets(cxyz0, ctris)
    N, L, T = len(cxyz), len(bary), len(trip)
    rows = min(rows, N)
    tsel = np.arange(max(1, T * rows // N))
    cf = pyorc.CostFunction()
    cf.set_params(threads=threads, **PARAMS)
    cf.set_meshes(txyz, ttris, S.sxyz, cxyz0, ctris)
    cf.reset_cpgrid(cxyz); cf.set_orig(cxyz0)
    cf.set_features(S.ref, S.inp, C > 1)
    cf.set_spacings(maxsep, mvdmax)
    cf.set_labels(bary, pyorc.rotations(centre, cxyz))
    cf.set_triplets(trip)
    lab = np.random.default_rng(7).integers(0, L, N).astype(np.int32)
    q = np.empty((L, len(tsel), 8, 4), np.int32)
    for a in range(L):
        for combo in range(8):
            q[a, :, combo, 0] = tsel
            q[a, :, combo, 1] = np.where(combo & 4, a, lab[trip[tsel, 0]])
            q[a, :, combo, 2] = np.where(combo & 2, a, lab[trip[tsel, 1]])
            q[a, :, combo, 3] = np.where(combo & 1, a, lab[trip[tsel, 2]])
    q = q.reshape(-1, 4)
    evals = rows * L + len(q)
    times = {"patches": [], "unary": [], "triplet": []}
    for it in range(warmup + steps):
        cf.set_params(threads=1, **PARAMS)
        t0 = time.perf_counter(); cf.get_source_data(0, rows); t1 = time.perf_counter()
        cf.set_params(threads=threads, **PARAMS)
        cf.unary_costs(0, rows); t2 = time.perf_counter()
        cf.triplet_costs(q); t3 = time.perf_counter()
        if it >= warmup:
            times["patches"].append(t1 - t0); times["unary"].append(t2 - t1); times["triplet"].append(t3 - t2)
    tp, tu, tt = (float(np.mean(times[k])) for k in ("patches", "unary", "triplet"))
    return dict(evals=evals, rows=rows, N=N, L=L, T=T, ntrip=len(tsel), t_patches=tp, t_unary=tu, t_triplet=tt,
                evals_per_s_resident=evals / (tu + tt), evals_per_s_e2e=evals / (tp + tu + tt))


def reference_sample(data_level, cp_level, C, rows, threads, seed=1000, steps=1, warmup=0, tables=None):
    """The same bounded sample through THE REFERENCE'S OWN cost-function classes (oracle/_ref/libref_cost.so = its
    src/DiscreteCostFunction.cpp, similarities.cc, reg_tools.cpp compiled in place; host/harness.cpp ref_cost_sample).
    get_source_data + computeUnaryCosts run on ONE thread (the reference's threaded unary loop is a data race, DESIGN.md
    §1), the re-entrant computeTripletCost on `threads` OpenMP threads as in include/Fusion/Fusion.h:180."""
    import ctypes as Ct
    from oracle import pyorc, pyrefcost
    L_ = pyrefcost.lib()
    txyz, ttris = pyorc.icosphere(data_level)
    cxyz0, ctris = pyorc.icosphere(cp_level)
    maxsep, mvdmax, mvd = pyorc.spacings(cxyz0, ctris)
    S = Subject(txyz, ttris, C, seed, mvd)
    samples, bary, centre = pyorc.label_grid(cp_level + 2, 0.5 * mvdmax)
    cxyz = smooth_warp(cxyz0, seed + 200, 0.1 * mvd)
    trip = pyorc.triplets(cxyz0, ctris)
    N, L, T = len(cxyz), len(bary), len(trip)
    rows = min(rows, N)
    ntrip = max(1, T * rows // N)
    lab = np.random.default_rng(7).integers(0, L, N).astype(np.int32)
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    dp, ip = Ct.POINTER(Ct.c_double), Ct.POINTER(Ct.c_int)
    A = dict(txyz=d(txyz), ttris=i(ttris), sxyz=d(S.sxyz), inp=d(S.inp), ref=d(S.ref), cp0=d(cxyz0), cp=d(cxyz), ctris=i(ctris),
             maxsep=d(maxsep), labels=d(bary), trip=i(trip[:ntrip]), lab=i(lab))
    P = {k: Ct.c_double(PARAMS[k]) for k in ("lam", "rexp", "mu", "kappa", "k_exp", "rng")}
    sec = np.zeros(3)
    times = {"patches": [], "unary": [], "triplet": []}
    u_out = m_out = None
    if tables is not None:   # tests: hand the evaluated sample back
        tables["unary"] = np.zeros((L, rows)); tables["moves"] = np.zeros((L, ntrip, 8)); tables["labeling"] = lab
        u_out, m_out = tables["unary"].ctypes.data_as(dp), tables["moves"].ctypes.data_as(dp)
    for it in range(warmup + steps):
        st = L_.ref_cost_sample(A["txyz"].ctypes.data_as(dp), len(txyz), A["ttris"].ctypes.data_as(ip), len(ttris), A["sxyz"].ctypes.data_as(dp),
                                A["inp"].ctypes.data_as(dp), A["ref"].ctypes.data_as(dp), C, int(C > 1), A["cp0"].ctypes.data_as(dp),
                                A["cp"].ctypes.data_as(dp), N, A["ctris"].ctypes.data_as(ip), len(ctris), A["maxsep"].ctypes.data_as(dp),
                                Ct.c_double(mvdmax), A["labels"].ctypes.data_as(dp), L, P["lam"], P["rexp"], P["mu"], P["kappa"], P["k_exp"],
                                P["rng"], int(PARAMS["simmeasure"]), int(PARAMS["rmode"]), rows, ntrip, A["trip"].ctypes.data_as(ip),
                                A["lab"].ctypes.data_as(ip), threads, u_out, m_out, sec.ctypes.data_as(dp))
        if st != 0:
            raise RuntimeError(L_.msmhost_last_error().decode())
        if it >= warmup:
            times["patches"].append(sec[0]); times["unary"].append(sec[1]); times["triplet"].append(sec[2])
    tp, tu, tt = (float(np.mean(times[k])) for k in ("patches", "unary", "triplet"))
    evals = rows * L + L * ntrip * 8
    return dict(evals=evals, rows=rows, N=N, L=L, T=T, ntrip=ntrip, t_patches=tp, t_unary=tu, t_triplet=tt,
                evals_per_s_resident=evals / (tu + tt), evals_per_s_e2e=evals / (tp + tu + tt))


def cpu_sample(data_level, cp_level, C, rows, threads, steps=1, warmup=0):
    """CPU arm: the reference's own classes when oracle/_ref/libref_cost.so was built (kind "reference"), else the
    oracle port (kind "port").  Returns (result dict, kind, description of the sample)."""
    from oracle import pyrefcost
    if pyrefcost.available():
        r = reference_sample(data_level, cp_level, C, rows, threads, steps=steps, warmup=warmup)
        kind = "reference"
        how = (f"reference's own DiscreteCostFunction classes compiled in place (oracle/_ref/libref_cost.so): get_source_data + "
               f"computeUnaryCosts on 1 thread (more is a data race in the reference), computeTripletCost on {threads} threads")
    else:
        r = oracle_sample(data_level, cp_level, C, rows, threads, steps=steps, warmup=warmup)
        kind = "port"
        how = f"oracle port (oracle/_ref/libref_cost.so not built): patch construction 1 thread as in the reference, tables {threads} threads"
    sample = (f"{r['rows']} of {r['N']} control points + {r['ntrip']} of {r['T']} triplets x {r['L']} labels x 8 moves; {how}; "
              f"patches {r['t_patches']:.2f}s, unary {r['t_unary']:.2f}s, triplets {r['t_triplet']:.3f}s")
    return r, kind, sample


# ---------------------------------------------------------------------------------------------- per-level wall time
def level_wall(data_level, cp_level, C, arm, threads):
    """Second half of BASELINE's metric: wall time of ONE resolution level (one outer iteration: cost evaluation + the
    reference's HOCR optimiser + label application + source warp), restating Mesh_registration::run_discrete_opt
    (src/mesh_registration.cpp:316-397) as tests/level_walltime.py does.  arm "gpu": this repository's cost functions and
    host classes; the OPTIMISER is the reference's own unmodified Fusion / ELC / FastPD code in both arms (it stays on the
    host by north_star; compiled in place into oracle/_ref, which is why this figure is only produced where that exists).
    arm "reference": the reference's own cost functions as well, one host thread (more is a data race in it)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import level_walltime as lw
    import problems as pb
    P = pb.make_problem(data_level, cp_level, C, seed=51, warp=0.0)
    P.seed = 51
    run = lw.run_gpu if arm == "gpu" else lw.run_reference
    if arm == "gpu":   # the table-aware fusion driver (host/src/FusionTables.h): same labellings as Fusion::optimize, batched tables
        os.environ["MSM_HOCR_DRIVER"] = "tables"
    t0 = time.perf_counter()
    try:
        src, labs, rep = run(P, [cp_level], 1, C, threads, f32(0.1))
    finally:
        os.environ.pop("MSM_HOCR_DRIVER", None)
    r = rep[0]
    return {"level": f"ico{data_level} data / ico{cp_level} control grid, C={C}, strain regulariser, HOCR, 1 outer iteration", "arm": arm,
            "level_wall_s": r["level_wall_s"], "setup_s": r["setup_s"], "cost_eval_s": r["cost_eval_s"], "optimiser_s": r["optimiser_s"],
            "warp_s": r["warp_s"], "labels_changed": int((labs[0][0] != 0).sum()), "N": r["N"], "host_threads": threads if arm == "gpu" else 1,
            "total_s": time.perf_counter() - t0,
            "optimiser": ("FusionTables driver over the reference's unmodified ELC.h / FastPD.h cores (batched move tables, flat binary MRF)"
                          if arm == "gpu" else "reference Fusion.h / ELC.h / FastPD.h, unmodified")}


# ---------------------------------------------------------------------------------------------- feature pre-processing
def feature_preprocessing(arm, level=6, D=4):
    """featurespace::initialize (src/featurespace.cc:18-65) of one resolution level — data of an irregular ico`level`-sized mesh
    resampled onto the ico`level` template (adaptive barycentric), smoothed (sigma 5), histogram matched, variance normalised —
    arm "gpu": the product's featurespace (host/src/featurespace.h over csrc/features.cuh); arm "reference": the reference's own
    featurespace.cc + reg_tools.cpp compiled in place (oracle/_ref/libref_feat.so; the oracle port where that is absent).
    Wall clock from host arrays to host arrays, mesh containers included on both sides."""
    from msm_newmeshreg_b200 import host
    xyz, tris = host.icosphere(level)
    xi = smooth_warp(xyz, 5, 1.0)
    rng = np.random.default_rng(3)
    di = np.stack([2.0 + 1.5 * random_field(xi, 21 + d) + 0.3 * rng.normal(size=len(xi)) for d in range(D)])
    dr = np.stack([-1.0 + 0.7 * random_field(xyz, 71 + d) + 0.3 * rng.normal(size=len(xyz)) for d in range(D)])
    kw = dict(sigma=(5.0, 5.0), intensitynorm=True, varnorm=True)
    if arm == "gpu":
        fn, kind = host.featurespace_initialize, "b200"
        fn(level, (xi, tris), di, (xyz, tris), dr, **kw)          # first call: CUDA context of the feature path, icosphere cache
    else:
        from oracle import pyorc, pyreffeat
        fn, kind = (pyreffeat.featurespace_initialize, "reference") if pyreffeat.available() else (pyorc.featurespace_initialize, "port")
    ts = []
    for _ in range(3 if arm == "gpu" else 1):
        t0 = time.perf_counter(); out = fn(level, (xi, tris), di, (xyz, tris), dr, **kw); ts.append(time.perf_counter() - t0)
    return {"what": f"featurespace::initialize, ico{level} ({len(xyz)} vertices), {D} feature dimensions, sigma 5, histogram matching + "
                    f"variance normalisation", "arm": arm, "kind": kind, "seconds": min(ts),
            "checksum": float(np.abs(out[0]).sum() + np.abs(out[1]).sum())}


# ---------------------------------------------------------------------------------------------- CP-sharded mode
def bench_cp_sharded(args, rank, world, local, C, metric, unit, emit=True):
    """ONE HCP-scale registration (ico7 / ico5): control points and triplets split into contiguous ranges over the GPUs
    (SURVEY.md §8e rows 1-2), every GPU holds the replicated meshes, table slices gathered with one NCCL all_gather each
    per step (never inside a per-label loop).  Strong scaling: total work fixed."""
    import torch
    import torch.distributed as dist
    from msm_newmeshreg_b200 import host, sharding
    txyz, ttris = host.icosphere(args.data_level)
    cxyz0, ctris = host.icosphere(args.cp_level)
    mvd = float(np.linalg.norm(cxyz0[ctris[:, 0]] - cxyz0[ctris[:, 1]], axis=1).mean())
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    S = Subject(txyz, ttris, C, 1000, mvd)                       # the same subject on every rank (replicated meshes)
    m = host.Model(sgres=args.cp_level + 2, cpres=args.cp_level, multivariate=C > 1, dopt="HOCR", threads=1, **PARAMS)
    m.set_data(txyz, ttris, txyz, ttris, S.inp, S.ref)
    m.initialize()
    m.reset_meshspace(S.sxyz)
    m.reset_cpgrid(smooth_warp(cxyz0, 1200, 0.1 * mvd))
    m.setup()
    N, L, T = m.N, m.L, m.T
    ctx = m.context()
    ctx.set_stream(stream.cuda_stream)
    k0, k1 = sharding.shard_range(N, rank, world)
    t0, t1 = sharding.shard_range(T, rank, world)
    ctx.set_shard(k0, k1)
    ctx.build_patches()
    lab_dev = torch.from_numpy(np.random.default_rng(7).integers(0, L, N).astype(np.int32)).cuda()
    d_unary = torch.empty((L, k1 - k0), dtype=torch.float32, device="cuda")
    d_moves = torch.empty((L, (t1 - t0) * 8), dtype=torch.float32, device="cuda")

    def step_nccl():
        ctx.unary_table_dev(d_unary.data_ptr())
        ctx.triplet_moves_range_dev(lab_dev.data_ptr(), -1, t0, t1, d_moves.data_ptr())
        U = sharding.allgather_label_major(d_unary, N)
        M = sharding.allgather_label_major(d_moves, T, unit=8)
        return U, M

    # Fused exchange: the kernels store every finished entry straight into the full table of EVERY rank through
    # symmetric-memory peer mappings (NVLink stores) — no all-gather, no re-layout; one device-side barrier per step.
    # (A consumer that reads the tables between steps needs a second barrier, or two table sets, before the next step
    # may overwrite them; the bench only reads after the last step.)
    exchange, why = args.cp_exchange, ""
    full_u = full_m = hdl = None
    if exchange in ("p2p", "mc"):
        try:
            if world > 1:
                import torch.distributed._symmetric_memory as symm_mem
                full_u = symm_mem.empty(L * N, dtype=torch.float32, device=f"cuda:{local}")
                full_m = symm_mem.empty(L * T * 8, dtype=torch.float32, device=f"cuda:{local}")
                hdl = symm_mem.rendezvous(full_u, dist.group.WORLD)
                hdl_m = symm_mem.rendezvous(full_m, dist.group.WORLD)
                peer_u, peer_m = [int(p) for p in hdl.buffer_ptrs], [int(p) for p in hdl_m.buffer_ptrs]
                mc_u, mc_m = int(getattr(hdl, "multicast_ptr", 0) or 0), int(getattr(hdl_m, "multicast_ptr", 0) or 0)
                if exchange == "mc" and not (mc_u and mc_m):
                    exchange, why = "p2p", " (no NVSwitch multicast address on this box: plain peer stores)"
            else:
                full_u = torch.empty(L * N, dtype=torch.float32, device="cuda")
                full_m = torch.empty(L * T * 8, dtype=torch.float32, device="cuda")
                peer_u, peer_m = [full_u.data_ptr()], [full_m.data_ptr()]
                mc_u = mc_m = 0
                exchange = "p2p"
        except Exception as e:   # no symmetric memory on this box: fall back to the NCCL gather
            exchange, why = "nccl", f" (p2p unavailable: {type(e).__name__}: {e})"

    def step_p2p():
        ctx.unary_table_peers(peer_u, mc_u if exchange == "mc" else 0)
        ctx.triplet_moves_range_peers(lab_dev.data_ptr(), -1, t0, t1, peer_m, mc_m if exchange == "mc" else 0)
        if hdl is not None:
            hdl.barrier(channel=0)
        return full_u.view(L, N), full_m.view(L, T * 8)

    if exchange in ("p2p", "mc"):   # the exchanges must produce the same tables
        Un, Mn = step_nccl()
        Up, Mp = step_p2p()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        assert torch.equal(Un, Up) and torch.equal(Mn.reshape(L, -1), Mp), "fused exchange differs from the NCCL gather"
    step = step_p2p if exchange in ("p2p", "mc") else step_nccl

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    for _ in range(args.warmup):
        U, M = step()
    sync_all()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        U, M = step()
    e1.record(stream)
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    evals = N * L + L * T * 8
    checksum = float(U.double().sum().item()), float(M.double().sum().item())
    line = None
    if rank == 0:
        line = ({
            "metric": metric, "value": evals / (ms_per_step * 1e-3), "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (unary) / f64 (triplet)", "data": "synthetic",
            "config": {"workload": f"C3 CP-sharded: ONE ico{args.data_level}/ico{args.cp_level} registration, control points and triplets "
                                   f"split over {world} GPU(s), unary [L][N] and fusion-move [L][T][8] tables "
                                   + ("written by the kernels straight into every rank's table (fused exchange: "
                                      + ("ONE NVSwitch multicast store per entry" if exchange == "mc" else "one NVLink peer store per rank and entry")
                                      + ", one device-side barrier per step)" + why if exchange in ("p2p", "mc") else "gathered with NCCL all_gather" + why),
                       "exchange": exchange, "N": N, "L": L, "T": T, "gathered_bytes_per_step": 4 * (L * N + L * T * 8),
                       "l2": "single subject: inputs are L2 resident between steps (mode used to show the sharded path, not the headline)"},
            "gpu_launches": int(ctx.launch_count() - launches0), "checksum": checksum})
    m.close()
    if emit:
        if rank == 0:
            emit_line(json.dumps(line))
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
    return line


# ---------------------------------------------------------------------------------------------- groupwise mode
def bench_group(args, rank, world, local, metric, unit, emit=True):
    """BASELINE.json configs[4]: groupwise registration (DiscreteGroupModel) of S subjects with group triplet costs
    (src/DiscreteGroupCostFunction.cpp:9-30), subjects sharded over the GPUs.  The S control meshes are replicated (S*N
    points, small); each GPU evaluates the fusion-move triplet costs of every label for ITS subjects' triangles and stores
    them straight into every rank's [L][S*T][8] table over peer mappings (fused exchange), one barrier per step — the joint
    host optimiser needs the whole table.  (Parity of this configuration — a sample of the table against the oracle port and
    the reference's own DiscreteGroupCostFunction — is tests/test_gpu_parity.py::test_group_triplets_full_size.)"""
    import torch
    import torch.distributed as dist
    from msm_newmeshreg_b200 import Context, host, sharding
    S, lvl = args.group_subjects, args.group_cp_level
    cxyz0, ctris = host.icosphere(lvl)
    mvd = float(np.linalg.norm(cxyz0[ctris[:, 0]] - cxyz0[ctris[:, 1]], axis=1).mean())
    # label set and cliques from the host model glue (src/DiscreteModel.cpp:91-162,267-282) on a throw-away model
    m = host.Model(sgres=lvl + 2, cpres=lvl, multivariate=False, dopt="HOCR", threads=1, **PARAMS)
    feat = random_field(cxyz0, 1)[None]
    m.set_data(cxyz0, ctris, cxyz0, ctris, feat, feat); m.initialize(); m.setup()
    bary, trip1 = m.labels(), m.triplets()
    centre = bary[0]
    m.close()
    N, T, L = len(cxyz0), len(trip1), len(bary)
    cps = np.stack([smooth_warp(cxyz0, 500 + s, 0.25 * mvd) for s in range(S)])
    trip = np.concatenate([trip1 + s * N for s in range(S)]).astype(np.int32)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    lam, mu, kappa = f32(0.2), f32(0.4), f32(1.6)
    ctx = Context(local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_params(lam=lam, rexp=2.0, mu=mu, kappa=kappa, rmode=3, group=True)
    ctx.set_cpgrid(cps.reshape(-1, 3), np.zeros((0, 3), np.int32), np.ones(S * N), np.tile(cxyz0, (S, 1)))
    ctx.set_group(S, N, T)
    ctx.set_labels(bary, centre)
    ctx.set_triplets(trip)
    lab = np.random.default_rng(7).integers(0, L, S * N).astype(np.int32)
    lab_dev = torch.from_numpy(lab).cuda()
    s0, s1 = sharding.shard_range(S, rank, world)        # whole subjects per rank
    t0, t1 = s0 * T, s1 * T
    hdl, nccl_gather, why = None, False, ""
    if world > 1 and args.cp_exchange != "nccl":
        try:
            import torch.distributed._symmetric_memory as symm_mem
            full = symm_mem.empty(L * S * T * 8, dtype=torch.float32, device=f"cuda:{local}")
            hdl = symm_mem.rendezvous(full, dist.group.WORLD)
            peers = [int(p) for p in hdl.buffer_ptrs]
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if args.cp_exchange == "mc" else 0
        except Exception as e:   # no symmetric memory on this box: kernel + NCCL all-gather instead
            nccl_gather, why = True, f" (p2p unavailable: {type(e).__name__})"
    elif world > 1:
        nccl_gather = True
    if world == 1 or nccl_gather:
        full = torch.empty(L * S * T * 8, dtype=torch.float32, device="cuda")
        peers, mc = [full.data_ptr()], 0
    d_slice = torch.empty((L, (t1 - t0) * 8), dtype=torch.float32, device="cuda") if nccl_gather else None

    def step():
        if nccl_gather:
            ctx.triplet_moves_range_dev(lab_dev.data_ptr(), -1, t0, t1, d_slice.data_ptr())
            full.view(L, S * T * 8).copy_(sharding.allgather_label_major(d_slice, S * T, unit=8).reshape(L, -1))
            return
        ctx.triplet_moves_range_peers(lab_dev.data_ptr(), -1, t0, t1, peers, mc)
        if hdl is not None:
            hdl.barrier(channel=0)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    checksum = float(full.double().sum().item())
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    evals = L * S * T * 8
    line = None
    if rank == 0:
        line = ({
            "metric": metric, "value": evals / (ms_per_step * 1e-3), "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"groupwise (BASELINE.json configs[4]): {S} subjects, ico{lvl} control grids ({N} points, {T} triangles each), "
                                   f"L={L}; step = group fusion-move triplet costs of every label, subjects sharded over {world} GPU(s), table "
                                   f"assembled on every rank by " + ("an NCCL all-gather" + why if nccl_gather else ("NVSwitch multicast stores" if mc else "peer stores") + " (fused exchange)"),
                       "S": S, "N": N, "T": T, "L": L, "table_bytes": 4 * evals, "checksum": checksum},
            "gpu_launches": int(ctx.launch_count() - launches0)})
    ctx.close()
    if args.group_data_term:
        try:
            dt = group_data_term(args, rank, world, local)
        except Exception as e:
            dt = {"error": f"{type(e).__name__}: {e}"}
        if line is not None:
            line["data_term"] = dt
    if emit:
        if rank == 0:
            emit_line(json.dumps(line))
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
    return line


def group_data_term(args, rank, world, local):
    """The groupwise DATA term (src/DiscreteGroupCostFunction.cpp:32-54 and its producers, src/DiscreteGroupModel.cpp:72-117)
    through the product's DiscreteGroupModel, subjects sharded over the GPUs: every rank produces the patch-wise rotated and
    resampled data of ITS subjects (S L adaptive-barycentric resamplings: GPU range queries + GPU point location), ONE
    all-gather of the [S][L][V_template] values (NCCL), then every rank assembles the patch maps and evaluates the 4
    fusion-move costs of every inter-subject pair for its share of the proposal labels."""
    import torch
    import torch.distributed as dist
    from msm_newmeshreg_b200 import host, sharding
    S, cpl, dl = args.group_subjects, args.group_cp_level, args.group_data_level
    txyz, ttris = host.icosphere(dl)
    cxyz0, ctris = host.icosphere(cpl)
    mvd = float(np.linalg.norm(cxyz0[ctris[:, 0]] - cxyz0[ctris[:, 1]], axis=1).mean())
    shared = random_field(txyz, 299)
    data = np.stack([random_field(txyz, 300 + s) + 0.5 * shared for s in range(S)])
    m = host.GroupModel(lam=f32(0.2), sgres=cpl + 2, cpres=cpl, threads=max(1, (os.cpu_count() or 1) // world))
    m.set_data(txyz, ttris, data)
    t0 = time.perf_counter()
    m.initialize()
    for s in range(S):
        m.reset_meshspace(s, smooth_warp(txyz, 400 + s, 0.15 * mvd))
        m.reset_cpgrid(s, smooth_warp(cxyz0, 450 + s, 0.05 * mvd))
    s0, s1 = sharding.shard_range(S, rank, world)
    m.set_subject_shard(s0, s1)
    t1 = time.perf_counter()
    m.setup()                                            # cliques + this rank's S_local * L resamplings
    t2 = time.perf_counter()
    if world > 1:
        mine = torch.from_numpy(np.stack([m.rotated_all(s) for s in range(s0, s1)])).cuda()
        allv = sharding.allgather_subject_rows(mine, S).cpu().numpy()     # ONE all-gather of the [S][L][V_t] values (NCCL)
        for s in range(S):
            if not (s0 <= s < s1):
                m.set_rotated(s, allv[s])
    t3 = time.perf_counter()
    if world > 1:
        m.finish_setup()                                 # patch maps of every subject -> CSR -> device
    lab = np.random.default_rng(7).integers(0, m.L, m.N).astype(np.int32)
    m.pair_moves(lab, 0)                                 # first call uploads the maps
    t4 = time.perf_counter()
    mine_labels = [a_ for a_ in range(m.L) if a_ % world == rank]
    chk = 0.0
    for a_ in mine_labels:
        chk += float(m.pair_moves(lab, a_).sum())
    t5 = time.perf_counter()
    tt = torch.tensor([t2 - t1, t3 - t2, t4 - t3, t5 - t4, chk], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        tt = torch.cat([mx[:4], sm[4:]])
    tt = tt.cpu().numpy()
    out = {"S": S, "data_mesh": f"ico{dl}", "control_grid": f"ico{cpl}", "nodes": m.N, "pairs": m.P, "L": m.L,
           "producers_s": float(tt[0]), "exchange_s": float(tt[1]), "patch_maps_and_upload_s": float(tt[2]),
           "pair_moves_all_labels_s": float(tt[3]), "pair_move_evals": int(m.P) * 4 * int(m.L),
           "pair_move_evals_per_s": int(m.P) * 4 * int(m.L) / float(tt[3]), "checksum": float(tt[4]),
           "exchanged_bytes": int(S) * int(m.L) * int(m.Vt) * 8 if world > 1 else 0,
           "note": "wall clock, max over ranks; producers = S L adaptive-barycentric resamplings sharded by subject"}
    m.close()
    return out


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--subjects-per-gpu", type=int, default=8, help="independent registrations per GPU (BASELINE C4: 8 per GPU)")
    ap.add_argument("--data-level", type=int, default=7)
    ap.add_argument("--cp-level", type=int, default=5)
    ap.add_argument("--channels", type=int, default=1)
    ap.add_argument("--cpu-rows", type=int, default=1024, help="control points in the bounded CPU sample")
    ap.add_argument("--e2e-threads", type=int, default=0,
                    help="host threads driving the subjects of one GPU in the e2e leg (0 = one per subject, at most host cores / ranks)")
    ap.add_argument("--mode", default="subjects", choices=["subjects", "cp", "group"],
                    help="subjects: B independent subjects per GPU (weak scaling, default); cp: ONE subject, control points "
                         "and triplets sharded over the GPUs (strong scaling); group: groupwise registration "
                         "(DiscreteGroupCostFunction triplet costs of --group-subjects subjects, subjects sharded over the GPUs)")
    ap.add_argument("--resident-streams", type=int, default=8,
                    help="streams the subjects of one GPU are spread over in the resident (device-timed) step: independent registrations "
                         "overlap on the device, fork/join by events around the timed region (1 = everything on one stream)")
    ap.add_argument("--level-wall", type=int, default=1,
                    help="1: also time one resolution level end to end (one outer iteration; N = 1 only; needs oracle/_ref for the "
                         "reference's optimiser); 0: skip")
    ap.add_argument("--secondary", type=int, default=1,
                    help="1: after the headline measurement also run the CP-sharded (configs[2] as written) and groupwise (configs[4]) "
                         "steps on the same GPUs and attach them as `secondary`; 0: skip")
    ap.add_argument("--group-data-term", type=int, default=0,
                    help="--mode group: also run the groupwise DATA term (producers sharded by subject + all-gather + pair moves); "
                         "tens of seconds at the default size")
    ap.add_argument("--group-data-level", type=int, default=6)
    ap.add_argument("--group-subjects", type=int, default=16)
    ap.add_argument("--group-cp-level", type=int, default=4)
    ap.add_argument("--cp-exchange", default="mc", choices=["mc", "p2p", "nccl"],
                    help="--mode cp / group: mc = kernels store each entry ONCE to the NVSwitch multicast address of the table "
                         "(falls back to p2p); p2p = one peer store per rank through symmetric-memory mappings; "
                         "nccl = all_gather_into_tensor + re-layout (cp mode only)")
    args = ap.parse_args()
    capture_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    C = args.channels
    workload = (f"C3 (BASELINE.json configs[2]): ico{args.data_level} data x ico{args.cp_level} control grid, L=19 labels, C={C}, "
                f"range 1, Pearson; step = patch construction + unary table + all-label fusion-move triplet costs for {args.subjects_per_gpu} independent subjects per GPU")
    metric, unit = "unary+triplet cost evals/s", "evals/s"

    # ------------------------------------------------------------------ reference (CPU) arm
    if args.impl == "reference":
        if rank != 0:
            return
        import __graft_entry__ as g
        try:
            g.build_oracle()
        except Exception as e:  # prebuilt oracle/liborc.so travels with the snapshot
            print(f"# oracle rebuild skipped: {e}", file=sys.stderr)
        threads = os.cpu_count() or 1
        r, kind, sample = cpu_sample(args.data_level, args.cp_level, C, args.cpu_rows, threads, steps=max(args.steps, 1), warmup=min(args.warmup, 1))
        v = r["evals_per_s_e2e"]
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * (r["t_patches"] + r["t_unary"] + r["t_triplet"]), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": workload, "sample": sample},
                "cpu_baseline": {"value": v, "unit": unit, "cores": threads, "kind": kind, "sample": sample,
                                 "value_tables_only": r["evals_per_s_resident"], "t_patches_s": r["t_patches"], "t_unary_s": r["t_unary"],
                                 "t_triplet_s": r["t_triplet"]},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
                "value_tables_only": r["evals_per_s_resident"]}
        if args.level_wall and kind == "reference":
            try:
                line["level_wall"] = level_wall(args.data_level, args.cp_level, C, "reference", 1)
            except Exception as e:
                line["level_wall"] = {"error": f"{type(e).__name__}: {e}"}
        if args.secondary:
            try:
                line["feature_preprocessing"] = feature_preprocessing("reference")
            except Exception as e:
                line["feature_preprocessing"] = {"error": f"{type(e).__name__}: {e}"}
        emit_line(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    numa = bind_to_gpu_numa_node(local, world) if world > 1 else {"bound": False, "why": "single GPU"}
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    os.environ["MSM_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    if local == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from msm_newmeshreg_b200 import host

    if args.mode == "cp":
        return bench_cp_sharded(args, rank, world, local, C, metric, unit)
    if args.mode == "group":
        return bench_group(args, rank, world, local, metric, unit)
    B = args.subjects_per_gpu
    txyz, ttris = host.icosphere(args.data_level)
    cxyz0, ctris = host.icosphere(args.cp_level)
    edge = np.linalg.norm(cxyz0[ctris[:, 0]] - cxyz0[ctris[:, 1]], axis=1)
    mvd = float(edge.mean())
    # Explicit (non-default) streams.  `stream` is the timing stream: the events below are recorded on it.  The subjects of this
    # GPU are independent registrations and are spread round-robin over `work_streams`, so that the small latency-bound kernels
    # of one subject (voxel grid, per-control-point set-up) overlap the big ones of another; every timed region forks from
    # `stream` (each work stream waits for the start event) and joins back into it (it waits for every work stream's end event)
    # before the stop event is recorded — all on the device.
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    n_streams = max(1, min(args.resident_streams, args.subjects_per_gpu))
    work_streams = [stream] if n_streams == 1 else [torch.cuda.Stream() for _ in range(n_streams)]
    models, ctxs, subs, host_bufs = [], [], [], []
    for s in range(B):
        seed = 1000 + 97 * rank + s
        S = Subject(txyz, ttris, C, seed, mvd)
        m = host.Model(sgres=args.cp_level + 2, cpres=args.cp_level, multivariate=C > 1, dopt="HOCR", threads=1, **PARAMS)
        m.set_data(txyz, ttris, txyz, ttris, S.inp, S.ref)   # level start: source == undeformed template (_ORIG)
        m.initialize()
        m.reset_meshspace(S.sxyz)                               # deformed source of the current outer iteration
        cxyz = smooth_warp(cxyz0, seed + 200, 0.1 * mvd)
        m.reset_cpgrid(cxyz)
        m.setup()
        c = m.context()
        c.set_stream(work_streams[s % n_streams].cuda_stream)
        models.append(m); ctxs.append(c); subs.append(S)
        host_bufs.append((torch.from_numpy(S.sxyz.copy()).pin_memory(), torch.from_numpy(cxyz.copy()).pin_memory()))
    N, L, T = models[0].N, models[0].L, models[0].T
    lab_host = np.random.default_rng(7).integers(0, L, N).astype(np.int32)
    lab_dev = torch.from_numpy(lab_host).cuda()
    d_unary = [torch.empty((L, N), dtype=torch.float32, device="cuda") for _ in range(B)]
    d_moves = [torch.empty((L, T, 8), dtype=torch.float32, device="cuda") for _ in range(B)]
    evals_per_subject = N * L + L * T * 8
    patch_totals = []
    alg_bytes_unary = []
    for c in ctxs:
        off, idx = c.patches()
        P = np.diff(off).astype(np.int64)
        alg_bytes_unary.append(int((P * ((48 + 12 * C) * L + (16 + 8 * C)) + 8 * L).sum()))
        patch_totals.append(int(P.sum()))

    redone_steps = [0]

    def subject_launches(s, with_patches):
        if with_patches:
            ctxs[s].build_patches_async()                # get_source_data: patches of the (device-resident) source / control grid
        ctxs[s].unary_table_dev(d_unary[s].data_ptr())
        ctxs[s].triplet_moves_dev(lab_dev.data_ptr(), -1, d_moves[s].data_ptr())

    def resident_step(with_patches=True):
        for s in range(B):
            subject_launches(s, with_patches)
        if with_patches:
            # the builds' status words are read once every subject's work is enqueued (no host wait inside the step's launches);
            # a flagged build (boundary ties / row overflow) has been repeated through the exact path: redo that subject
            for s in range(B):
                if ctxs[s].patches_resolve():
                    redone_steps[0] += 1
                    subject_launches(s, False)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        nonlocal n_streams
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record(stream)
        if n_streams > 1:
            for ws in work_streams:
                ws.wait_event(e0)                 # fork
        for _ in range(steps):
            fn()
        if n_streams > 1:
            for ws in work_streams:
                j = torch.cuda.Event()
                j.record(ws)
                stream.wait_event(j)              # join
        e1.record(stream)
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    for _ in range(args.warmup):
        resident_step()
    sync_all()
    ms_tables_only = timed(lambda: resident_step(False), args.steps)      # secondary: tables only (no patch construction)
    launches0 = sum(c.launch_count() for c in ctxs)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ms_per_step = timed(resident_step, args.steps)                          # headline: patches + unary table + fusion moves
    launches = sum(c.launch_count() for c in ctxs) - launches0
    # Per-kernel durations (the roofline's denominator) are only meaningful when kernels of different subjects do not share the
    # device: a second timed region of the SAME step with every subject on the timing stream, event pairs around each family
    ms_single_stream = ms_per_step
    if n_streams > 1:
        for c in ctxs:
            c.set_stream(stream.cuda_stream)
        n_streams_saved, n_streams = n_streams, 1
        resident_step()
    for c in ctxs:
        c.set_timing(2)
    ms_timed_kernels = timed(resident_step, args.steps)
    if args.resident_streams > 1 and B > 1:
        ms_single_stream = ms_timed_kernels
    fam = {}
    for c in ctxs:
        for f in ("unary_main", "triplet_moves", "patches", "unary"):
            a_, b_ = c.timing_collect(f)
            fam.setdefault(f, [0.0, 0]); fam[f][0] += a_; fam[f][1] += b_
        c.set_timing(0)
    um, un = fam["unary_main"]; tm_, tn_ = fam["triplet_moves"]
    clocks = sampler.stop() if sampler else None
    value = evals_per_subject * B * world / (ms_per_step * 1e-3)
    value_tables_only = evals_per_subject * B * world / (ms_tables_only * 1e-3)
    one_pass = [c.patch_build_stats() for c in ctxs]

    # ---- the triplet kernels against their byte model ((84/k) + 4 B per evaluation; k = 8 per fusion move,
    # k = L^3 for the full look-up table that serves the unmodified Fusion::optimize)
    trip_roof = None
    if rank == 0:
        try:
            c0 = ctxs[0]
            c0.set_timing(1)
            full = torch.empty((T, L, L, L), dtype=torch.float32, device="cuda")
            tt = []
            for _ in range(5):
                c0.triplet_table_dev(0, T, full.data_ptr()); tt.append(c0.last_kernel_ms("triplet_table"))
            mm = []
            for _ in range(5):
                c0.triplet_moves_dev(lab_dev.data_ptr(), -1, d_moves[0].data_ptr()); mm.append(c0.last_kernel_ms("triplet_moves"))
            c0.set_timing(0)
            del full
            peak_, _src = measured_peak()
            tb, mb = (84.0 / L ** 3 + 4.0) * T * L ** 3, (84.0 / 8 + 4.0) * T * L * 8
            trip_roof = {"bound": "hbm", "unit": "GB/s", "peak": peak_,
                         "table": {"kernel": "k_triplet_prep + k_triplet_table_fast", "evals": T * L ** 3, "algorithmic_bytes": tb, "ms": min(tt),
                                   "achieved": tb / (min(tt) * 1e-3) / 1e9, "frac": tb / (min(tt) * 1e-3) / 1e9 / peak_},
                         "moves": {"kernel": "k_triplet_moves_fast", "evals": T * L * 8, "algorithmic_bytes": mb, "ms": min(mm),
                                   "achieved": mb / (min(mm) * 1e-3) / 1e9, "frac": mb / (min(mm) * 1e-3) / 1e9 / peak_},
                         "note": "best of 5 launches each, one subject, CUDA events on the launching stream (msm_set_timing 1)"}
        except Exception as e:
            trip_roof = {"error": f"{type(e).__name__}: {e}"}

    # ---- end to end through the host classes, host buffers in/out
    # The B subjects of a GPU are independent registrations: each one is driven by its own host thread on its context's
    # private stream, so one subject's uploads, kernels, read-backs and host-side mesh copies overlap the others'
    # (PCIe is full duplex; ctypes releases the GIL inside the host library).  --e2e-threads 1 is the serial loop.
    def subject_step(s):
        sx, cx = host_bufs[s]
        models[s].reset_meshspace(sx.numpy())
        models[s].reset_cpgrid(cx.numpy())
        models[s].setup()                       # patch construction + unary table -> host double[L][N]
        models[s].triplet_moves(lab_host, -1, out=moves_host[s].numpy())   # -> pinned host float[L][T][8] (fp32 table)

    for c in ctxs:
        c.set_stream(0)                         # back to the context's own stream
    moves_host = [torch.empty((L, T, 8), dtype=torch.float32).pin_memory() for _ in range(B)]
    k_e2e = max(1, min(args.steps, 10))

    def time_e2e(nthreads):
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(nthreads) if nthreads > 1 else None

        def subject_steps(s, k):
            for _ in range(k):
                subject_step(s)

        def e2e_steps(k):     # k steps of every subject; no barrier between the steps of different subjects
            if pool:
                list(pool.map(lambda s: subject_steps(s, k), range(B)))
            else:
                for _ in range(k):
                    for s in range(B):
                        subject_step(s)
        e2e_steps(1)
        sync_all()
        reps = []
        for _ in range(3):   # host-side timing is noisy: median of three repetitions of k_e2e steps
            t0 = time.perf_counter()
            e2e_steps(k_e2e)
            torch.cuda.synchronize()
            reps.append((time.perf_counter() - t0) / k_e2e)
        if pool:
            pool.shutdown()
        t = torch.tensor([sorted(reps)[1]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return evals_per_subject * B * world / float(t.item())

    cores_here = len(os.sched_getaffinity(0)) if numa.get("bound") else (os.cpu_count() or 1) // world
    ranks_sharing = numa.get("ranks_on_node", 1) if numa.get("bound") else 1
    e2e_threads = args.e2e_threads if args.e2e_threads > 0 else max(1, cores_here // ranks_sharing)
    e2e_threads = max(1, min(B, e2e_threads))
    e2e_serial = time_e2e(1)
    e2e_value = time_e2e(e2e_threads) if e2e_threads > 1 else e2e_serial
    Vs = len(txyz)
    h2d = B * (Vs * 24 + N * 24 + N * 4)
    d2h = B * (L * N * 4 + L * T * 8 * 4)

    if rank == 0:
        peak, peak_src = measured_peak()
        unary_ms = um / max(un, 1)
        achieved = (np.mean(alg_bytes_unary) / (unary_ms * 1e-3)) / 1e9 if un else None
        traffic = l2_traffic = l1_traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("k_unary_dram_bytes_per_launch")
                l2_traffic, l1_traffic = tj.get("k_unary_l2_bytes_per_launch"), tj.get("k_unary_l1_global_load_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (unary) / f64 (triplet)",
            "data": "synthetic",
            "config": {"workload": workload, "subjects_per_gpu": B, "N": N, "L": L, "T": T, "mean_patch": float(np.mean(patch_totals)) / N,
                       "evals_per_subject": evals_per_subject, "l2": "working set of B subjects per GPU exceeds the 126 MB L2 (no flush needed)",
                       "parallelism": f"subjects x{world} (no data-path collective)",
                       "resident_streams": max(1, min(args.resident_streams, B))},
            "roofline": {"bound": "hbm", "kernel": "k_unary", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_l2": l2_traffic,
                         "traffic_l1_global_loads": l1_traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": float(np.mean(alg_bytes_unary)), "avg_launch_ms": unary_ms, "launches_timed": un,
                         "triplet_moves_avg_ms": tm_ / max(tn_, 1)},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": k_e2e,
                    "host_threads": e2e_threads, "value_one_host_thread": e2e_serial, "numa": numa,
                    "api": "NonLinearSRegDiscreteModel::reset_meshspace/reset_CPgrid/setupCostFunction + computeTripletMoves"},
            "gpu_launches": int(launches), "clocks": clocks,
            "value_tables_only": value_tables_only, "ms_per_step_tables_only": ms_tables_only,
            "ms_per_step_one_stream": ms_single_stream,
            "step": "per subject: get_source_data (patch construction, one-pass path) + whole unary table + all-label fusion-move triplet costs",
            "kernel_ms_per_subject": {f: (v[0] / v[1] if v[1] else None) for f, v in fam.items()},
            "patch_builds_one_pass_vs_exact": [sum(x[0] for x in one_pass), sum(x[1] for x in one_pass)],
            "patch_builds_repeated_in_timed_steps": redone_steps[0],
            "roofline_triplet": trip_roof,
        }
        if world == 1:
            try:
                threads = os.cpu_count() or 1
                r, kind, sample = cpu_sample(args.data_level, args.cp_level, C, args.cpu_rows, threads, steps=1, warmup=0)
                line["cpu_baseline"] = {"value": r["evals_per_s_e2e"], "unit": unit, "cores": threads, "kind": kind, "sample": sample,
                                        "value_tables_only": r["evals_per_s_resident"]}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": unit, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    else:
        line = None
    # ---- secondary: the other sharded configurations of BASELINE.json on the same GPUs (all ranks take part)
    secondary = {}
    if args.secondary:
        for m_ in models:
            m_.close()
        models.clear(); ctxs.clear()
        torch.cuda.empty_cache()
        for name, fn in (("cp_sharded", lambda: bench_cp_sharded(args, rank, world, local, C, metric, unit, emit=False)),
                         ("groupwise", lambda: bench_group(args, rank, world, local, metric, unit, emit=False))):
            try:
                secondary[name] = fn()
            except Exception as e:
                secondary[name] = {"error": f"{type(e).__name__}: {e}"}
            if world > 1:
                dist.barrier()
    if rank == 0:
        if args.secondary:
            try:
                secondary["feature_preprocessing"] = feature_preprocessing("gpu")
            except Exception as e:
                secondary["feature_preprocessing"] = {"error": f"{type(e).__name__}: {e}"}
        if secondary:
            line["secondary"] = secondary
        if args.level_wall and world == 1:
            try:
                line["level_wall"] = level_wall(args.data_level, args.cp_level, C, "gpu", os.cpu_count() or 1)
            except Exception as e:
                line["level_wall"] = {"error": f"{type(e).__name__}: {e}"}
        emit_line(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
