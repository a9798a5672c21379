"""Ad-hoc GPU exploration (not a pytest file): prints parity errors and kernel timings for a few configs.
Usage on the GPU box: python tests/gpu_explore.py [quick|big]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import problems as pb  # noqa: E402


def report(name, g, o, floor):
    e = pb.rel_err(g, o, floor)
    print(f"  {name}: max rel {e.max():.3e}  mean rel {e.mean():.3e}  max abs {np.abs(np.asarray(g) - np.asarray(o)).max():.3e}"
          f"  |ref| range [{np.abs(o).min():.3e},{np.abs(o).max():.3e}]", flush=True)


def run(data_level, cp_level, C, warp, deform_cp=0.0, rough=False, check_trip_table=True, threads=8, oracle_rows=None, params=None):
    print(f"== data ico{data_level} cp ico{cp_level} C={C} warp={warp} deform_cp={deform_cp} rough={rough}", flush=True)
    t0 = time.time()
    P = pb.make_problem(data_level, cp_level, C, seed=1, warp=warp, deform_cp=deform_cp, rough=rough)
    print(f"  problem built in {time.time() - t0:.1f}s N={P.N} L={P.L} T={P.T}", flush=True)
    ctx = pb.gpu_context(P, params=params)
    ctx.set_timing(True)
    t0 = time.time(); ctx.build_patches(); t1 = time.time()
    off, idx = ctx.patches()
    print(f"  patches: total {len(idx)} mean {len(idx) / P.N:.1f} max {np.diff(off).max()}  wall {1e3 * (t1 - t0):.2f} ms kernel-family {ctx.last_kernel_ms('patches'):.3f} ms", flush=True)
    U = ctx.unary_table()
    for _ in range(3):
        U = ctx.unary_table()
    ms = ctx.last_kernel_ms('unary')
    nsamp = len(idx) * P.L
    algb = sum((off[k + 1] - off[k]) * ((48 + 12 * C) * P.L + (16 + 8 * C)) + 8 * P.L for k in range(P.N))
    print(f"  unary kernel {ms:.4f} ms  ({P.N * P.L / ms / 1e6:.2f} G evals/s... {P.N * P.L / (ms * 1e-3):.3e} evals/s, {nsamp / (ms * 1e-3):.3e} samples/s, alg {algb / 1e9:.3f} GB -> {algb / (ms * 1e-3) / 1e9:.0f} GB/s)", flush=True)
    cf = pb.oracle_costfunction(P, params=params, threads=threads)
    rows = P.N if oracle_rows is None else min(oracle_rows, P.N)
    t0 = time.time(); cf.get_source_data(0, rows); t1 = time.time()
    po = cf.patches()
    same = all(np.array_equal(po[k], idx[off[k]:off[k + 1]]) for k in range(rows))
    print(f"  oracle patches rows={rows} {t1 - t0:.2f}s identical membership: {same}", flush=True)
    t0 = time.time(); Uo = cf.unary_costs(0, rows); t1 = time.time()
    print(f"  oracle unary {t1 - t0:.2f}s ({rows * P.L / (t1 - t0):.3e} evals/s, {threads} threads)", flush=True)
    report("unary", U[:, :rows], Uo, 1e-2)
    rng = np.random.default_rng(0)
    nq = 20000
    q = np.stack([rng.integers(0, P.T, nq), rng.integers(0, P.L, nq), rng.integers(0, P.L, nq), rng.integers(0, P.L, nq)], 1).astype(np.int32)
    q[:100, 1:] = 0
    tg = ctx.triplet_costs(q); to = cf.triplet_costs(q)
    report("triplet queries", tg, to, 1e-3)
    lab = rng.integers(0, P.L, P.N).astype(np.int32)
    mv = ctx.triplet_moves(lab, 3)
    print(f"  triplet moves kernel {ctx.last_kernel_ms('triplet_moves'):.4f} ms", flush=True)
    if check_trip_table:
        nt = min(P.T, 64)
        tt = ctx.triplet_table(0, nt); tto = cf.triplet_table(0, nt)
        report("triplet table", tt, tto, 1e-3)
    tall = min(P.T, 20480)
    import ctypes
    ctx.triplet_table(0, min(P.T, 512))
    print(f"  triplet table kernel ({min(P.T, 512)} triplets) {ctx.last_kernel_ms('triplet_table'):.4f} ms", flush=True)
    ctx.close()
    # pairwise model
    ctx = pb.gpu_context(P, params=dict(rmode=1), cliques="pairs")
    ctx.set_timing(True)
    pt = ctx.pair_table()
    print(f"  pair table kernel {ctx.last_kernel_ms('pair_table'):.4f} ms  P={len(P.pairs)}", flush=True)
    cfp = pb.oracle_costfunction(P, params=dict(rmode=1), cliques="pairs", threads=threads)
    if len(P.pairs) * P.L * P.L <= 3_000_000:
        pto = cfp.pair_table()
        report("pair table", pt, pto, 1e-3)
    ctx.close()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    run(4, 2, 1, 0.3)
    run(4, 2, 1, 0.0)
    run(5, 3, 4, 0.3, deform_cp=0.2)
    run(5, 3, 1, 0.3, deform_cp=0.2, rough=True)
    if mode == "big":
        run(6, 2, 1, 0.3)
        run(6, 4, 4, 0.3, deform_cp=0.2)
        run(7, 5, 1, 0.3, deform_cp=0.1, check_trip_table=True, oracle_rows=600)
        run(7, 5, 4, 0.3, deform_cp=0.1, check_trip_table=False, oracle_rows=600)
