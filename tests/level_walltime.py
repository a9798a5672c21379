"""Per-level wall time of a complete multi-resolution registration (BASELINE.json configs[1]: pairwise ico6, 4-channel
features, 3 levels with control grids ico2 -> ico3 -> ico4, triplet strain regulariser, HOCR) — test infrastructure:
it drives the GPU-backed host classes with the REFERENCE's own optimisers from oracle/_ref (not part of the product),
and with --check-ref / --reference-only THE REFERENCE ITSELF on the CPU (its cost functions + optimisers compiled in place).

Restates Mesh_registration::run_multiresolutions / project_CPgrid / run_discrete_opt (src/mesh_registration.cpp:9-29,
282-314,316-397) for a fixed data resolution.  With --check the same registration is assembled from oracle pieces and the
labellings of every iteration of every level, and the final warped sphere, are compared.

  python tests/level_walltime.py [--data 6] [--levels 2,3,4] [--channels 4] [--iters 3] [--check] [--out profiles/x.json]
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import problems as pb  # noqa: E402
from oracle import pyorc, pyref  # noqa: E402


class GpuBackend:
    """GPU-backed mirror classes + the reference's optimisers (oracle/_ref/libref_optim.so)."""
    name = "gpu"

    def model(self, threads, **kw):
        from msm_newmeshreg_b200 import host
        return host.Model(_lib_override=pyref.lib(), threads=threads, **kw)

    def mesh_interpolate(self, *a):
        from msm_newmeshreg_b200 import host
        return host.mesh_interpolate(*a, _lib_override=pyref.lib())

    def unfold(self, *a):
        from msm_newmeshreg_b200 import host
        return host.unfold(*a, _lib_override=pyref.lib())

    def run_discrete_opt(self, m, src, iters, threads):
        return pyref.run_discrete_opt(m, src, iters, threads=threads)


class ReferenceBackend:
    """THE REFERENCE on the CPU: its own cost functions, model and optimisers (oracle/_ref/libref_cost.so).  Everything
    runs on ONE thread: the reference's threaded unary loop is racy (tests/test_reference_pin.py), and Fusion::optimize
    calls computeUnaryCost from an OpenMP loop (include/Fusion/Fusion.h:147-150) while the multivariate cost function
    clears and resizes shared per-vertex vectors there (src/DiscreteCostFunction.cpp:487-505) — with more than one thread
    that is a heap race (AddressSanitizer: heap-use-after-free at :502), so numthreads > 1 is undefined behaviour in the
    reference itself.  The GPU arm's virtuals are re-entrant table look-ups and may use every host thread."""
    name = "reference"

    def model(self, threads, **kw):
        from oracle import pyrefcost
        return pyrefcost.model(threads=1, **kw)

    def mesh_interpolate(self, *a):
        from oracle import pyrefcost
        return pyrefcost.mesh_interpolate(*a)

    def unfold(self, *a):
        from oracle import pyrefcost
        return pyrefcost.unfold(*a)

    def run_discrete_opt(self, m, src, iters, threads):
        from oracle import pyrefcost
        return pyrefcost.run_discrete_opt(m, src, iters, threads=1)


def run_levels(B, P, levels, iters, C, threads, lam):
    src = P.txyz.copy()
    report, labs_all = [], []
    for cp_level in levels:
        t0 = time.perf_counter()
        pr = dict(pb.DEFAULT_PARAMS); pr.update(rmode=3, dopt="HOCR", lam=lam)
        m = B.model(threads, sgres=cp_level + 2, cpres=cp_level, multivariate=C > 1, **pr)
        m.set_data(P.txyz, P.ttris, P.txyz, P.ttris, P.inp, P.ref)
        m.initialize()
        # project_CPgrid (:282-314): carry the registration of the previous level to the new control grid
        cxyz0, ctris = pyorc.icosphere(cp_level)
        if not np.array_equal(src, P.txyz):
            cp = B.mesh_interpolate(m.cpgrid(), P.txyz, P.ttris, src)
            cp = B.unfold(cp, ctris)
            m.reset_cpgrid(cp)
            src = B.unfold(src, P.ttris)
        t1 = time.perf_counter()
        src, labs, en, tm = B.run_discrete_opt(m, src, iters, threads)
        t2 = time.perf_counter()
        labs_all.append(labs)
        report.append({"cp_level": cp_level, "N": m.N, "L": m.L, "T": m.T, "iterations": len(labs), "setup_s": t1 - t0, "level_wall_s": t2 - t0,
                       "cost_eval_s": tm["cost"].tolist(), "optimiser_s": tm["optimiser"].tolist(), "warp_s": tm["warp"].tolist(),
                       "energies": en.tolist(), "labelling_sha1": [hashlib.sha1(l.tobytes()).hexdigest()[:12] for l in labs]})
        m.close()
    return src, labs_all, report


def run_gpu(P, levels, iters, C, threads, lam):
    return run_levels(GpuBackend(), P, levels, iters, C, threads, lam)


def run_reference(P, levels, iters, C, threads, lam):
    return run_levels(ReferenceBackend(), P, levels, iters, C, threads, lam)


def run_oracle(P, levels, iters, C, threads, lam):
    src = P.txyz.copy()
    labs_all = []
    for cp_level in levels:
        Q0 = pb.make_problem(P.data_level, cp_level, C, seed=P.seed, warp=0.0)
        cp = Q0.cxyz0.copy()
        if not np.array_equal(src, P.txyz):
            cp = pyorc.mesh_interpolate(cp, P.txyz, P.ttris, src)
            cp, _ = pyorc.unfold(cp, Q0.ctris)
            src, _ = pyorc.unfold(src, P.ttris)
        energy, labs = 0.0, []
        for it in range(1, iters + 1):
            Q = pb.make_problem(P.data_level, cp_level, C, seed=P.seed, warp=0.0, iteration=it)
            Q.ref, Q.inp = P.ref, P.inp
            Q.sxyz = src; Q.cxyz = cp; Q.rot = pyorc.rotations(Q.centre, cp)
            cf = pb.oracle_costfunction(Q, params=dict(rmode=3, lam=lam), threads=threads); cf.get_source_data()
            lab, newenergy = pyref.fusion_tables(cf.unary_costs(), Q.triplets, cf.triplet_table(), threads=threads)
            labs.append(lab)
            if it > 1 and (it - 1) % 2 == 0 and energy - newenergy < 0.001:
                break
            newcp = np.einsum("nij,nj->ni", Q.rot, Q.labels[lab])
            src = pyorc.mesh_interpolate(src, cp, Q.ctris, newcp)
            cp, _ = pyorc.unfold(newcp, Q.ctris)
            src, _ = pyorc.unfold(src, P.ttris)
            energy = newenergy
        labs_all.append(np.array(labs))
    return src, labs_all


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", type=int, default=6)
    ap.add_argument("--levels", default="2,3,4")
    ap.add_argument("--channels", type=int, default=4)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--lam", type=float, default=0.1)
    ap.add_argument("--check", action="store_true", help="also assemble the registration from oracle pieces and compare")
    ap.add_argument("--check-ref", action="store_true", help="also run THE REFERENCE (all CPU, oracle/_ref/libref_cost.so) and compare")
    ap.add_argument("--reference-only", action="store_true", help="run only the reference arm (no GPU needed)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    levels = [int(x) for x in a.levels.split(",")]
    lam = pb.f32(a.lam)
    P = pb.make_problem(a.data, levels[0], a.channels, seed=51, warp=0.0)
    P.seed = 51
    if a.reference_only:
        t0 = time.perf_counter()
        src_ref, labs_ref, rep_ref = run_reference(P, levels, a.iters, a.channels, a.threads, lam)
        result = {"config": f"pairwise ico{a.data}, C={a.channels}, control grids {levels}, strain regulariser (regoption 3), HOCR, {a.iters} iterations/level",
                  "arm": "reference: its own cost functions + its own optimisers, 1 host thread (more is a data race in the reference)",
                  "levels": rep_ref, "wall_s": time.perf_counter() - t0,
                  "max_vertex_displacement": float(np.linalg.norm(src_ref - P.txyz, axis=1).max())}
        print(json.dumps(result, indent=1))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(result, f, indent=1)
        return
    src_gpu, labs_gpu, report = run_gpu(P, levels, a.iters, a.channels, a.threads, lam)
    result = {"config": f"pairwise ico{a.data}, C={a.channels}, control grids {levels}, strain regulariser (regoption 3), HOCR, {a.iters} iterations/level",
              "optimiser": "reference Fusion.h/ELC.h/FastPD.h unmodified (oracle/_ref), host threads %d" % a.threads, "levels": report,
              "max_vertex_displacement": float(np.linalg.norm(src_gpu - P.txyz, axis=1).max())}
    if a.check:
        t0 = time.perf_counter()
        src_orc, labs_orc = run_oracle(P, levels, a.iters, a.channels, a.threads, lam)
        result["oracle_wall_s"] = time.perf_counter() - t0
        result["labellings_identical"] = [bool(g.shape == o.shape and np.array_equal(g, o)) for g, o in zip(labs_gpu, labs_orc)]
        result["labels_differing"] = [int((g != o).sum()) if g.shape == o.shape else -1 for g, o in zip(labs_gpu, labs_orc)]
        result["warp_max_abs_diff"] = float(np.abs(src_gpu - src_orc).max())
    if a.check_ref:
        t0 = time.perf_counter()
        src_ref, labs_ref, rep_ref = run_reference(P, levels, a.iters, a.channels, a.threads, lam)
        result["reference_wall_s"] = time.perf_counter() - t0
        result["reference_levels"] = rep_ref
        result["reference_labellings_identical"] = [bool(g.shape == o.shape and np.array_equal(g, o)) for g, o in zip(labs_gpu, labs_ref)]
        result["reference_labels_differing"] = [int((g != o).sum()) if g.shape == o.shape else -1 for g, o in zip(labs_gpu, labs_ref)]
        result["reference_warp_max_abs_diff"] = float(np.abs(src_gpu - src_ref).max())
        result["level_speedup_vs_reference"] = [r["level_wall_s"] / g["level_wall_s"] for g, r in zip(report, rep_ref)]
    print(json.dumps(result, indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(result, f, indent=1)


if __name__ == "__main__":
    main()
