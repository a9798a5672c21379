"""Timing of the feature pre-processing (src/featurespace.cc:18-108) on the GPU against the reference's own featurespace on the
host (oracle/_ref/libref_feat.so; falls back to the oracle port).  Usage: python tests/prof_features.py [data_level] [ico] [D]
Prints one JSON line.  Not a test."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
import feature_inputs as fi  # noqa: E402
from msm_newmeshreg_b200 import _capi, host  # noqa: E402
from oracle import pyorc, pyreffeat  # noqa: E402


def best(fn, n=3):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return min(ts), r


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    ico = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    mi, di, mr, dr = fi.make(level, level, D=D, seed=3, warp=1.0)
    kw = dict(sigma=(5.0, 5.0), intensitynorm=True, varnorm=True)
    out = {"data_level": level, "ico": ico, "D": D, "vertices": len(mi[0])}
    ctx = _capi.Context()
    xyz, _ = pyorc.icosphere(ico)
    vals = np.stack([np.random.default_rng(d).normal(size=len(xyz)) for d in range(D)])
    ctx.smooth_data(xyz, vals, xyz, 5.0)
    ctx.set_timing(1)
    t, sm = best(lambda: ctx.smooth_data(xyz, vals, xyz, 5.0))
    out["smooth_gpu_call_s"] = t
    out["smooth_gpu_kernels_ms"] = ctx.last_kernel_ms("smooth")
    t, m = best(lambda: ctx.histogram_match(vals, vals[::-1] * 2 + 1)); out["histmatch_gpu_call_s"] = t; out["histmatch_gpu_kernels_ms"] = ctx.last_kernel_ms("histmatch")
    t, v = best(lambda: ctx.variance_normalise(vals)); out["varnorm_gpu_call_s"] = t; out["varnorm_gpu_kernels_ms"] = ctx.last_kernel_ms("varnorm")
    ctx.set_timing(0)
    t0 = time.perf_counter(); want = pyorc.smooth_data(xyz, vals, xyz, 5.0); out["smooth_cpu_port_s"] = time.perf_counter() - t0
    out["smooth_max_abs_diff"] = float(np.abs(sm - want).max())
    t0 = time.perf_counter(); wm = pyorc.histogram_normalization(vals, vals[::-1] * 2 + 1); out["histmatch_cpu_port_s"] = time.perf_counter() - t0
    out["histmatch_equal"] = bool(np.array_equal(m, wm))
    t0 = time.perf_counter(); wv = pyorc.variance_normalise(vals); out["varnorm_cpu_port_s"] = time.perf_counter() - t0
    out["varnorm_max_abs_diff"] = float(np.abs(v - wv).max())
    # the whole initialize(): product mirror (GPU) vs the reference's own class (host)
    host.featurespace_initialize(ico, mi, di, mr, dr, **kw)
    t, a = best(lambda: host.featurespace_initialize(ico, mi, di, mr, dr, **kw), 2)
    out["initialize_gpu_s"] = t
    ref = pyreffeat.featurespace_initialize if pyreffeat.available() else pyorc.featurespace_initialize
    out["initialize_cpu_kind"] = "reference" if pyreffeat.available() else "port"
    t0 = time.perf_counter(); b = ref(ico, mi, di, mr, dr, **kw); out["initialize_cpu_s"] = time.perf_counter() - t0
    out["initialize_max_abs_diff"] = [float(np.abs(u - w).max()) for u, w in zip(a[:2], b[:2])]
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
