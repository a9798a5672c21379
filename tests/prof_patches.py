"""Profiling driver (not a pytest file): patch construction on the C3 config (ico7 data / ico5 control grid), second and
later builds of a level (the one-pass path).  Usage on the GPU box:  python tests/prof_patches.py [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import problems as pb  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
P = pb.make_problem(7, 5, 1, seed=1, warp=0.3, deform_cp=0.1)
ctx = pb.gpu_context(P)
ctx.build_patches()
ctx.set_timing(1)
ms = []
for r in range(reps):
    ctx.reset_source(P.sxyz)          # as every outer iteration does: new source coordinates
    ctx.build_patches()
    ms.append(ctx.last_kernel_ms("patches"))
print(f"patch construction (builds 2..{reps + 1}): ms {['%.4f' % m for m in ms]} best {min(ms):.4f}; paths (one-pass, exact) = {ctx.patch_build_stats()}")
off, idx = ctx.patches()
print("total samples", len(idx), "largest patch", int(np.diff(off).max()))
ctx.close()
