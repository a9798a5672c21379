"""Profiling driver (not a pytest file): sets up the C3 config (ico7 data / ico5 control grid) and launches the unary
table kernel a few times.  Used under ncu on the GPU box:  python tests/prof_unary.py [C] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import problems as pb  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
data_level = int(sys.argv[3]) if len(sys.argv) > 3 else 7
cp_level = int(sys.argv[4]) if len(sys.argv) > 4 else 5
P = pb.make_problem(data_level, cp_level, C, seed=1, warp=0.3, deform_cp=0.1)
ctx = pb.gpu_context(P)
ctx.set_timing(True)
ctx.build_patches()
ctx.build_patches()      # second build of a level: the one-pass path (fixed-capacity rows), what every later iteration sees
off, idx = ctx.patches()
ms = []
main = []
for _ in range(reps):
    U = ctx.unary_table()
    ms.append(ctx.last_kernel_ms("unary"))
    main.append(ctx.last_kernel_ms("unary_main"))
algb = sum((off[k + 1] - off[k]) * ((48 + 12 * C) * P.L + (16 + 8 * C)) + 8 * P.L for k in range(P.N))
best = min(ms)
print(f"C={C} N={P.N} L={P.L} samples={len(idx) * P.L} unary family ms: {['%.4f' % m for m in ms]}  best {best:.4f} ms "
      f"-> {P.N * P.L / best / 1e6:.3f} G evals/s, {algb / best / 1e6:.0f} GB/s algorithmic ({algb / best / 1e6 / 6549.1:.3f} of 6549.1)")
print("k_unary alone ms:", ["%.4f" % m for m in main], "best %.4f" % min(main))
print("checksum", float(np.abs(U).sum()), "launches", ctx.launch_count())
ctx.close()
