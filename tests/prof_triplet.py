"""Profiling driver (not a pytest file): triplet / pairwise table kernels on the C3 config (ico5 control grid).
Usage on the GPU box:  python tests/prof_triplet.py [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import problems as pb  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cp_level = int(sys.argv[2]) if len(sys.argv) > 2 else 5
P = pb.make_problem(cp_level, cp_level, 1, seed=1, warp=0.0, deform_cp=0.1)   # data mesh irrelevant here: keep it small
import torch  # noqa: E402
ctx = pb.gpu_context(P)
ctx.set_timing(1)
L, T = P.L, P.T
out = torch.empty((T, L, L, L), dtype=torch.float32, device="cuda")
ms = []
for _ in range(reps):
    ctx.triplet_table_dev(0, T, out.data_ptr())
    ms.append(ctx.last_kernel_ms("triplet_table"))
torch.cuda.synchronize()
n = T * L ** 3
best = min(ms)
print(f"triplet full table: T={T} L={L} evals={n} ms {['%.4f' % m for m in ms]} best {best:.4f} -> {n / best / 1e6:.2f} G evals/s, "
      f"{4 * n / best / 1e6:.0f} GB/s stored ({4 * n / best / 1e6 / 6549.1:.3f} of 6549.1)")
lab = torch.from_numpy(np.random.default_rng(0).integers(0, L, P.N).astype(np.int32)).cuda()
mv = torch.empty((L, T, 8), dtype=torch.float32, device="cuda")
ms = []
for _ in range(reps):
    ctx.triplet_moves_dev(lab.data_ptr(), -1, mv.data_ptr())
    ms.append(ctx.last_kernel_ms("triplet_moves"))
print(f"triplet moves (all labels): evals={L * T * 8} best {min(ms):.4f} ms -> {L * T * 8 / min(ms) / 1e6:.2f} G evals/s")
ctx.close()
ctx = pb.gpu_context(P, params=dict(rmode=1), cliques="pairs")
ctx.set_timing(1)
pt = torch.empty((len(P.pairs), L, L), dtype=torch.float32, device="cuda")
ms = []
for _ in range(reps):
    ctx.pair_table_dev(pt.data_ptr())
    ms.append(ctx.last_kernel_ms("pair_table"))
print(f"pair table: P={len(P.pairs)} evals={len(P.pairs) * L * L} best {min(ms):.4f} ms -> {len(P.pairs) * L * L / min(ms) / 1e6:.2f} G evals/s")
print("checksum", float(out.double().sum()), float(mv.double().sum()), float(pt.double().sum()))
