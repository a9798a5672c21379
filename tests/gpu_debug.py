"""scratch: sharded vs single context bits"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import problems as pb

P = pb.make_problem(5, 3, 1, seed=5, warp=0.3, deform_cp=0.1)
os.environ.pop("MSM_UNARY_LG", None)
ctx = pb.gpu_context(P); ctx.build_patches(); U = ctx.unary_table()
for lg in (19, 7, 3):
    os.environ["MSM_UNARY_LG"] = str(lg)
    V = ctx.unary_table()
    print("single ctx forced Lg", lg, "diff vs natural", (V != U).sum())
ctx.close()

def shard(k0, k1):
    c = pb.gpu_context(P); c.set_shard(k0, k1); c.build_patches()
    out = np.zeros_like(U); c.unary_table(out); c.close()
    return out[:, k0:k1]

for lg in (None, 19, 10, 7, 3, 1):
    if lg is None: os.environ.pop("MSM_UNARY_LG", None)
    else: os.environ["MSM_UNARY_LG"] = str(lg)
    a = shard(0, 214); d = a != U[:, :214]
    print("shard(0,214) Lg", lg, "diff", d.sum(), "labels", np.unique(np.nonzero(d)[0]))
