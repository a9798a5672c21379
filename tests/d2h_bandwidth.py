"""Plain pinned host<->device copy bandwidth of the box, one GPU alone and all GPUs at once (not a pytest file).
Settles whether the end-to-end scaling of bench.py (every subject-step reads a 12.5 MB fusion-move table back) is bound by
the box or by this repository's code: ONE cudaMemcpyAsync per GPU on its own stream, 256 MB pinned buffers, no kernels.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tests/d2h_bandwidth.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = 256
host = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
dev = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()


def bw(direction, reps=8, together=False):
    """together=True: EVERY rank calls this at the same time (one barrier inside); False: no collective inside."""
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            (host.copy_(dev, non_blocking=True) if direction == "d2h" else dev.copy_(host, non_blocking=True))
        torch.cuda.synchronize()
        if together and world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            (host.copy_(dev, non_blocking=True) if direction == "d2h" else dev.copy_(host, non_blocking=True))
        e1.record(stream)
        torch.cuda.synchronize()
    return reps * MB / 1024.0 / (e0.elapsed_time(e1) * 1e-3)     # GiB/s


def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world == 1:
        return [x]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


res = {"world": world, "buffer_MiB": MB, "host_cores": os.cpu_count()}
# one GPU at a time (the others idle)
mine = {"d2h": 0.0, "h2d": 0.0}
for r in range(world):
    for d in ("d2h", "h2d"):
        if r == rank:
            mine[d] = bw(d)                  # no collective inside: the other ranks wait at the barrier below
        if world > 1:
            dist.barrier()                   # every rank, once per (r, d)
res["alone_GiBps"] = {d: gather(mine[d]) for d in ("d2h", "h2d")}
# all GPUs at once
for d in ("d2h", "h2d"):
    per = gather(bw(d, together=True))
    res[f"concurrent_{d}_GiBps"] = {"per_gpu": per, "aggregate": sum(per)}
# both directions at once on every GPU (what the end-to-end step does: uploads of one subject overlap read-backs of another)
s2 = torch.cuda.Stream()
h2 = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(8):
    with torch.cuda.stream(stream):
        host.copy_(dev, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
per = gather(8 * MB / 1024.0 / dt)
res["concurrent_bidirectional_GiBps_each_way"] = {"per_gpu": per, "aggregate": sum(per)}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
