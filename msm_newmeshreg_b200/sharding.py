"""Multi-GPU plumbing for the cost tables: one process per GPU, control points (or triplets, or
subjects) split into contiguous ranges, no collective on the data path; the finished table slices are gathered once
per outer iteration with a single all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests).

torch is used for process-group plumbing and device memory only.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous range [begin, end) of `n` units owned by `rank` of `world` (balanced to within one unit)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def subjects_of_rank(n_subjects, rank, world):
    """Round-robin subject assignment for batched / groupwise registration (independent registrations: no exchange)."""
    return list(range(rank, n_subjects, world))


def allgather_label_major(local, n, group=None, unit=1):
    """All-gather per-rank table slices local[L, n_r*unit] (label-major rows, this rank's contiguous units as columns,
    `unit` values per unit) into the reference layout [L, n*unit] (unary[label*N + node], src/DiscreteCostFunction.cpp:312;
    fusion-move tables [L][T][8] use unit=8).

    Works on CUDA tensors (NCCL) and CPU tensors (gloo).  Slices are padded to the largest shard so that one
    fixed-size all_gather suffices."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    L = local.shape[0]
    sizes = [s * unit for s in shard_sizes(n, world)]
    mx = max(sizes)
    send = torch.zeros((L, mx), dtype=local.dtype, device=local.device)
    send[:, : local.shape[1]] = local
    recv = torch.empty((world, L, mx), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    return torch.cat([recv[r, :, : sizes[r]] for r in range(world)], dim=1)


def assemble_label_major(slices, n):
    """Host-side equivalent of allgather_label_major for numpy slices already collected in rank order."""
    world = len(slices)
    sizes = shard_sizes(n, world)
    assert all(s.shape[1] == sizes[r] for r, s in enumerate(slices))
    return np.concatenate(slices, axis=1)


def allgather_subject_rows(mine, n_subjects, group=None):
    """Groupwise registration, subjects sharded contiguously (shard_range): every rank contributes the rows of ITS subjects,
    mine[s_local, ...] (e.g. the [L][V_template] resampled values of src/DiscreteGroupModel.cpp:92), and receives all of them in
    subject order: returns a tensor [n_subjects, ...].  ONE fixed-size all_gather (shards are padded to the largest one), on CUDA
    tensors (NCCL) or CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        assert mine.shape[0] == n_subjects
        return mine
    rank = dist.get_rank(group)
    s0, s1 = shard_range(n_subjects, rank, world)
    assert mine.shape[0] == s1 - s0, "rows of this rank's subjects expected"
    per = max(shard_sizes(n_subjects, world))
    pad = torch.zeros((per,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    pad[: s1 - s0] = mine
    allv = torch.empty((world * per,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(allv, pad, group=group)
    rows = []
    for r in range(world):
        a, b = shard_range(n_subjects, r, world)
        rows.append(allv[r * per: r * per + (b - a)])
    return torch.cat(rows, dim=0)
