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
