"""Multi-GPU sharding of the target domain: one process per GPU, samples replicated,
rank r owns the contiguous linear-index slab [⌊T·r/R⌋, ⌊T·(r+1)/R⌋), no exchange during compute,
results gathered with one collective (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations


def slab_bounds(n_targets: int, rank: int, world: int):
    """(first, count) of rank's slab."""
    first = n_targets * rank // world
    last = n_targets * (rank + 1) // world
    return first, last - first


def gather_slabs(local, n_targets: int, group=None):
    """All-gather per-rank slabs (1-D tensors, slab order = rank order) into the full length-T tensor
    on every rank. Equal slabs use one all_gather_into_tensor; ragged slabs are padded to the widest."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = [slab_bounds(n_targets, r, world)[1] for r in range(world)]
    if local.shape[0] != counts[dist.get_rank(group)]:
        raise ValueError("local slab has the wrong length for this rank")
    if len(set(counts)) == 1:
        out = torch.empty(n_targets, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    width = max(counts)
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty(world * width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * width: r * width + counts[r]] for r in range(world)])
