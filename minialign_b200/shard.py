"""Read sharding across the GPUs of one box: contiguous input batches dealt round-robin to ranks, a
full index copy per GPU, and ONE exchange step: an all-gather of each rank's output byte count per batch wave followed by an
exclusive prefix sum, so every rank knows where its slice of the merged, input-ordered output stream starts.  The exchange
is 8 bytes per rank (latency-bound); NCCL on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def batches_of(rank: int, world: int, n_batches: int):
    """Batch ids handled by `rank`: round-robin, so wave w = batches [w*world, (w+1)*world)."""
    return list(range(rank, n_batches, world))


def output_offsets(local_bytes: int, device: torch.device | str = "cpu", group=None):
    """Exclusive prefix sum of the per-rank byte counts of one wave -> (my offset, total bytes of the wave)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0, int(local_bytes)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.tensor([int(local_bytes)], dtype=torch.int64, device=device)
    allv = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.cpu()
    return int(allv[:rank].sum()), int(allv.sum())


def merged_order(n_batches: int, world: int):
    """(batch id, owning rank) in output order: the reference drains batches strictly by id (minialign.c:4638-4643)."""
    return [(b, b % world) for b in range(n_batches)]
