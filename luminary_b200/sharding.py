"""Multi-GPU work partition of the path tracer.

Every (pixel, sample id) pair is independent and fully determined by the counter-based sampler, the scene is read-only,
so the path shards by *sample id* with the scene replicated: rank r of G renders ids r, r + G, r + 2G, ... of the whole
frame into its private accumulation planes. The only exchange step is a sum-reduce of the 4 float planes plus the
executed sample counts. The reference does the same partition through a shared sample-id allocator
(device/device_adaptive_sampler.c:58-71) and combines through pinned host memory + `buffer_add`
(device/device_result_interface.c:107-299); here the combine is one NCCL reduce over NVLink.
"""
from __future__ import annotations

from typing import Tuple


def rank_sample_ids(total_samples: int, rank: int, world: int, first_sample: int = 0) -> Tuple[int, int, int]:
    """Returns (first id, count, stride) of the sample ids rank `rank` renders out of `total_samples` (strong split).
    Ids are interleaved so that any prefix of the pass sequence is balanced across ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("invalid rank / world size")
    count = (total_samples - rank + world - 1) // world if total_samples > rank else 0
    return first_sample + rank, count, world


def reduce_planes(planes, sample_count: int, dst: int = 0):
    """Sum-reduces the accumulation planes (a torch tensor on this rank's device) and the executed sample counts onto
    rank `dst`. Returns the global sample count (valid on every rank)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sample_count
    dist.reduce(planes, dst=dst, op=dist.ReduceOp.SUM)
    n = torch.tensor([sample_count], dtype=torch.int64, device=planes.device)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return int(n.item())
