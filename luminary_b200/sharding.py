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


# ---------------------------------------------------------------------------------------------
# adaptive sampling across ranks: the shared sampler of the reference (device/device_adaptive_sampler.c:58-71: every execution
# gets its DeviceSampleAllocation from one allocator; the main device builds the stages from the combined result)
# ---------------------------------------------------------------------------------------------
ADAPTIVE_STAGES = 4


def adaptive_allocations(num_executions: int, update_interval: int):
    """The allocator: for global execution e -> (stage id, executions per stage finished before e). A stage lasts
    update_interval << stage executions (device_renderer.c:350-376); stage 4 is the last."""
    out = []
    ex = [0] * (ADAPTIVE_STAGES + 1)
    stage = 0
    for _ in range(num_executions):
        out.append((stage, list(ex)))
        ex[stage] += 1
        if stage < ADAPTIVE_STAGES and ex[stage] >= (update_interval << stage):
            stage += 1
    return out, stage, ex


def render_adaptive_sharded(renderer, num_executions: int, update_interval: int, rank: int, world: int, combine, broadcast_words):
    """Runs the adaptive schedule over `world` ranks. Executions of one stage are dealt round-robin; at every stage boundary the
    planes are combined onto rank 0 (`combine()`: sum onto rank 0, zero elsewhere), rank 0 builds the next stage and its counts are
    broadcast (`broadcast_words(words or None) -> words`), because every rank must derive sample ids from identical counts.
    `renderer` offers set_state(stage, executions, words), render(executions_before), build_stage() -> words (rank 0 only).
    Returns (stage, executions) at the end; rank 0's planes then hold everything but the last stage's un-combined part - call
    combine() once more before resolving."""
    alloc, end_stage, end_ex = adaptive_allocations(num_executions, update_interval)
    words = None
    cur_stage = 0
    k_in_stage = 0
    for e, (stage, before) in enumerate(alloc):
        if stage != cur_stage:
            # stage boundary: everything rendered so far must be in rank 0's planes
            combine()
            if rank == 0:
                renderer.set_state(cur_stage, before, None)
                words = renderer.build_stage()
            words = broadcast_words(words if rank == 0 else None)
            renderer.set_state(stage, before, words)
            cur_stage = stage
            k_in_stage = 0
        if k_in_stage % world == rank:
            renderer.render(before)
        k_in_stage += 1
    combine()
    renderer.set_state(end_stage if end_stage == cur_stage else cur_stage, end_ex, None)
    return cur_stage, end_ex


class DeviceAdaptiveRenderer:
    """Adapter of api.Device for render_adaptive_sharded (the device must have adaptive sampling enabled and a render started)."""

    def __init__(self, dev):
        self.dev = dev

    def set_state(self, stage, executions, words):
        self.dev.set_adaptive_state(stage, executions, words)

    def render(self, before):
        self.dev.render_allocated_execution(before)

    def build_stage(self):
        self.dev.build_adaptive_stage()
        return self.dev.adaptive_words()


def render_adaptive_on_devices(dev, planes, num_executions: int, update_interval: int, rank: int, world: int):
    """The torch.distributed flavour: `planes` is the torch tensor bound to `dev` (bind_frame_planes). Combines with an NCCL (or
    gloo) reduce onto rank 0, broadcasts the stage counts from rank 0."""
    import numpy as np
    import torch
    import torch.distributed as dist

    multi = dist.is_initialized() and world > 1

    def combine():
        dev.sync()
        if multi:
            dist.reduce(planes, dst=0, op=dist.ReduceOp.SUM)
            if rank != 0:
                planes.zero_()
            torch.cuda.synchronize()

    def broadcast_words(words):
        if not multi:
            return words
        st = dev.adaptive_state()
        t = torch.zeros(st["blocks_x"] * st["blocks_y"], dtype=torch.int64, device=planes.device)
        if rank == 0:
            t.copy_(torch.from_numpy(np.ascontiguousarray(words, np.int64).reshape(-1)))
        dist.broadcast(t, src=0)
        return t.cpu().numpy().astype(np.uint32)

    return render_adaptive_sharded(DeviceAdaptiveRenderer(dev), num_executions, update_interval, rank, world, combine, broadcast_words)
