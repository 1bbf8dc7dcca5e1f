"""World-size-2 gloo test (CPU) of the multi-GPU host logic: sample-id partition + plane reduce.

The renderer stand-in on CPU is the oracle (tests may use it): two ranks render disjoint sample ids of a tiny scene,
reduce with gloo, and rank 0 must hold exactly what a single process renders for all ids."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from luminary_b200 import scenes, sharding


def test_rank_sample_ids_partition_is_exact():
    for total in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, count, stride = sharding.rank_sample_ids(total, r, world)
                seen += [first + k * stride for k in range(count)]
            assert sorted(seen) == list(range(total))
    with pytest.raises(ValueError):
        sharding.rank_sample_ids(8, 2, 2)


def _scene():
    sc = scenes.example_with_light(width=32, height=18, sphere_subdiv=1, max_ray_depth=2)
    return sc


def _render(first, count, stride):
    import orc

    sc = _scene()
    osc = orc.OracleScene(sc)
    from luminary_b200 import api  # host-side C light tree builder (no GPU needed)

    osc.set_light_tree(*api.build_light_tree(sc))
    osc.set_bsdf_luts(np.full(1024, 60000, np.uint16), np.full(1024, 3000, np.uint16), np.full(32768, 65535, np.uint16),
                      np.full(32768, 65535, np.uint16))
    planes = np.zeros((4, sc.height, sc.width), np.float32)
    for k in range(count):
        p, _ = osc.render(first + k * stride, 1, threads=1)
        planes += p
    return planes


def _worker(rank, world, port, total, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count, stride = sharding.rank_sample_ids(total, rank, world)
    planes = torch.from_numpy(_render(first, count, stride).copy())
    n = sharding.reduce_planes(planes, count, dst=0)
    assert n == total
    if rank == 0:
        np.save(out_path, planes.numpy())
    dist.destroy_process_group()


def test_two_rank_reduce_equals_single_process(tmp_path):
    total = 5
    out = str(tmp_path / "planes.npy")
    mp.spawn(_worker, args=(2, 29517, total, out), nprocs=2, join=True)
    combined = np.load(out)
    single = _render(0, total, 1)
    assert np.allclose(combined, single, rtol=1e-5, atol=1e-6)
    assert single[:3].sum() > 0
