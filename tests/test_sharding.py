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


# ---------------------------------------------------------------------------------------------
# adaptive sampling across ranks (sharding.render_adaptive_sharded): the oracle is the renderer on CPU
# ---------------------------------------------------------------------------------------------
def test_adaptive_allocator_schedule():
    alloc, stage, ex = sharding.adaptive_allocations(2 + 4 + 8 + 3, 2)
    assert stage == 3 and ex == [2, 4, 8, 3, 0]
    assert alloc[0] == (0, [0, 0, 0, 0, 0]) and alloc[2] == (1, [2, 0, 0, 0, 0]) and alloc[6] == (2, [2, 4, 0, 0, 0])
    assert alloc[-1] == (3, [2, 4, 8, 2, 0])
    alloc, stage, ex = sharding.adaptive_allocations(1000, 1)  # 1 + 2 + 4 + 8 executions, then stage 4 for ever
    assert stage == 4 and ex == [1, 2, 4, 8, 985]


class _OracleAdaptiveRenderer:
    def __init__(self, params):
        import orc

        sc = _scene()
        self.orc = orc
        self.osc = orc.OracleScene(sc)
        from luminary_b200 import api

        self.osc.set_light_tree(*api.build_light_tree(sc))
        self.osc.set_bsdf_luts(np.full(1024, 60000, np.uint16), np.full(1024, 3000, np.uint16), np.full(32768, 65535, np.uint16),
                               np.full(32768, 65535, np.uint16))
        self.w, self.h = sc.width, sc.height
        self.params = params
        self.no_build = orc.AdaptiveParams.from_buffer_copy(params)
        self.no_build.update_interval = 1 << 20  # a single execution never triggers a stage build of its own
        self.planes = np.zeros((4, self.h, self.w), np.float32)
        self.words = np.zeros(((self.h + 3) // 4, (self.w + 3) // 4), np.uint32)
        self.stage, self.ex = 0, [0] * 5

    def set_state(self, stage, executions, words):
        self.stage, self.ex = stage, list(executions)
        if words is not None:
            self.words = np.array(words, np.uint32).reshape(self.words.shape)

    def render(self, before):
        st = dict(planes=self.planes, words=self.words.copy(), executions=list(before), stage=self.stage, paths=0, closest_rays=0, shadow_rays=0,
                  light_enum_rays=0)
        self.planes = self.osc.render_adaptive(self.no_build, 1, st, threads=1)["planes"]

    def build_stage(self):
        self.words, _, _ = self.orc.adaptive_stage_counts(self.planes, self.w, self.h, self.words, self.ex, self.stage, self.params)
        return self.words


def _adaptive_params():
    import orc

    return orc.adaptive_params(max_sampling_rate=6, avg_sampling_rate=2, update_interval=2, exposure_aware=True, exposure=1.0, tonemap=4)


def _adaptive_worker(rank, world, port, n_exec, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = _OracleAdaptiveRenderer(_adaptive_params())

    def combine():
        t = torch.from_numpy(r.planes)
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
        if rank != 0:
            r.planes[...] = 0.0

    def broadcast_words(words):
        t = torch.from_numpy(np.ascontiguousarray(words, np.int64)) if words is not None else torch.zeros(r.words.shape, dtype=torch.int64)
        dist.broadcast(t, src=0)
        return t.numpy().astype(np.uint32)

    stage, ex = sharding.render_adaptive_sharded(r, n_exec, 2, rank, world, combine, broadcast_words)
    if rank == 0:
        np.savez(out_path, planes=r.planes, words=r.words, stage=stage, ex=np.array(ex))
    dist.destroy_process_group()


def test_two_rank_adaptive_schedule_equals_single_process(tmp_path):
    n_exec = 2 + 4 + 3
    out = str(tmp_path / "adaptive.npz")
    mp.spawn(_adaptive_worker, args=(2, 29519, n_exec, out), nprocs=2, join=True)
    got = np.load(out)
    single = _OracleAdaptiveRenderer(_adaptive_params())
    st = single.osc.render_adaptive(single.params, n_exec, threads=1)
    assert list(got["ex"]) == st["executions"] == [2, 4, 3, 0, 0] and int(got["stage"]) == st["stage"] == 2
    # identical sample ids and counts; only the float summation order across ranks differs
    assert np.array_equal(got["words"], st["words"])
    assert np.allclose(got["planes"], st["planes"], rtol=2e-5, atol=1e-6)
    assert ((st["words"] & 0xFF).max() > 0)
