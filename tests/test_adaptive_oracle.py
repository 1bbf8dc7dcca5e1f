"""CPU checks of the adaptive-sampling restatement (oracle/orc_adaptive.c + orc_render_adaptive; reference
cuda/adaptive_sampling.cuh, device/device_adaptive_sampler.c, device_renderer.c:350-376)."""
import numpy as np

import orc
from luminary_b200 import api, scenes


def test_stage_counts_follow_the_variance():
    w, h = 32, 20
    bw, bh = 8, 5
    n = w * h
    rng = np.random.default_rng(3)
    ex = [8, 0, 0, 0, 0]
    p = orc.adaptive_params(max_sampling_rate=16, avg_sampling_rate=2, update_interval=8, exposure_aware=False)
    # first moments 0, second moment = 8 * variance: per-pixel variance is what we put in
    var = rng.random((h, w)).astype(np.float32)
    planes = np.zeros((4, h, w), np.float32)
    planes[3] = var * 8.0
    words, bvar, total = orc.adaptive_stage_counts(planes, w, h, np.zeros((bh, bw), np.uint32), ex, 0, p)
    blockmax = var.reshape(bh, 4, bw, 4).max(axis=(1, 3))
    assert np.allclose(bvar, blockmax, rtol=1e-6)  # max over the 16 pixels of the block (adaptive_sampling.cuh:189)
    assert abs(total - blockmax.sum()) < 1e-3
    expect = np.clip(np.floor(blockmax / blockmax.mean() * 2.0 + 0.5), 1, 16).astype(np.uint32)
    counts = (words & 0xFF) + 1
    assert np.array_equal(counts, expect)  # remap(variance, 0, avg, 0, avg_rate) rounded, clamped to [1, max] (:212-215)
    assert np.all(words >> 8 == 0)
    # the next stage fills byte 1 and keeps byte 0; samples per pixel so far = 8 + 5 * count0
    ex2 = [8, 5, 0, 0, 0]
    planes2 = planes.copy()
    words2, _, _ = orc.adaptive_stage_counts(planes2, w, h, words, ex2, 1, p)
    assert np.array_equal(words2 & 0xFF, words & 0xFF) and np.any((words2 >> 8) & 0xFF)
    # no variance anywhere: 0 / 0 -> NaN -> 0 -> clamped to one sample (the conversion of the reference kernel)
    flat, _, _ = orc.adaptive_stage_counts(np.zeros((4, h, w), np.float32), w, h, np.zeros((bh, bw), np.uint32), ex, 0, p)
    assert np.all(flat == 0)
    # uniform variance: every block gets the average rate
    planes3 = np.zeros((4, h, w), np.float32)
    planes3[3] = 4.0
    uni, _, _ = orc.adaptive_stage_counts(planes3, w, h, np.zeros((bh, bw), np.uint32), ex, 0, p)
    assert np.all((uni & 0xFF) + 1 == 2)
    # exposure-aware: the variance of bright blocks is compressed by the tone map (Reinhard: 1 / (1 + lum)) and weighs less
    planes4 = np.zeros((4, h, w), np.float32)
    planes4[:3, :, :16] = 8 * 0.2   # dark half: mean 0.2, variance 0.3
    planes4[:3, :, 16:] = 8 * 50.0  # bright half: mean 50, variance 1.0
    planes4[3, :, :16] = 8 * (0.3 + 0.2 ** 2)
    planes4[3, :, 16:] = 8 * (1.0 + 50.0 ** 2)
    pe = orc.adaptive_params(max_sampling_rate=16, avg_sampling_rate=2, update_interval=8, exposure_aware=True, exposure=1.0, tonemap=2)
    we, _, _ = orc.adaptive_stage_counts(planes4, w, h, np.zeros((bh, bw), np.uint32), ex, 0, pe)
    wn, _, _ = orc.adaptive_stage_counts(planes4, w, h, np.zeros((bh, bw), np.uint32), ex, 0, p)
    ce, cn = (we & 0xFF) + 1, (wn & 0xFF) + 1
    assert np.all(cn[:, :4] == 1) and np.all(cn[:, 4:] == 3)  # absolute variance: the bright half wins
    assert np.all(ce[:, :4] == 4) and np.all(ce[:, 4:] == 1)  # displayed variance: the dark half wins


def test_adaptive_render_schedule_and_unbiasedness():
    sc = scenes.example_with_light(width=48, height=28, sphere_subdiv=2, max_ray_depth=2)
    lt = api.build_light_tree(sc)
    osc = orc.OracleScene(sc)
    osc.set_light_tree(*lt)
    p = orc.adaptive_params(max_sampling_rate=8, avg_sampling_rate=2, update_interval=2, exposure=1.0, tonemap=4)
    st = osc.render_adaptive(p, 2)
    assert st["stage"] == 1 and st["executions"] == [2, 0, 0, 0, 0] and st["paths"] == 2 * 48 * 28
    st = osc.render_adaptive(p, 3, st)  # stage 1 lasts 2 << 1 = 4 executions
    assert st["stage"] == 1 and st["executions"] == [2, 3, 0, 0, 0]
    c0 = (st["words"] & 0xFF) + 1
    assert st["paths"] == 2 * 48 * 28 + 3 * int(c0.sum()) * 16
    st = osc.render_adaptive(p, 1 + 8 + 16 + 3, st)
    assert st["stage"] == 4 and st["executions"] == [2, 4, 8, 16, 3]
    # stage 0 of the adaptive schedule is the uniform render: identical planes
    fresh = osc.render_adaptive(p, 2)
    uni, _ = osc.render(0, 2)
    assert np.array_equal(fresh["planes"], uni)
    # means agree with a uniform render of comparable cost (same estimator, different sample allocation)
    img = osc.adaptive_resolve(st)
    ref, _ = osc.render(0, 48)
    ref = ref[:3] / 48
    assert abs(img.mean() - ref.mean()) < 0.08 * ref.mean()
    # per-pixel sample counts: every pixel of a block received executions . counts samples
    n = st["executions"][0] + sum(st["executions"][k + 1] * (((st["words"] >> (8 * k)) & 0xFF) + 1) for k in range(4))
    assert n.min() >= 2 + 4 + 8 + 16 + 3 and n.max() <= 2 + 8 * (4 + 8 + 16 + 3)


def test_adaptive_words_match_golden_fixture():
    """tests/golden/textured_adaptive.json (generator: tests/golden/make_textured_adaptive.py): stage sample counts after 2 + 4 executions."""
    import hashlib
    import json
    import os

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "textured_adaptive.json")))["adaptive"]
    sc = scenes.example_with_light(width=48, height=28, sphere_subdiv=2, max_ray_depth=2)
    osc = orc.OracleScene(sc)
    osc.set_light_tree(*api.build_light_tree(sc))
    st = osc.render_adaptive(orc.adaptive_params(max_sampling_rate=8, avg_sampling_rate=2, update_interval=2, exposure=1.0, tonemap=4), 2 + 4)
    assert st["executions"] == g["executions"] and st["stage"] == g["stage"] and st["paths"] == g["paths"]
    assert hashlib.sha256(np.ascontiguousarray(st["words"]).tobytes()).hexdigest() == g["words_sha256"]
    assert np.bincount(((st["words"] & 0xFF) + 1).reshape(-1), minlength=9).tolist() == g["count_histogram_stage1"]
