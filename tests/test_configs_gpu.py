"""Image and ray-count parity of the product against the oracle on the BASELINE.json configurations themselves
(SURVEY 8d scenes at their full triangle counts): config 2 (atrium-1M, 5 bounces), config 3 (terrain-10M + 100 000 emissive
triangles, light-tree NEE through a deep tree), config 4 (divergence stress: glossy / translucent / emissive / diffuse materials
hashed per triangle, open ceiling, 8 bounces).

Three comparisons per configuration, identical random numbers on both sides:
  * per path vertex (the sharp one): k_shade / k_trace_shadow against the oracle on the vertices of wavefront iterations 0 and 1;
  * full frame at 480 x 270 (the CPU oracle finishes it in seconds): planes, ray counts of every kind;
  * a centred region of the 1920 x 1080 frame the bench renders (the product renders the whole frame).

Why images of THESE scenes cannot agree to the 70 - 97 dB of the small rooms (test_render_gpu.py), measured on B200 at 2 spp:
atrium-1M 42.5 dB, terrain-10M 51.4 dB, divergence 41.1 dB, with 16 - 28 % of the pixels differing by more than 1e-3. The reference's
light-tree root pass (light_tree.cuh:191-262, ris.cuh:114-151) streams ALL root children (48 on the atrium, 128 on the other two) through
8 reservoir lanes that each re-use ONE 23-bit random number: every child consumes H(p) bits of it (u' = u / p or (u - p) / (1 - p)),
about 16 bits over 48 children, so the last decisions of a lane are taken on 7 or fewer significant bits and a 1-ulp difference in
any importance value (fast-math rcp / rsqrt on the device - as in the reference's own build - against IEEE + libm in the oracle)
is amplified by 2^16. Per vertex about 2 - 3 % of the vertices therefore select a different (equally distributed) light than the
oracle; the per-vertex test below pins exactly that: where the decision agrees the continuous outputs agree to 1e-3, where it
does not the estimate is an equally valid sample, and the sums over all vertices agree to a fraction of a percent. The image
thresholds are set from these measurements: PSNR >= 35 dB (33 dB on the glass-heavy divergence scene), mean radiance within 1 %
(measured 0.06 - 0.7 %), ray counts within 0.5 %.
Pixels poisoned by the reference's NaN quirk (0 x inf in light_bsdf_get_probability, about 1 path in 500 000) are
rare on both sides (at most 1e-4 of the pixels) and are excluded from the image statistics."""
import numpy as np
import pytest

import orc
from luminary_b200 import api, scenes

pytestmark = pytest.mark.gpu

CONFIGS = {
    "config2_atrium1m": dict(make=lambda w, h: scenes.atrium(1_000_000, w, h, 5), spp=2, min_psnr=35.0),
    "config3_terrain10m": dict(make=lambda w, h: scenes.terrain(2236, 50_000, w, h, 5), spp=2, min_psnr=35.0),
    "config4_divergence": dict(make=lambda w, h: scenes.divergence(1_000_000, w, h, 8), spp=2, min_psnr=33.0),
}


def _psnr(a, b):
    a = a / (1.0 + a)
    b = b / (1.0 + b)
    mse = float(np.mean((a - b) ** 2))
    return 150.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)


@pytest.fixture(scope="module")
def luts():
    dev = api.Device(0)
    dev.build_bsdf_lut()
    out = dev.get_bsdf_lut()
    dev.destroy()
    return out


def _compare(name, gpu, ref, spp, min_psnr, mean_tol=1e-2):
    """gpu / ref: (4, h, w) plane sums of the same pixels"""
    g, r = gpu[:3] / spp, ref[:3] / spp
    nan_g, nan_r = ~np.isfinite(g).all(axis=0), ~np.isfinite(r).all(axis=0)
    print(f"  {name}: non-finite pixels gpu {int(nan_g.sum())} oracle {int(nan_r.sum())} (of {nan_g.size})")
    # color_any (math.cuh:944-946) is "any component > 0": a NaN NEE term (about 1e-6 of the path vertices of the atrium: a light
    # sample with a degenerate solid angle) fails it and is dropped, in the reference, the oracle and the product alike
    assert nan_g.sum() == 0 and nan_r.sum() == 0
    ok = ~(nan_g | nan_r)
    g, r = g[:, ok], r[:, ok]
    psnr = _psnr(g, r)
    mean_rel = abs(g.mean() - r.mean()) / r.mean()
    second = abs(gpu[3][ok].mean() - ref[3][ok].mean()) / max(ref[3][ok].mean(), 1e-20)
    diff_px = (np.abs(g - r).max(axis=0) > 1e-3 * (1.0 + r.max(axis=0))).mean()
    print(f"  {name}: PSNR {psnr:.1f} dB, mean {g.mean():.6f} vs {r.mean():.6f} (rel {mean_rel:.2e}), second moment rel {second:.2e}, "
          f"pixels differing by > 1e-3: {100.0 * diff_px:.3f} %")
    assert r.mean() > 1e-3
    assert psnr >= min_psnr
    assert mean_rel <= mean_tol
    return psnr


@pytest.mark.parametrize("cfg", sorted(CONFIGS))
def test_baseline_config_full_frame_480x270(cfg, luts):
    c = CONFIGS[cfg]
    sc = c["make"](480, 270)
    spp = c["spp"]
    dev = api.Device(0)
    dev.set_bsdf_lut(*luts)
    lt = dev.load_scene(sc, light_tree="auto")
    st0 = dev.stats()
    print(f"  {cfg}: {sc.num_tris} triangles, BVH8 {st0['bvh_nodes']} nodes, depth {st0['bvh_depth']}, SAH {st0['bvh_sah_cost']:.3f}, "
          f"PLOC radius {st0['bvh_ploc_radius']}, emitter BVH depth {st0['light_bvh_depth']}")
    dev.start_render()
    dev.render_samples(0, spp)
    gpu = dev.download_frame_planes().reshape(4, sc.height, sc.width)
    st = dev.stats()
    dev.destroy()
    assert st["stack_overflows"] == 0 and st["nonfinite_samples"] == 0
    osc = orc.OracleScene(sc)
    osc.set_bsdf_luts(*luts)
    if lt is not None:
        osc.set_light_tree(*lt)
    ref, info = osc.render(0, spp)
    _compare(cfg, gpu, ref.reshape(4, sc.height, sc.width), spp, c["min_psnr"])
    for mine, theirs in (("closest_rays", "closest_rays"), ("shadow_rays", "shadow_rays"), ("light_rays", "light_enum_rays")):
        a, b = int(st[mine]), int(info[theirs])
        print(f"  {cfg}: {mine} {a} vs oracle {b} (rel {abs(a - b) / max(b, 1):.2e})")
        if mine == "shadow_rays":
            # the oracle counts what the reference EXECUTEs: every segment with a valid light / a non-zero packed ambient colour. The
            # product does not trace a segment whose contribution x path throughput is exactly 0 (paths whose bounce weight was 0
            # stay alive in the reference too - Russian roulette skips delta paths - and their NEE adds exactly nothing): fewer
            # rays, identical image. Measured: 76 % / 99 % / 73 % of the oracle's count.
            assert 0.6 * b <= a <= 1.002 * b
        else:
            assert abs(a - b) <= 2e-3 * max(b, 1000)


@pytest.mark.parametrize("cfg", sorted(CONFIGS))
def test_baseline_config_1080p_region(cfg, luts):
    c = CONFIGS[cfg]
    sc = c["make"](1920, 1080)
    spp = c["spp"]
    dev = api.Device(0)
    dev.set_bsdf_lut(*luts)
    lt = dev.load_scene(sc, light_tree="auto")
    dev.start_render()
    dev.render_samples(0, spp)
    gpu = dev.download_frame_planes().reshape(4, sc.height, sc.width)
    st = dev.stats()
    dev.destroy()
    assert st["stack_overflows"] == 0 and st["nonfinite_samples"] == 0
    osc = orc.OracleScene(sc)
    osc.set_bsdf_luts(*luts)
    if lt is not None:
        osc.set_light_tree(*lt)
    x0, y0, x1, y1 = 800, 450, 1120, 630   # 320 x 180 pixels around the image centre
    ref, _ = osc.render(0, spp, region=(x0, y0, x1, y1))
    ref = ref.reshape(4, sc.height, sc.width)
    assert not ref[:, :y0].any() and not ref[:, :, :x0].any()   # the oracle only touched the region
    _compare(cfg + " region", gpu[:, y0:y1, x0:x1], ref[:, y0:y1, x0:x1], spp, c["min_psnr"])


@pytest.mark.parametrize("cfg", sorted(CONFIGS))
def test_baseline_config_per_vertex(cfg, luts):
    """The surface stages on the path vertices of the configuration itself (480 x 270 frame, wavefront iterations 0 and 1).
    Measured on B200 at iteration 0: the light-tree NEE selects the oracle's light on 91.4 % (atrium, 48 lights), 88.4 % (terrain,
    100 000 lights) and 70.4 % (divergence, 250 000 lights behind a 128-child root and a deep tree) of the vertices - the reservoir
    chaos of the module docstring; tests/test_shade_vertices_gpu.py shows that the REFERENCE's own kernel disagrees with the oracle
    just as often on a many-light scene. Everything that does not hang on that choice agrees on >= 99.5 % (Russian roulette, bounce
    direction, emission); where the light agrees the ray agrees (>= 97 %) and the median colour agrees to 1e-3 (the colour carries the
    reservoir weight of all 8 lanes); summed over ALL vertices the unshadowed NEE energy agrees within 1 % and the energy that
    survives k_trace_shadow within 1.5 % (measured 0.01 - 0.3 %): the differing choices are equally valid samples."""
    from test_shade_vertices_gpu import product_vertices

    c = CONFIGS[cfg]
    sc = c["make"](480, 270)
    dev = api.Device(0)
    dev.set_bsdf_lut(*luts)
    lt = dev.load_scene(sc, light_tree="auto")
    osc = orc.OracleScene(sc)
    osc.set_bsdf_luts(*luts)
    osc.set_light_tree(*lt)
    for iteration in (0, 1):
        vin, _ = osc.path_vertices(1, iteration)
        want = osc.shade_vertices(vin, iteration)
        seg = osc.nee_segments(vin, iteration)
        got = dev.shade_vertices(product_vertices(vin), 1, iteration, False)
        g0, w0 = got["nee"][:, 0], seg[:, 0]
        both = (g0["valid"] != 0) & (w0["valid"] != 0)
        present = ((g0["valid"] != 0) == (w0["valid"] != 0)).mean()
        same = g0["target_prim"][both] == w0["target_prim"][both]
        gs, ws = g0[both][same], w0[both][same]
        ray_ok = np.abs(gs["ray"] - ws["ray"]).max(axis=1) < 1e-3
        col_rel = np.abs(gs["color"][ray_ok] - ws["color"][ray_ok]).max(axis=1) / np.maximum(np.abs(ws["color"][ray_ok]).max(axis=1), 1e-4)
        e_g, e_w = got["nee"]["color"].sum(), seg["color"].sum()
        v_g, v_w = got["nee"]["visible"].sum(), (seg["color"] * seg["visibility"]).sum()
        alive = ((got["alive"] != 0) == (want["bounce_alive"] != 0)).mean()
        m = (got["alive"] != 0) & (want["bounce_alive"] != 0)
        bounce_ray = (np.abs(got["ray"][m] - want["bounce_ray"][m]).max(axis=1) < 2e-3).mean()
        print(f"  {cfg} iter {iteration}: {vin.size} vertices, light-tree segment present equal {present:.4f}, same light {same.mean():.4f}, "
              f"ray equal {ray_ok.mean():.4f}, colour p50 rel {np.percentile(col_rel, 50):.2e}, NEE energy {e_g:.5g} vs {e_w:.5g}, "
              f"visible {v_g:.5g} vs {v_w:.5g}, rr equal {alive:.4f}, bounce ray equal {bounce_ray:.4f}")
        assert present >= 0.95 and same.mean() >= 0.65 and ray_ok.mean() >= 0.97
        assert np.percentile(col_rel, 50) <= 1e-3
        assert abs(e_g - e_w) <= 1e-2 * e_w and abs(v_g - v_w) <= 1.5e-2 * max(v_w, 1e-6)
        assert alive >= 0.995 and bounce_ray >= 0.99
        assert np.abs(got["emission"] - want["emission"]).max() <= 1e-4 * max(1.0, np.abs(want["emission"]).max())
    assert dev.stats()["stack_overflows"] == 0
    dev.destroy()


def test_terrain_closest_hits_bit_exact():
    """Closest hit on the terrain (tiny nodes far from the ray origins: the stress case for the quantisation margin of the BVH8
    child boxes): primary rays and 400 000 long random rays, ids and t bit-identical to the oracle."""
    scene = scenes.terrain(1000, 5000, 480, 270, 2)
    dev = api.Device(0)
    dev.load_scene(scene)
    osc = orc.OracleScene(scene)
    inst, tri, t, u, v = dev.trace_primary(0)
    ref = osc.trace_primary(0)
    assert np.array_equal(inst, ref["instance"]) and np.array_equal(tri, ref["tri"])
    assert np.array_equal(t.view(np.uint32), ref["t"].view(np.uint32))
    rng = np.random.default_rng(5)
    n = 400_000
    o = np.stack([rng.uniform(-500, 500, n), rng.uniform(5, 120, n), rng.uniform(-500, 500, n)], axis=1).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:, 1] = -np.abs(d[:, 1]) * 0.3
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    gi, gtri, gt, gu, gv = dev.trace_rays(o, d)
    r = osc.trace_rays(o, d)
    hit = r["prim"] != 0xFFFFFFFE
    assert hit.mean() > 0.5
    assert np.array_equal(gi == 0xFFFFFFFE, ~hit)
    assert np.array_equal(gt.view(np.uint32), r["t"].view(np.uint32))
    assert np.array_equal(gu.view(np.uint32)[hit], r["u"].view(np.uint32)[hit]) and np.array_equal(gv.view(np.uint32)[hit], r["v"].view(np.uint32)[hit])
    assert dev.stats()["stack_overflows"] == 0
    dev.destroy()
