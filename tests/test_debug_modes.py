"""Debug shading modes (LuminaryShadingMode 1..5): the reference's one-bounce queue (device_renderer.c:136-182) with
geometry_process_tasks_debug (cuda/geometry.cuh:182-246) and sky_process_tasks_debug (cuda/sky.cuh:635-668).

CPU part: the oracle's restatement against closed forms (identification colour = Squares hash of the hit ids, depth = saturate(2 / t)).
GPU part: the product's k_shade_debug through the C ABI against the oracle on the same sample ids. Identification is integer work and
must be bit-exact; depth / albedo / normal / lights go through --use_fast_math on the device (as in the reference), tolerances below.
"""
import numpy as np
import pytest

import orc
from luminary_b200 import scenes

MODES = {1: "albedo", 2: "depth", 3: "normal", 4: "identification", 5: "lights"}


def _swap16(x):
    return ((x << np.uint32(16)) | (x >> np.uint32(16))).astype(np.uint32)


def _squares32(key, counter):  # random_uint32_t_base, cuda/random.cuh:172-194 (uint32 arithmetic wraps)
    key = np.uint32(key)
    with np.errstate(over="ignore"):
        y = (counter.astype(np.uint32) * key).astype(np.uint32)
        x = y.copy()
        z = (y + key).astype(np.uint32)
        x = _swap16((x * x + y).astype(np.uint32))
        x = _swap16((x * x + z).astype(np.uint32))
        x = _swap16((x * x + y).astype(np.uint32))
        x = (x * x + z).astype(np.uint32)
        zz = x.copy()
        x = _swap16(x)
        return (zz ^ (x * x + y).astype(np.uint32)).astype(np.uint32)


def _scene():
    """The lit box seen from outside through its (two-sided) walls, so that the frame holds hits AND misses; a coloured constant sky."""
    sc = scenes.example_with_light(width=96, height=54, sphere_subdiv=2, max_ray_depth=3)
    sc.camera = dict(sc.camera, pos=(0.3, 1.6, 5.0))
    sc.sky_color = (0.25, 0.5, 0.75)
    return sc


def test_oracle_identification_and_depth_closed_forms():
    sc = _scene()
    osc = orc.OracleScene(sc)
    prim = osc.trace_primary(0)
    hit = prim["instance"] < 0xFFFFFFF0  # misses carry the sky handle
    assert hit.any() and (~hit).any()

    ident = osc.render_debug(4, 0, 1)
    v = _squares32(0x55555555, ((prim["instance"].astype(np.uint32) << np.uint32(16)) | prim["tri"].astype(np.uint32)).astype(np.uint32))
    want = np.stack([(v & 0x7FF).astype(np.float32) / np.float32(0x7FF), ((v >> 10) & 0x7FF).astype(np.float32) / np.float32(0x7FF),
                     ((v >> 20) & 0x7FF).astype(np.float32) / np.float32(0x7FF)])
    sky = np.array([0.0, 0.63, 1.0], np.float32)[:, None]
    want = np.where(hit.reshape(1, -1), want.reshape(3, -1), sky)
    assert np.array_equal(ident[:3].reshape(3, -1), want)

    depth = osc.render_debug(2, 0, 1)
    t = prim["t"].reshape(-1)
    with np.errstate(divide="ignore"):
        d = np.clip((np.float32(1.0) / t) * np.float32(2.0), 0.0, 1.0).astype(np.float32)
    d = np.where(hit.reshape(-1), d, 0.0)
    for c in range(3):
        assert np.array_equal(depth[c].reshape(-1), d)
    # plane 3 accumulates luminance(colour^2) like the beauty pass
    assert (depth[3].reshape(-1)[~hit.reshape(-1)] == 0).all() and (depth[3].reshape(-1)[hit.reshape(-1)] > 0).all()

    # albedo + emission; misses show the sky itself (sky_color_main with the camera state); lights mode keeps 2.5 % of the albedo
    albedo = osc.render_debug(1, 0, 1)
    assert np.array_equal(albedo[:3].reshape(3, -1)[:, ~hit.reshape(-1)], np.repeat(np.array(sc.sky_color, np.float32)[:, None], (~hit).sum(), 1))
    lights = osc.render_debug(5, 0, 1)
    normal = osc.render_debug(3, 0, 1)
    assert np.all(lights[:3] <= albedo[:3] + 1e-6)
    assert np.all(lights[:3].reshape(3, -1)[:, ~hit.reshape(-1)] == 0) and np.all(normal[:3].reshape(3, -1)[:, ~hit.reshape(-1)] == 0)
    assert normal[:3].min() >= 0.0 and normal[:3].max() <= 1.0
    dark = (albedo[:3].reshape(3, -1) <= 1.0).all(axis=0) & hit.reshape(-1)
    assert np.allclose(lights[:3].reshape(3, -1)[:, dark], albedo[:3].reshape(3, -1)[:, dark] * 0.025, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("textured", [False, True])
def test_debug_modes_match_oracle(textured):
    from luminary_b200 import api

    sc = scenes.textured_example(width=96, height=54, sphere_subdiv=2, max_ray_depth=3) if textured else _scene()
    dev = api.Device(0)
    dev.build_bsdf_lut()
    dev.load_scene(sc, light_tree="auto")
    osc = orc.OracleScene(sc)
    spp = 2
    try:
        for mode, name in MODES.items():
            dev.set_shading_mode(mode)
            dev.start_render()
            dev.render_samples(0, spp)
            gpu = dev.download_frame_planes()
            cpu = osc.render_debug(mode, 0, spp)
            err = np.abs(gpu[:3] - cpu[:3])
            print(f"debug mode {name} (textured={textured}): max |gpu - oracle| {err.max():.3e}, mean {err.mean():.3e}, image mean {cpu[:3].mean():.4f}")
            assert cpu[:3].max() > 0
            if mode == 4:
                # the 11-bit hash fields are integer work and must agree exactly; the division by 0x7ff is a fast-math reciprocal
                # multiply on the device (as in the reference's build) and may differ from the oracle's IEEE division by one ulp
                q = lambda a: np.rint(a[:3] / spp * 0x7FF).astype(np.int64)
                assert np.array_equal(q(gpu), q(cpu)) and err.max() <= 2.4e-7, "identification colours differ"
            else:
                # fast-math reciprocal / normalisation on the device; alpha cut-out edges of the textured room may pick the other
                # side of a texel boundary for a handful of pixels
                close = err <= 2e-4 * np.maximum(1.0, np.abs(cpu[:3]))
                assert close.mean() >= (0.998 if textured else 1.0), (name, err.max(), close.mean())
        # the output chain leaves debug images un-tone-mapped (tonemap.cuh:207-208): no exposure, no tone map, only the sRGB transfer
        # and the 8-bit quantisation of convert_RGBF_to_ARGB8 (kernels.cuh:615-644)
        dev.set_shading_mode(4)
        dev.start_render()
        dev.render_samples(0, 1)
        img = dev.download_output_argb8(1, exposure=3.0, tonemap=4, dithering=False)
        planes = dev.download_frame_planes()
        lin = planes[:3].astype(np.float64)
        srgb = np.where(lin <= 0.0031308, 12.92 * lin, 1.055 * np.power(np.maximum(lin, 1e-30), 0.416666666667) - 0.055)
        want = np.floor(np.clip(0.5 + 255.0 * srgb, 0.0, 255.9999))
        rgb = np.stack([img[..., 2], img[..., 1], img[..., 0]]).astype(np.float64)  # LuminaryARGB8 is b, g, r, a
        assert np.abs(rgb - want).max() <= 1 and (rgb == want).mean() > 0.99 and (img[..., 3] == 255).all()
        # and back to the path tracer
        dev.set_shading_mode(0)
        dev.start_render()
        dev.render_samples(0, 1)
        assert dev.stats()["shadow_rays"] > 0
    finally:
        dev.destroy()
