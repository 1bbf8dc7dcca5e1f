"""GPU parity of material textures: the texture fetch, alpha cut-outs in the closest-hit any-hit,
textured shadow transparency, albedo / roughness / normal / luminance maps in shading.

Tolerances.
  * Fetch: the reference samples with the hardware texture unit (tex2DLod, cuda/texture_utils.cuh:36). The CPU
    restatement (oracle/orc_texture.c) follows the weight / rounding rules measured on the B200 (tools/tex_probe.py):
    unorm (u8 / u16) textures with power-of-two extents must agree BIT FOR BIT; for other extents the unit's coordinate
    precision is not published: >= 98 % of the linear fetches bit-identical, the rest within one 1/256 weight step; fp32
    textures within 2 ulp-of-range (the order of the four fp32 multiply-adds inside the unit is not observable); point
    filter and texel centres exact.
  * Closest-hit ids: a hit is cut out iff the fetched alpha is exactly 0. Away from block edges this is exact;
    within a texel of an alpha edge the 8-bit weight rounding may differ, so <= 0.2 % of the pixels may differ
    (documented exception, in addition to the exact-t ties of test_trace_gpu.py). Where ids agree, t is bit-identical.
  * Images: as test_render_gpu.py (PSNR >= 30 dB on x / (1 + x), mean radiance within 2 %).
"""
import numpy as np
import pytest

import orc
from luminary_b200 import scenes

pytestmark = pytest.mark.gpu


def _psnr(a, b):
    a = a / (1.0 + a)
    b = b / (1.0 + b)
    mse = float(np.mean((a - b) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)


@pytest.fixture(scope="module")
def device_luts():
    from luminary_b200 import api

    dev = api.Device(0)
    dev.build_bsdf_lut()
    luts = dev.get_bsdf_lut()
    dev.destroy()
    return luts


def test_texture_fetch_matches_oracle():
    from luminary_b200 import api

    scene = scenes.textured_example()
    textures = [t for t in scene.textures if t.get("data") is not None]
    # one more: every address mode on a tiny 2-component u16 texture
    rng = np.random.default_rng(7)
    tiny = (rng.random((3, 5, 2)) * 65535).astype(np.uint16)
    for wrap in range(4):
        textures.append(dict(data=tiny, wrap_u=wrap, wrap_v=(wrap + 1) % 4, filter=1, gamma=1.0))
    # widths / heights that are not powers of two, every address mode, both filters
    for k, (w, h) in enumerate(((7, 3), (100, 37), (1000, 3), (33, 129))):
        odd = (rng.random((h, w, 4)) * 255).astype(np.uint8)
        for wrap in range(4):
            textures.append(dict(data=odd, wrap_u=wrap, wrap_v=(wrap + k) % 4, filter=(wrap + k) & 1, gamma=1.0))
    dev = api.Device(0, load_embedded_data=False)
    dev.add_textures(textures)
    uv = (rng.random((4096, 2)) * 4.0 - 1.5).astype(np.float32)
    failures = []
    for tid, t in enumerate(textures):
        h, w = t["data"].shape[:2]
        centres = np.stack([(np.arange(64) % w + 0.5) / w, (np.arange(64) // w % h + 0.5) / h], axis=1).astype(np.float32)
        got_c = dev.sample_texture(tid, centres)
        ref_c = orc.texture_fetch(t, centres)
        assert np.array_equal(got_c, ref_c), f"texture {tid}: texel centres differ"
        got = dev.sample_texture(tid, uv)
        ref = orc.texture_fetch(t, uv)
        diff = np.abs(got - ref).max()
        exact = float(np.mean(got == ref))
        print(f"texture {tid} ({w}x{h}x{t['data'].shape[2] if t['data'].ndim == 3 else 1}, wrap {t.get('wrap_u')}/{t.get('wrap_v')}, "
              f"filter {t.get('filter')}): max |diff| {diff:.3e}, bit-identical {100 * exact:.1f} %")
        pow2 = (w & (w - 1)) == 0 and (h & (h - 1)) == 0
        if t["data"].dtype != np.float32 and not pow2:
            if exact < 0.98 or diff > 1.05 / 256.0:
                failures.append((tid, exact, diff))
        elif t["data"].dtype != np.float32:
            if exact != 1.0:
                bad = np.where(np.any(got != ref, axis=1))[0]
                failures.append((tid, exact, uv[bad[:4]].tolist(), got[bad[:4]].tolist(), ref[bad[:4]].tolist()))
        elif diff > 4e-7 * max(1.0, float(np.abs(ref).max())):
            failures.append((tid, diff))
    dev.destroy()
    assert not failures, failures


def test_generated_mip_chain_matches_oracle():
    """Lumb200Texture.mipmap = 1: the chain generated on the device (k_mipmap_level through surface writes) against
    orc_texture_next_mip, level by level, read back with tex2DLod at the texel centres. Power-of-two extents: bit-identical."""
    from luminary_b200 import api

    rng = np.random.default_rng(4)
    u8 = (rng.random((64, 128, 4)) * 255).astype(np.uint8)
    u8[8:24, 8:24, 3] = 0
    u8[40:44, 40:44, 3] = (rng.random((4, 4)) < 0.3).astype(np.uint8)  # sparse alpha 1: the opacity rule keeps it non-zero
    u16 = (rng.random((32, 32, 4)) * 65535).astype(np.uint16)
    f32 = rng.random((16, 64, 4)).astype(np.float32)
    textures = [dict(data=u8, wrap_u=0, wrap_v=0, filter=1, mipmap=1), dict(data=u16, wrap_u=1, wrap_v=2, filter=1, mipmap=1),
                dict(data=f32, wrap_u=1, wrap_v=1, filter=1, mipmap=1), dict(data=u8, wrap_u=0, wrap_v=0, filter=1, mipmap=0)]
    dev = api.Device(0, load_embedded_data=False)
    dev.add_textures(textures)
    for tid, t in enumerate(textures[:3]):
        levels = orc.texture_mip_chain(t)
        assert len(levels) >= 4
        for l, ref in enumerate(levels):
            h, w = ref.shape[:2]
            yy, xx = np.mgrid[0:h, 0:w]
            centres = np.stack([(xx.reshape(-1) + 0.5) / w, (yy.reshape(-1) + 0.5) / h], axis=1).astype(np.float32)
            got = dev.sample_texture(tid, centres, lod=float(l)).reshape(h, w, 4)
            if ref.dtype == np.float32:
                assert np.abs(got - ref).max() <= 2e-7, (tid, l)
            else:
                full = np.float32(255.0 if ref.dtype == np.uint8 else 65535.0)
                assert np.array_equal(got, ref.astype(np.float32) / full), (tid, l, np.abs(got * full - ref).max())
    # a lod beyond the chain clamps to the last level; a texture without a chain ignores the lod
    last = orc.texture_mip_chain(textures[0])[-1]
    got = dev.sample_texture(0, np.array([[0.3, 0.6]], np.float32), lod=50.0)
    ref = orc.texture_fetch(dict(textures[0], data=last), np.array([[0.3, 0.6]], np.float32))
    assert np.array_equal(got, ref)
    uv = rng.random((64, 2)).astype(np.float32)
    assert np.array_equal(dev.sample_texture(3, uv, lod=3.0), dev.sample_texture(3, uv, lod=0.0))
    dev.destroy()


def test_alpha_cutout_closest_hit_ids():
    from luminary_b200 import api

    scene = scenes.textured_example(width=384, height=216)
    dev = api.Device(0)
    dev.load_scene(scene)
    inst, tri, t, u, v = dev.trace_primary(3)
    dev.destroy()
    ref = orc.OracleScene(scene).trace_primary(3)
    same = (inst == ref["instance"]) & (tri == ref["tri"])
    print(f"closest-hit ids equal on {100 * same.mean():.3f} % of {same.size} pixels; "
          f"screen hits {int((inst == 1).sum())}, seen through cut-outs {int(((ref['instance'] != 1)).sum())}")
    assert same.mean() >= 0.998
    assert np.array_equal(t[same].view(np.uint32), ref["t"][same].view(np.uint32))
    # the cut-outs are really there: without the alpha test every pixel of the screen's footprint would hit instance 1
    plain = scenes.textured_example(width=384, height=216)
    plain.textures = []
    for m in plain.materials:
        for k in ("albedo_tex", "luminance_tex", "roughness_tex", "metallic_tex", "normal_tex"):
            m[k] = 0xFFFF
    ref_plain = orc.OracleScene(plain).trace_primary(3)
    assert int((ref_plain["instance"] == 1).sum()) > int((ref["instance"] == 1).sum()) * 1.2


def test_textured_hits_match_golden_fixture():
    """The CUDA path against the committed fixture (tests/golden/textured_adaptive.json), independent of the oracle library on the box."""
    import hashlib
    import json
    import os

    from luminary_b200 import api

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "textured_adaptive.json")))["textured_hits"]
    dev = api.Device(0)
    dev.load_scene(scenes.textured_example(width=384, height=216))
    inst, tri, t, u, v = dev.trace_primary(3)
    dev.destroy()
    d = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert d(inst) == g["sha256"]["instance"] and d(tri) == g["sha256"]["tri"] and d(t.view(np.uint32)) == g["sha256"]["t"]


def _render_both(scene, spp, device_luts):
    from luminary_b200 import api

    dev = api.Device(0)
    dev.set_bsdf_lut(*device_luts)
    dev.load_scene(scene, light_tree=None)
    lt = dev.build_light_tree(scene)  # luminance-textured emitters: intensities integrated on the device
    dev.update_light_tree(*lt)
    dev.build_accel()
    dev.start_render()
    dev.render_samples(0, spp)
    gpu = dev.download_frame_planes()
    stats = dev.stats()
    dev.destroy()
    osc = orc.OracleScene(scene)
    osc.set_bsdf_luts(*device_luts)
    osc.set_light_tree(*lt)
    ref, info = osc.render(0, spp)
    return gpu, ref, stats, info


def test_textured_room_image_parity(device_luts):
    scene = scenes.textured_example(width=160, height=90, max_ray_depth=4)
    spp = 32
    gpu, ref, stats, info = _render_both(scene, spp, device_luts)
    assert np.isfinite(gpu).all()
    g, r = gpu[:3] / spp, ref[:3] / spp
    print(f"textured room: mean gpu {g.mean():.5f} oracle {r.mean():.5f}, PSNR {_psnr(g, r):.1f} dB, rays {stats['closest_rays']} / "
          f"{info['closest_rays']} closest, {stats['shadow_rays']} / {info['shadow_rays']} shadow")
    assert r.mean() > 0.005
    assert abs(g.mean() - r.mean()) <= 0.02 * r.mean()
    assert _psnr(g, r) >= 30.0
    assert abs(int(stats["closest_rays"]) - info["closest_rays"]) <= 0.005 * info["closest_rays"]
    assert abs(int(stats["shadow_rays"]) - info["shadow_rays"]) <= 0.01 * info["shadow_rays"]


def test_textured_emitter_intensities_match_oracle():
    """lumb200_device_compute_light_intensities (the reference's light_compute_intensity kernel) against its CPU restatement:
    fast-math powf of the gamma vs libm => 2e-4 relative; and the intensities reach the light tree."""
    from luminary_b200 import api

    scene = scenes.textured_example()
    # a second, larger emitter with a gamma-encoded 8-bit luminance map and skewed texture coordinates
    rng = np.random.default_rng(11)
    scene.textures.append(dict(data=(rng.random((64, 32, 4)) * 255).astype(np.uint8), wrap_u=0, wrap_v=2, filter=1, gamma=2.2))
    scene.materials.append(scenes.default_material(emission=(1.0, 1.0, 1.0), emission_scale=5.0, emission_active=True,
                                                   luminance_tex=len(scene.textures) - 1))
    quad = scenes.quad_uv((-1.9, 1.0, -3.9), (-1.9, 2.0, -3.9), (-1.9, 2.0, -2.9), (-1.9, 1.0, -2.9), len(scene.materials) - 1, 0.37)
    quad.uv[1, 2] = (0.9, -0.2)
    scene.meshes.append(quad)
    scene.instances.append(scenes.Instance(len(scene.meshes) - 1))
    mesh_ids, tri_ids = api.textured_emitter_triangles(scene)
    assert mesh_ids.size == 4
    dev = api.Device(0)
    dev.load_scene(scene, light_tree=None)
    got = dev.compute_light_intensities(mesh_ids, tri_ids)
    lt = dev.build_light_tree(scene)
    lt_plain = api.build_light_tree(scene)
    dev.destroy()
    ref = orc.light_intensities(orc.OracleScene(scene), mesh_ids, tri_ids)
    print("intensities", got, ref)
    assert np.all(ref > 0.2) and np.all(ref <= 1.0)
    assert np.allclose(got, ref, rtol=2e-4, atol=0)
    assert lt[2].shape == lt_plain[2].shape and lt[0] != lt_plain[0]  # same lights, different powers in the root


def test_textures_change_the_image(device_luts):
    """The textured kernel variants are really the ones that ran: stripping the textures changes the render."""
    from luminary_b200 import api

    def render(scene):
        lt = api.build_light_tree(scene)
        dev = api.Device(0)
        dev.set_bsdf_lut(*device_luts)
        dev.load_scene(scene, light_tree=lt)
        dev.start_render()
        dev.render_samples(0, 8)
        img = dev.download_frame_planes()[:3] / 8
        dev.destroy()
        return img

    a = render(scenes.textured_example(width=96, height=54))
    plain = scenes.textured_example(width=96, height=54)
    plain.textures = []
    for m in plain.materials:
        for k in ("albedo_tex", "luminance_tex", "roughness_tex", "metallic_tex", "normal_tex"):
            m[k] = 0xFFFF
    b = render(plain)
    assert _psnr(a, b) < 28.0
