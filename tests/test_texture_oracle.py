"""CPU checks of the texture restatement (oracle/orc_texture.c): hand-computed known answers of the published CUDA
filtering rule (programming guide, "Texture Fetching": xB = x - 0.5, i = floor(xB), 1.8 fixed-point weights), the
address modes, and the alpha cut-out semantics of optix_alpha_test (cuda/optix_common.cuh:20-46)."""
import numpy as np

import orc
from luminary_b200 import scenes


def test_point_and_linear_known_answers():
    data = np.array([[[0], [255]], [[51], [102]]], np.uint8)  # 2x2, one component
    t = dict(data=data, wrap_u=1, wrap_v=1, filter=1, gamma=1.0)
    # texel centres return the texel; missing components read 0 (measured on B200, alpha included)
    c = orc.texture_fetch(t, np.array([[0.25, 0.25], [0.75, 0.25], [0.25, 0.75], [0.75, 0.75]], np.float32))
    assert np.array_equal(c[:, 0], np.array([0, 255, 51, 102], np.float32) / np.float32(255.0))
    assert np.all(c[:, 1:] == 0)
    # midway between the two texels of the first row: weight 0.5 exactly
    m = orc.texture_fetch(t, np.array([[0.5, 0.25]], np.float32))
    assert m[0, 0] == np.float32(32768) / np.float32(65535)  # 128/256 of 65535, rounded to a 16-bit unorm
    # the centre of the texture: mean of the four texels
    m = orc.texture_fetch(t, np.array([[0.5, 0.5]], np.float32))
    assert abs(m[0, 0] - (0 + 255 + 51 + 102) / 4 / 255.0) < 1e-6
    # weights are quantised to 1/256: u = 0.25 + 0.3/2 -> frac 0.3 -> round(76.8) / 256 = 77 / 256
    m = orc.texture_fetch(t, np.array([[0.25 + 0.15, 0.25]], np.float32))
    assert abs(m[0, 0] - 77.0 / 256.0) < 2e-5
    # clamp: outside the texture the edge texel repeats
    m = orc.texture_fetch(t, np.array([[-3.0, 0.25], [7.0, 0.25]], np.float32))
    assert m[0, 0] == 0.0 and m[1, 0] == 1.0
    tp = dict(t, filter=0)
    m = orc.texture_fetch(tp, np.array([[0.49, 0.1], [0.51, 0.1]], np.float32))
    assert m[0, 0] == 0.0 and m[1, 0] == 1.0


def test_address_modes():
    row = np.arange(4, dtype=np.float32).reshape(1, 4, 1)  # 4x1 fp32: 0 1 2 3
    uv = lambda xs: np.array([[x, 0.5] for x in xs], np.float32)
    centres = lambda idx: [(i + 0.5) / 4 for i in idx]
    wrap = orc.texture_fetch(dict(data=row, wrap_u=0, wrap_v=1, filter=0), uv(centres([-1, 4, 5, -4])))
    assert wrap[:, 0].tolist() == [3, 0, 1, 0]
    mirror = orc.texture_fetch(dict(data=row, wrap_u=2, wrap_v=1, filter=0), uv(centres([-1, -2, 4, 5, 8])))
    assert mirror[:, 0].tolist() == [0, 1, 3, 2, 0]
    border = orc.texture_fetch(dict(data=row, wrap_u=3, wrap_v=1, filter=0), uv(centres([-1, 0, 3, 4])))
    assert border[:, 0].tolist() == [0, 0, 3, 0]
    # linear + wrap blends the last and the first texel across the seam
    seam = orc.texture_fetch(dict(data=row, wrap_u=0, wrap_v=1, filter=1), uv([1.0]))
    assert seam[0, 0] == 1.5


def test_alpha_cutouts_open_the_screen():
    scene = scenes.textured_example(width=96, height=54)
    tex = orc.OracleScene(scene).trace_primary(0)
    plain = scenes.textured_example(width=96, height=54)
    plain.textures = []
    for m in plain.materials:
        for k in ("albedo_tex", "luminance_tex", "roughness_tex", "metallic_tex", "normal_tex"):
            m[k] = 0xFFFF
    ref = orc.OracleScene(plain).trace_primary(0)
    on_screen = ref["instance"] == 1
    through = on_screen & (tex["instance"] != 1)
    # roughly a third of the 4x4-texel blocks have alpha 0
    assert 0.15 < through.sum() / on_screen.sum() < 0.55
    # pixels that do not look at the screen are untouched
    assert np.array_equal(tex["instance"][~on_screen], ref["instance"][~on_screen])
    assert np.array_equal(tex["t"][~on_screen].view(np.uint32), ref["t"][~on_screen].view(np.uint32))
    # brute force agrees with the BVH path on the textured scene
    osc = orc.OracleScene(scene)
    assert osc is not None


def test_textured_render_differs_and_is_finite():
    scene = scenes.textured_example(width=48, height=27, max_ray_depth=2)
    from luminary_b200 import api

    lt = api.build_light_tree(scene)
    osc = orc.OracleScene(scene)
    osc.set_light_tree(*lt)
    img, info = osc.render(0, 2)
    assert np.isfinite(img).all() and img[:3].mean() > 0
    assert info["shadow_rays"] > 0


def test_fetch_reproduces_b200_texture_unit_golden():
    """tests/golden/tex_probe_b200.npz holds tex2D<float4> results measured on a B200 by tools/tex_probe.py (the script
    that made the fixture): fine sweeps between two texels of fp32 rows {0, 1, 0, 1, ...} of width 2 / 64 / 4096, random
    (u, v) on a 2x2 fp32 texture {0, 1; 2, 4}, and sweeps on the u8 rows {0, 255} and {51, 102}. The restatement must
    reproduce every value bit for bit."""
    import os

    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "tex_probe_b200.npz"))
    half = lambda u: np.stack([u, np.full_like(u, 0.5)], axis=1)
    for w in (2, 64, 4096):
        row = np.zeros((1, w, 1), np.float32)
        row[0, 1::2, 0] = 1.0
        got = orc.texture_fetch(dict(data=row, wrap_u=1, wrap_v=1, filter=1), half(d[f"u_{w}"]))
        assert np.array_equal(got[:, 0], d[f"v_{w}"]), w
    t = dict(data=np.array([[[0.0], [1.0]], [[2.0], [4.0]]], np.float32), wrap_u=1, wrap_v=1, filter=1)
    assert np.array_equal(orc.texture_fetch(t, d["uv2d"])[:, 0], d["v2d"])
    for name, row in (("v_u8a", [0, 255]), ("v_u8b", [51, 102])):
        t = dict(data=np.array([[[row[0]], [row[1]]]], np.uint8), wrap_u=1, wrap_v=1, filter=1)
        assert np.array_equal(orc.texture_fetch(t, half(d["u_u8"]))[:, 0], d[name]), name


def test_mip_chain_restatement():
    """orc_texture_next_mip (device_texture.c:128-245, cuda/mipmap.cuh): for power-of-two extents every level is the 2x2 box
    filter of the one above, rounded half up, and a non-zero alpha never becomes zero."""
    rng = np.random.default_rng(1)
    data = (rng.random((16, 32, 4)) * 255).astype(np.uint8)
    data[4:8, 4:8, 3] = 0
    data[0:2, 0:2, 3] = [[1, 0], [0, 0]]  # mean 0.25 -> rounds to 0 without the opacity rule
    levels = orc.texture_mip_chain(dict(data=data, wrap_u=0, wrap_v=0, filter=1))
    assert [l.shape for l in levels] == [(16, 32, 4), (8, 16, 4), (4, 8, 4), (2, 4, 4)]  # floor(log2(min(w, h))) levels
    a = levels[0].astype(np.float64)
    box = (a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2]) / 4.0
    assert np.array_equal(levels[1][..., :3], np.floor(box[..., :3] + 0.5).astype(np.uint8))
    assert levels[1][0, 0, 3] == 1 and levels[1][2, 2, 3] == 0 and levels[1][3, 3, 3] == 0
    fp = rng.random((8, 8, 4)).astype(np.float32)
    lf = orc.texture_mip_chain(dict(data=fp, wrap_u=1, wrap_v=1, filter=1))
    assert len(lf) == 3 and np.allclose(lf[1], (fp[0::2, 0::2] + fp[1::2, 0::2] + fp[0::2, 1::2] + fp[1::2, 1::2]) / 4.0, atol=1e-6)


def test_textured_hits_match_golden_fixture():
    """tests/golden/textured_adaptive.json (generator: tests/golden/make_textured_adaptive.py): closest hits with alpha cut-outs."""
    import hashlib
    import json
    import os

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "textured_adaptive.json")))["textured_hits"]
    ref = orc.OracleScene(scenes.textured_example(width=384, height=216)).trace_primary(3)
    d = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert ref["tri"].size == g["count"] and int((ref["instance"] == 1).sum()) == g["screen_hits"]
    assert d(ref["instance"]) == g["sha256"]["instance"] and d(ref["tri"]) == g["sha256"]["tri"] and d(ref["t"].view(np.uint32)) == g["sha256"]["t"]
