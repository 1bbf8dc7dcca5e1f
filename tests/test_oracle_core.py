"""CPU tests of the oracle's integer paths (PathID, RNG, packing) and its geometry (BVH2 vs brute force).

The reference ships no tests or golden vectors, so the known answers here come from a second,
independent restatement of the published algorithms in pure Python (tests/golden/make_rng_kat.py -> rng_kat.json):
two implementations written separately from the same source must agree bit for bit.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import orc
from luminary_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_path_id_roundtrip():
    L = orc.lib()
    for x, y, s in [(0, 0, 0), (1919, 1079, 63), (16383, 16383, (1 << 20) - 1), (123, 456, 70000), (5, 7, 0x3FFFF), (9, 1, 0xC0001)]:
        pid = L.orc_path_id_get(x, y, s)
        px, py = C.c_uint32(), C.c_uint32()
        L.orc_path_id_pixel(pid, C.byref(px), C.byref(py))
        assert (px.value, py.value) == (x, y)
        assert L.orc_path_id_sample(pid) == s


def test_rng_known_answers():
    L = orc.lib()
    kat = json.load(open(os.path.join(GOLDEN, "rng_kat.json")))
    for key, counter, want in kat["squares32"]:
        assert L.orc_squares32(key, counter) == want
    for key, counter, want in kat["squares16"]:
        assert L.orc_squares16(key, counter) == want
    for offset, dim, wx, wy in kat["sobol"]:
        r = L.orc_sobol(offset, dim)
        assert (r.x, r.y) == (wx, wy)
    for target, px, py, seq, depth, wx, wy in kat["random_2d_base"]:
        r = L.orc_random_2d_base(target, px, py, seq, depth)
        assert (r.x, r.y) == (wx, wy)


def test_rng_float_range_and_low_bit():
    L = orc.lib()
    for i in range(200):
        r = L.orc_random_2d_base(orc_target := 39 + (i % 3), i * 7, i * 13, i, i % 6)
        f = L.orc_u32_to_float(r.x)
        assert 0.0 <= f < 1.0
    assert L.orc_u32_to_float(0xFFFFFFFF) < 1.0
    assert L.orc_u32_to_float(0) == 0.0


def test_sobol_is_stratified():
    """(0,2)-sequence property survives Owen scrambling: any 16 consecutive-power-of-two block of a dimension pair
    puts exactly one point in each of the 16 elementary intervals 1/16 x 1 and 1 x 1/16."""
    L = orc.lib()
    for dim in (0, 39, 400, 577 * 3 + 51):
        pts = [L.orc_sobol(i, dim) for i in range(16)]
        xs = sorted(p.x >> 28 for p in pts)
        ys = sorted(p.y >> 28 for p in pts)
        assert xs == list(range(16))
        assert ys == list(range(16))


def test_packing_roundtrips():
    L = orc.lib()
    rng = np.random.default_rng(1)
    for _ in range(200):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        for pack in (L.orc_pack_normal, L.orc_pack_normal_host):
            u = L.orc_unpack_normal(pack(orc.vec3(n)))
            assert np.allclose([u.x, u.y, u.z], n, atol=1e-4)
        r = L.orc_ray_unpack(L.orc_ray_pack(orc.vec3(n)))
        assert np.allclose([r.x, r.y, r.z], n, atol=1e-6)
    # record: 21-bit floats, truncation => relative error < 2^-12, never larger than the input
    for c in [(1.0, 1.0, 1.0), (0.123, 4.5, 1e-3), (0.0, 0.0, 0.0), (1000.0, 0.5, 7.25)]:
        p = L.orc_record_pack(orc.RGB(*c))
        u = L.orc_record_unpack(p)
        for a, b in zip((u.r, u.g, u.b), c):
            assert a <= b and (b == 0 or (b - a) / b < 2.0 ** -12)
    p = L.orc_record_pack(orc.RGB(1.0, 1.0, 1.0))
    assert (p.x, p.y) == (0x7F000 | ((0x7F000 << 21) & 0xFFFFFFFF), (0x7F000 >> 11) | (0x7F000 << 10))
    # uv: bfloat16 truncation of both components
    uv = L.orc_unpack_uv(L.orc_pack_uv(0.337, 7.75))
    assert abs(uv.x - 0.337) < 0.337 * 2.0 ** -7 and uv.y == 7.75
    assert L.orc_ior_compress(1.0) == 0
    assert abs(L.orc_ior_decompress(L.orc_ior_compress(1.5)) - 1.5) < 2.0 ** -7


def test_material_pack_layout():
    m = scenes.default_material(albedo=(0.25, 0.5, 0.75, 1.0), roughness=0.3, metallic=True, emission=(3.0, 1.5, 0.0), emission_active=True,
                                refraction_index=1.5, base_substrate=1)
    p = orc.pack_material(m)
    assert C.sizeof(orc.MaterialPacked) == 32
    assert p.flags == (0x01 | 0x02 | 0x08 | 0x40)
    assert p.albedo_a == 0xFFFF and p.albedo_g == int(0.5 * 65535 + 0.5)
    assert p.roughness == int(np.float32(0.3) * np.float32(65535.0) + np.float32(0.5))
    assert p.refraction_index == int(0.25 * 65535 + 0.5)
    # emission is normalised by max + 1 and the scale stored as an unsigned 8.8 float
    assert p.emission_r == int(np.float32(3.0) * np.float32(0.25) * 65535 + 0.5)
    assert p.emission_scale == (np.float32(4.0).view(np.uint32) >> 15) & 0xFFFF


def test_transform_inverse():
    L = orc.lib()
    t = orc.Transform()
    t.translation = orc.vec3((1.0, -2.0, 3.0))
    t.scale = orc.vec3((1.5, 0.5, 2.0))
    t.rotation = L.orc_quat_pack(L.orc_euler_to_quat(orc.vec3((0.3, -1.1, 2.0))))
    v = orc.vec3((0.7, 0.1, -0.4))
    w = L.orc_transform_apply(C.byref(t), v)
    b = L.orc_transform_apply_inv(C.byref(t), w)
    assert np.allclose([b.x, b.y, b.z], [0.7, 0.1, -0.4], atol=2e-4)  # quaternion16 is only ~1e-4 accurate


def test_triangle_tests_agree_inside():
    L = orc.lib()
    rng = np.random.default_rng(7)
    tri = np.array([[0, 0, -3], [1, 0, -3.5], [0, 1, -2.5]], np.float32).reshape(-1)
    hits = 0
    for _ in range(500):
        d = np.array([rng.uniform(-0.3, 1.1) - 0.1, rng.uniform(-0.3, 1.1) - 0.1, -3.0])
        d /= np.linalg.norm(d)
        u1, v1 = C.c_float(), C.c_float()
        t1 = L.orc_tri_mt(orc.fptr(tri), orc.vec3((0.1, 0.1, 0)), orc.vec3(d), C.byref(u1), C.byref(v1))
        t2, u2, v2 = C.c_float(), C.c_float(), C.c_float()
        ok = L.orc_tri_watertight(orc.fptr(tri), orc.vec3((0.1, 0.1, 0)), orc.vec3(d), C.byref(t2), C.byref(u2), C.byref(v2))
        inside = t1 < 3e38
        margin = min(u1.value, v1.value, 1 - u1.value - v1.value)
        if abs(margin) > 1e-5:
            assert inside == bool(ok)
        if inside and ok:
            hits += 1
            assert abs(t1 - t2.value) <= 1e-5 * abs(t1)
            assert abs(u1.value - u2.value) < 1e-5 and abs(v1.value - v2.value) < 1e-5
    assert hits > 50


def test_bvh_matches_bruteforce_small_scene():
    """The BVH2 must return exactly the brute-force closest hit (same watertight arithmetic, ties to the lower index)."""
    L = orc.lib()
    sc = scenes.example(width=64, height=36, sphere_subdiv=2)
    osc = orc.OracleScene(sc)
    o, d = osc.camera_rays(0)
    res = osc.trace_rays(o, d)
    mism = 0
    for i in range(0, o.shape[0], 3):
        h = L.orc_closest_hit_bruteforce(osc.handle, orc.vec3(o[i]), orc.vec3(d[i]), 0.0, 3.4e38, 0xFFFFFFFF, 0)
        if h.prim != res["prim"][i] or h.t != res["t"][i]:
            mism += 1
    assert mism == 0
    assert (res["prim"] != 0xFFFFFFFE).all()  # closed room: every primary ray hits


def test_scene_generators_are_deterministic_and_sized():
    a = scenes.example()
    b = scenes.example()
    assert a.num_tris == 40972
    assert all(np.array_equal(x.vertex, y.vertex) for x, y in zip(a.meshes, b.meshes))
    at = scenes.atrium(target_tris=20000, width=64, height=36)
    assert at.num_tris == 20000
    assert len(at.materials) == 17


def test_config1_hits_match_golden_fixture():
    """BASELINE config 1 on the CPU: the oracle must reproduce tests/golden/config1_hits.json (digests of ids, t, u, v of
    the 518 400 primary rays of the Example scene; generator: tests/golden/make_config1_hits.py)."""
    import hashlib

    g = json.load(open(os.path.join(GOLDEN, "config1_hits.json")))
    ref = orc.OracleScene(scenes.example()).trace_primary(0)
    assert ref["tri"].size == g["count"]
    dig = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert dig(ref["instance"]) == g["sha256"]["instance"]
    assert dig(ref["tri"]) == g["sha256"]["tri"]
    for k in ("t", "u", "v"):
        assert dig(ref[k].view(np.uint32)) == g["sha256"][k]
    for i, inst, tri, tbits in g["spots"]:
        assert (int(ref["instance"][i]), int(ref["tri"][i]), int(ref["t"][i].view(np.uint32))) == (inst, tri, tbits)
