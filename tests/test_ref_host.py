"""Pins the host-side restatements against the REFERENCE's own C code, compiled from /root/reference into
oracle/_ref/libref_host.so (oracle/ref/Makefile; device_structs.c, device_packing.c, device_light.c, host_math.c ...).

  * the product's light-tree builder (luminary_b200/csrc/host/light_tree.c) must produce byte-identical root / node blobs
    and the same light-id -> triangle map as the reference's light_tree_build (device_light.c:2236-2268);
  * the oracle's packers (oracle/orc_core.c) must equal device_struct_material_convert, device_pack_normal / _uv,
    device_struct_instance_transform_convert bit for bit.
Skipped when the reference library is not built (it can only be built where /root/reference exists)."""
import ctypes as C

import numpy as np
import pytest

import orc
import refhost
from luminary_b200 import api, scenes

pytestmark = pytest.mark.skipif(not refhost.available(), reason="oracle/_ref/libref_host.so not built (needs /root/reference)")


def _compare_tree(scene):
    mine = api.build_light_tree(scene)
    ref = refhost.build_light_tree(scene)
    assert (mine is None) == (ref is None)
    if mine is None:
        return 0
    assert mine[0] == ref[0], "root header / sections differ from light_tree_build"
    assert mine[1] == ref[1], "8-wide nodes differ from light_tree_build"
    assert np.array_equal(mine[2], ref[2]), "light id -> (instance, triangle) map differs"
    return ref[2].shape[0]


@pytest.mark.parametrize("name,make", [
    ("example", lambda: scenes.example_with_light(96, 54, 2, 3)),
    ("atrium-100k (48 emitters, root only)", lambda: scenes.atrium(100_000, 192, 108, 5)),
    ("divergence-30k (rotated instances, 4 emissive materials)", lambda: scenes.divergence(30_000, 192, 108, 8)),
    ("divergence-200k (3374 emitters, 939 nodes)", lambda: scenes.divergence(200_000, 192, 108, 8)),
    ("terrain (10k emitters)", lambda: scenes.terrain(300, 5000, 192, 108, 5)),
])
def test_light_tree_matches_reference_build(name, make):
    assert _compare_tree(make()) > 0


def test_light_tree_without_emitters():
    sc = scenes.example(64, 36, 1)
    assert api.build_light_tree(sc) is None
    assert refhost.build_light_tree(sc) is None


def test_emitter_vertices_are_world_space_triangles():
    """bvh_vertex_buffer_data (device_light.c:1226-1250) = world-space emitter triangles in light-id order; the device builds its
    emitter BVH from the same triangles (device_api.cu: ensure_light_records)."""
    sc = scenes.divergence(30_000, 192, 108, 8)
    root, nodes, handles, verts = refhost.build_light_tree(sc)
    osc = orc.OracleScene(sc)
    tris = osc.world_tris().reshape(-1, 3, 3)
    # flattened prim index of (instance, tri)
    n = osc.num_prims()
    inst = np.zeros(n, np.uint32)
    tri = np.zeros(n, np.uint32)
    for p in range(n):
        a, b = C.c_uint32(), C.c_uint32()
        orc.lib().orc_scene_prim_handle(osc.handle, p, C.byref(a), C.byref(b))
        inst[p], tri[p] = a.value, b.value
    lookup = {(int(i), int(t)): p for p, (i, t) in enumerate(zip(inst, tri))}
    for lid in range(handles.shape[0]):
        p = lookup[(int(handles[lid, 0]), int(handles[lid, 1]))]
        # the light tree rotates with the float quaternion, the device transforms (and the flattened world triangles) with the
        # 16-bit quaternion of DeviceTransform (device_structs.c:388-399): same triangle, <= 1e-3 m apart on rotated instances
        np.testing.assert_allclose(verts[lid, :, :3], tris[p], rtol=0, atol=1e-3)


def test_material_packing_matches_device_struct_material_convert():
    rng = np.random.default_rng(7)
    for k in range(200):
        m = dict(base_substrate=int(rng.integers(0, 2)), albedo=tuple(rng.random(4).astype(np.float32)),
                 emission=tuple((rng.random(3) * (20.0 if k % 3 else 0.0)).astype(np.float32)), emission_scale=float(rng.random() * 4),
                 roughness=float(rng.random()), roughness_clamp=float(rng.random()), refraction_index=float(1.0 + 2.0 * rng.random()),
                 emission_active=bool(k % 3), thin_walled=bool(rng.integers(0, 2)), metallic=bool(rng.integers(0, 2)),
                 colored_transparency=bool(rng.integers(0, 2)), roughness_as_smoothness=bool(rng.integers(0, 2)),
                 normal_map_is_compressed=bool(rng.integers(0, 2)), bidirectional_emission=bool(rng.integers(0, 2)))
        if k % 2:  # texture ids travel unchanged (device_structs.c:322-326)
            for key in ("albedo_tex", "luminance_tex", "roughness_tex", "metallic_tex", "normal_tex"):
                m[key] = int(rng.integers(0, 0x10000))
        mine = bytes(orc.pack_material(m))
        ref = refhost.material_convert(m)
        assert mine == ref, (k, m, mine.hex(), ref.hex())


def test_normal_and_uv_packing_match_device_packing():
    rng = np.random.default_rng(11)
    v = rng.normal(size=(4000, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v[:6] = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float32)
    for x, y, z in v:
        assert orc.lib().orc_pack_normal_host(orc.vec3((x, y, z))) == refhost.lib().refhost_pack_normal(float(x), float(y), float(z))
    uv = (rng.random((4000, 2)) * 6.0 - 3.0).astype(np.float32)
    for u, w in uv:
        assert orc.lib().orc_pack_uv(float(u), float(w)) == refhost.lib().refhost_pack_uv(float(u), float(w))


def test_instance_transform_matches_reference():
    rng = np.random.default_rng(5)
    L = orc.lib()
    for k in range(300):
        t = (rng.random(3) * 20 - 10).astype(np.float32)
        r = (rng.random(3) * 6.28 - 3.14).astype(np.float32) if k else np.zeros(3, np.float32)
        s = (rng.random(3) * 3 + 0.1).astype(np.float32)
        ref = refhost.instance_transform_convert(t, r, s)
        tr = orc.Transform()
        tr.translation = orc.vec3(t)
        tr.scale = orc.vec3(s)
        tr.rotation = L.orc_quat_pack(L.orc_euler_to_quat(orc.vec3(r)))
        assert bytes(tr) == ref, (k, bytes(tr).hex(), ref.hex())


def test_mesh_vertex_and_texture_triangle_records():
    """device_struct_vertex_convert / device_struct_triangle_texture_convert (device_structs.c:351-374) against the records
    the oracle scene holds (same layout the device uploads: 3 x {pos, packed normal}, {3 packed uv, material})."""
    sc = scenes.example_with_light(64, 36, 2, 1)
    for m in sc.meshes:
        verts, tex = refhost.mesh_convert(m)
        pos = np.asarray(m.vertex, np.float32).reshape(-1, 3)
        assert np.array_equal(verts[:, :3].view(np.float32), pos)
        nrm = np.asarray(m.normal, np.float32).reshape(-1, 3)
        mine_n = np.array([orc.lib().orc_pack_normal_host(orc.vec3(n)) for n in nrm], np.uint32)
        assert np.array_equal(verts[:, 3], mine_n)
        uv = np.asarray(m.uv, np.float32).reshape(-1, 2)
        mine_uv = np.array([orc.lib().orc_pack_uv(float(a), float(b)) for a, b in uv], np.uint32).reshape(-1, 3)
        assert np.array_equal(tex[:, :3], mine_uv)
        assert np.array_equal(tex[:, 3] & 0xFFFF, np.asarray(m.material, np.uint32).reshape(-1))


# ---------------------------------------------------------------------------------------------------------------------
# the PRODUCT's upload-path packers (csrc/device_api.cu, reached through lumb200_host_pack_*), byte for byte against the
# reference's host code: what the kernels read is what device_struct_*_convert would have produced
# ---------------------------------------------------------------------------------------------------------------------
def test_product_material_packer_matches_reference_bytes():
    from luminary_b200 import api

    rng = np.random.default_rng(23)
    for k in range(300):
        m = dict(base_substrate=int(rng.integers(0, 2)), albedo=tuple(rng.random(4).astype(np.float32)),
                 emission=tuple((rng.random(3) * (20.0 if k % 3 else 0.0)).astype(np.float32)), emission_scale=float(rng.random() * 4),
                 roughness=float(rng.random()), roughness_clamp=float(rng.random()), refraction_index=float(1.0 + 2.0 * rng.random()),
                 emission_active=bool(k % 3), thin_walled=bool(rng.integers(0, 2)), metallic=bool(rng.integers(0, 2)),
                 colored_transparency=bool(rng.integers(0, 2)), roughness_as_smoothness=bool(rng.integers(0, 2)),
                 normal_map_is_compressed=bool(rng.integers(0, 2)), bidirectional_emission=bool(rng.integers(0, 2)))
        if k % 2:
            for key in ("albedo_tex", "luminance_tex", "roughness_tex", "metallic_tex", "normal_tex"):
                m[key] = int(rng.integers(0, 0x10000))
        if k < 8:  # exact end points of every quantiser
            m["albedo"] = (0.0, 1.0, 0.5, 1.0 if k & 1 else 0.0)
            m["roughness"] = float(k & 1)
            m["refraction_index"] = 1.0 if k & 2 else 3.0
        mine, ref = api.pack_material(m), refhost.material_convert(m)
        assert mine == ref, (k, m, mine.hex(), ref.hex())


def test_product_triangle_packer_matches_reference_records():
    from luminary_b200 import api

    rng = np.random.default_rng(29)
    sc = scenes.example_with_light(64, 36, 2, 1)
    for m in sc.meshes:
        # perturb the normals / uvs so that every code path of the octahedral and bf16 packers is hit
        mm = scenes.Mesh(m.vertex, (m.normal + rng.normal(scale=0.3, size=m.normal.shape)).astype(np.float32),
                         (rng.random(m.uv.shape) * 8.0 - 4.0).astype(np.float32), m.material)
        rv, rt = refhost.mesh_convert(mm)
        pv, pt = api.pack_triangles(mm)
        assert np.array_equal(pv, np.asarray(rv).view(np.uint32).reshape(pv.shape))
        assert np.array_equal(pt[:, :3], np.asarray(rt).view(np.uint32).reshape(pt.shape)[:, :3])
        assert np.array_equal(pt[:, 3] & 0xFFFF, np.asarray(rt).view(np.uint32).reshape(pt.shape)[:, 3] & 0xFFFF)


def test_product_transform_packer_matches_reference_bytes():
    from luminary_b200 import api

    rng = np.random.default_rng(31)
    for k in range(400):
        t = (rng.random(3) * 20 - 10).astype(np.float32)
        r = (rng.random(3) * 6.28 - 3.14).astype(np.float32) if k else np.zeros(3, np.float32)
        s = (rng.random(3) * 3 + 0.1).astype(np.float32)
        mine, ref = api.pack_transform(t, r, s), refhost.instance_transform_convert(t, r, s)
        assert mine == ref, (k, mine.hex(), ref.hex())
