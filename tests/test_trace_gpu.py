"""GPU parity of the closest-hit path through the C ABI.

Bar (BASELINE.json north_star): closest-hit triangle ids bit-exact against the oracle except documented exact-t
ties, hit t and barycentrics within 1e-5 relative. Because the CUDA geometry kernels evaluate the oracle's
watertight test operation for operation (no FMA contraction, IEEE div/sqrt) and equal-t ties resolve to the lower
primitive index in both, the expectation here is stricter: ids AND t/u/v bit-identical, zero exceptions.
"""
import numpy as np
import pytest

import orc
from luminary_b200 import scenes

pytestmark = pytest.mark.gpu

SKY = 0xFFFFFFFE


def _device_for(scene):
    from luminary_b200 import api

    dev = api.Device(0)
    dev.load_scene(scene)
    return dev


def _compare_primary(scene, sample_id=0):
    dev = _device_for(scene)
    inst, tri, t, u, v = dev.trace_primary(sample_id)
    ref = orc.OracleScene(scene).trace_primary(sample_id)
    stats = dev.stats()
    dev.destroy()
    return (inst, tri, t, u, v), ref, stats


def test_config1_example_primary_ids_bit_exact():
    scene = scenes.example()  # 960 x 540, 40 972 triangles
    (inst, tri, t, u, v), ref, stats = _compare_primary(scene)
    assert stats["bvh_tris"] == 40972
    n = inst.size
    assert n == 960 * 540
    id_mismatch = np.count_nonzero((inst != ref["instance"]) | (tri != ref["tri"]))
    assert id_mismatch == 0, f"{id_mismatch} of {n} closest-hit ids differ from the oracle"
    hit = ref["instance"] != SKY
    assert hit.all()  # closed room
    # north-star tolerance ...
    rel = np.abs(t[hit] - ref["t"][hit]) / np.maximum(np.abs(ref["t"][hit]), 1e-30)
    assert rel.max() <= 1e-5
    assert np.abs(u[hit] - ref["u"][hit]).max() <= 1e-5 and np.abs(v[hit] - ref["v"][hit]).max() <= 1e-5
    # ... and the stricter expectation of identical arithmetic
    assert np.array_equal(t.view(np.uint32), ref["t"].view(np.uint32))
    assert np.array_equal(u.view(np.uint32), ref["u"].view(np.uint32))
    assert np.array_equal(v.view(np.uint32), ref["v"].view(np.uint32))


def test_watertight_vs_moller_trumbore_t_within_tolerance():
    """The reference re-intersects with its own Moeller-Trumbore (cuda/math.cuh:1337-1358); our t must agree with it to 1e-5."""
    import ctypes as C

    scene = scenes.example(width=192, height=108, sphere_subdiv=3)
    dev = _device_for(scene)
    inst, tri, t, u, v = dev.trace_primary(0)
    dev.destroy()
    osc = orc.OracleScene(scene)
    o, d = osc.camera_rays(0)
    world = osc.world_tris()
    ref = osc.trace_rays(o, d)
    L = orc.lib()
    worst = 0.0
    for i in range(0, o.shape[0], 7):
        p = int(ref["prim"][i])
        uu, vv = C.c_float(), C.c_float()
        tm = L.orc_tri_mt(orc.fptr(np.ascontiguousarray(world[p].reshape(-1))), orc.vec3(o[i]), orc.vec3(d[i]), C.byref(uu), C.byref(vv))
        if tm > 3e38:  # MT rejects an edge-grazing hit the watertight test accepts: documented, must be at the rim
            assert min(u[i], v[i], 1 - u[i] - v[i]) < 1e-4
            continue
        worst = max(worst, abs(tm - t[i]) / abs(tm))
    assert worst <= 1e-5


def test_instanced_scene_with_rotation_and_scale():
    scene = scenes.atrium(target_tris=60000, width=320, height=180)
    (inst, tri, t, u, v), ref, stats = _compare_primary(scene, sample_id=3)
    assert stats["bvh_tris"] == 60000
    assert np.array_equal(inst, ref["instance"]) and np.array_equal(tri, ref["tri"])
    assert np.array_equal(t.view(np.uint32), ref["t"].view(np.uint32))
    assert len(np.unique(inst)) > 5  # columns are separate instances


def test_random_rays_including_misses_and_axis_aligned():
    scene = scenes.example(width=64, height=36, sphere_subdiv=3)
    # open the room: drop the ceiling and one wall so that rays can miss
    room = scene.meshes[0]
    keep = np.ones(room.num_tris, bool)
    keep[2:4] = False
    keep[10:12] = False
    scene.meshes[0] = scenes.Mesh(room.vertex[keep], room.normal[keep], room.uv[keep], room.material[keep])
    rng = np.random.default_rng(5)
    n = 20000
    o = rng.uniform([-1.9, 0.1, -3.9], [1.9, 2.9, -0.1], size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:300] = 0.0
    d[:100, 0] = 1.0
    d[100:200, 1] = -1.0
    d[200:300, 2] = -1.0
    dev = _device_for(scene)
    inst, tri, t, u, v = dev.trace_rays(o, d)
    dev.destroy()
    osc = orc.OracleScene(scene)
    ref = osc.trace_rays(o, d)
    miss = ref["prim"] == SKY
    assert 0 < miss.sum() < n
    assert np.array_equal(inst == SKY, miss)
    # map oracle flattened prim -> (instance, tri)
    offs = np.cumsum([0] + [scene.meshes[i.mesh_id].num_tris for i in scene.instances])
    ri = np.searchsorted(offs, ref["prim"][~miss], side="right") - 1
    rt = ref["prim"][~miss] - offs[ri]
    assert np.array_equal(inst[~miss], ri.astype(np.uint32)) and np.array_equal(tri[~miss], rt.astype(np.uint32))
    assert np.array_equal(t[~miss].view(np.uint32), ref["t"][~miss].view(np.uint32))
    assert (t[miss] > 3e38).all()


def test_empty_scene_and_single_triangle():
    from luminary_b200 import api

    sc = scenes.example(width=32, height=18, sphere_subdiv=1)
    dev = api.Device(0)
    dev.update_instances([])
    dev.update_materials(sc.materials)
    dev.update_settings(32, 18, 0)
    dev.update_camera(sc.camera)
    dev.build_accel()
    inst, tri, t, u, v = dev.trace_primary(0)
    assert (inst == SKY).all() and (t > 3e38).all()
    dev.destroy()

    one = scenes.mesh_from_tris(np.array([[[-5, -5, -3], [5, -5, -3], [0, 5, -3]]], np.float32), 0)
    dev = api.Device(0)
    dev.add_mesh(one.vertex, one.normal, one.uv, one.material)
    dev.update_instances([scenes.Instance(0)])
    dev.update_materials(sc.materials)
    dev.update_settings(32, 18, 0)
    dev.update_camera(scenes.default_camera())
    dev.build_accel()
    inst, tri, t, u, v = dev.trace_primary(0)
    assert (inst == 0).sum() > 100 and ((inst == 0) | (inst == SKY)).all()
    assert np.allclose(t[inst == 0] * 1.0, t[inst == 0])
    dev.destroy()


def test_api_errors_match_reference_codes():
    from luminary_b200 import api

    dev = api.Device(0)
    with pytest.raises(api.LuminaryError) as e:
        dev.trace_primary(0)  # no settings yet
    assert e.value.code == 7
    with pytest.raises(api.LuminaryError) as e:
        dev.update_settings(0, 10, 1)
    assert e.value.code == 3
    with pytest.raises(api.LuminaryError) as e:
        dev.update_instances([scenes.Instance(3)])  # mesh does not exist
    assert e.value.code == 3
    with pytest.raises(api.LuminaryError) as e:
        api.Device(99)
    assert e.value.code == 13
    dev.destroy()


def test_config1_matches_golden_fixture():
    """The CUDA path against the committed fixture (tests/golden/config1_hits.json), independent of the oracle library
    being present: digests of ids and of the bit patterns of t, u, v."""
    import hashlib
    import json
    import os

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "config1_hits.json")))
    dev = _device_for(scenes.example())
    inst, tri, t, u, v = dev.trace_primary(0)
    dev.destroy()
    dig = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert inst.size == g["count"]
    assert dig(inst) == g["sha256"]["instance"] and dig(tri) == g["sha256"]["tri"]
    assert dig(t.view(np.uint32)) == g["sha256"]["t"] and dig(u.view(np.uint32)) == g["sha256"]["u"] and dig(v.view(np.uint32)) == g["sha256"]["v"]


def test_hierarchy_builders_agree(monkeypatch):
    """Closest hits do not depend on the acceleration structure: the PLOC hierarchy (default) and the Karras radix tree
    (LUMB200_BVH_BUILDER=lbvh) must return identical ids and bit-identical t / u / v; so do different search radii."""
    scene = scenes.atrium(target_tris=80000, width=320, height=180)
    results = []
    for env in ({}, {"LUMB200_BVH_BUILDER": "lbvh"}, {"LUMB200_PLOC_RADIUS": "4"}):
        for k in ("LUMB200_BVH_BUILDER", "LUMB200_PLOC_RADIUS"):
            monkeypatch.delenv(k, raising=False)
        for k, val in env.items():
            monkeypatch.setenv(k, val)
        dev = _device_for(scene)
        results.append(dev.trace_primary(3) + (dev.stats()["bvh_nodes"],))
        dev.destroy()
    base = results[0]
    assert len({r[5] for r in results}) > 1  # the structures really differ
    for r in results[1:]:
        assert np.array_equal(r[0], base[0]) and np.array_equal(r[1], base[1])
        for k in (2, 3, 4):
            assert np.array_equal(r[k].view(np.uint32), base[k].view(np.uint32))


def test_reinsertion_keeps_hits_and_lowers_sah(monkeypatch):
    """Parallel reinsertion (bvh_build.cu section 4c, on by default) restructures the binary hierarchy under the collapse: every
    primitive must stay reachable (ids and t / u / v bit-identical to the un-optimised tree's, which the other tests pin to the oracle),
    the SAH cost of the collapsed tree must drop, and a far larger pass budget than the default must leave a valid tree too."""
    scene = scenes.atrium(target_tris=120000, width=320, height=180)
    results = {}
    for passes in ("0", "4", "200"):
        monkeypatch.setenv("LUMB200_BVH_REINSERT", passes)
        monkeypatch.setenv("LUMB200_BVH_REINSERT_MIN_GAIN", "0")  # spend the whole budget (stops only when a pass moves nothing)
        dev = _device_for(scene)
        st = dev.stats()
        results[passes] = dev.trace_primary(1) + (st["bvh_sah_cost"], st["bvh_tris"], st["stack_overflows"])
        dev.destroy()
    base = results["0"]
    for passes in ("4", "200"):
        r = results[passes]
        assert r[6] == base[6] and r[7] == 0
        assert np.array_equal(r[0], base[0]) and np.array_equal(r[1], base[1])
        for k in (2, 3, 4):
            assert np.array_equal(r[k].view(np.uint32), base[k].view(np.uint32))
    assert results["4"][5] < 0.98 * base[5], (results["4"][5], base[5])
    assert results["200"][5] < results["4"][5]


def test_full_size_atrium_sampled_rays_and_shadow_consistency():
    """BASELINE config 2 geometry at full size (1M triangles, 1920x1080 primary rays): the oracle re-traces a random
    sample of 20 000 of the rays (ids and t bit-exact), and a size-independent property is checked on ALL rays: the
    reported hit point lies on the reported triangle (|barycentric reconstruction - (o + t d)| small)."""
    scene = scenes.atrium(1_000_000, 1920, 1080, 5)
    dev = _device_for(scene)
    inst, tri, t, u, v = dev.trace_primary(0)
    st = dev.stats()
    dev.destroy()
    assert st["bvh_tris"] == 1_000_000
    osc = orc.OracleScene(scene)
    rng = np.random.default_rng(11)
    idx = np.sort(rng.choice(inst.size, 20000, replace=False))
    import ctypes as C

    L = orc.lib()
    w = scene.width
    o = np.empty((idx.size, 3), np.float32)
    d = np.empty((idx.size, 3), np.float32)
    oo, dd = orc.Vec3(), orc.Vec3()
    for k, i in enumerate(idx):
        L.orc_camera_sample(C.byref(osc.camera), C.byref(osc.settings), L.orc_path_id_get(int(i % w), int(i // w), 0), C.byref(oo), C.byref(dd))
        o[k] = (oo.x, oo.y, oo.z)
        d[k] = (dd.x, dd.y, dd.z)
    ref = osc.trace_rays(o, d)
    # flattened primitive index of the device's (instance, tri) handles
    offs = np.cumsum([0] + [scene.meshes[i.mesh_id].num_tris for i in scene.instances if i.active])
    hit = inst[idx] != SKY
    flat = np.where(hit, offs[np.minimum(inst[idx], len(offs) - 2)] + tri[idx], SKY).astype(np.uint32)
    assert np.array_equal(flat, ref["prim"])
    assert np.array_equal(t[idx].view(np.uint32)[hit], ref["t"].view(np.uint32)[hit])
    # property on the sampled rays' geometry: o + t d == v0 + u e1 + v e2 within fp32 slack
    wt = osc.world_tris()[ref["prim"][hit]]
    p_ray = o[hit] + t[idx][hit, None] * d[hit]
    p_tri = wt[:, 0] + u[idx][hit, None] * (wt[:, 1] - wt[:, 0]) + v[idx][hit, None] * (wt[:, 2] - wt[:, 0])
    assert np.abs(p_ray - p_tri).max() <= 2e-4 * max(1.0, np.abs(p_ray).max())
    assert hit.mean() > 0.99  # closed hall


@pytest.mark.parametrize("shape,blades", [(0, 0), (1, 6)])
def test_thin_lens_aperture_primary_hits(shape, blades):
    """Thin-lens camera with an open aperture (camera_thin_lens.cuh:8-86: round disc, and the bladed polygon with its
    LENS_BLADE random number): the product's primary hits against the oracle's, and the oracle's rays against the REFERENCE's
    own tasks_create kernel where oracle/_ref/librefdev.so exists. Lens sampling goes through sinf / cosf / sqrtf (IEEE in the
    product's geometry TU, libm in the oracle, fast math in the reference): origins may differ in the last bits, so ids are
    required to agree on all but rim pixels and t to 1e-4 where they do."""
    import refdev
    from luminary_b200 import api

    scene = scenes.example(width=192, height=108, sphere_subdiv=3)
    scene.camera = dict(scene.camera, aperture_size=0.04, object_distance=2.5, aperture_shape=shape, aperture_blade_count=max(blades, 3))
    for sample_id in (0, 7):
        (inst, tri, t, u, v), ref, stats = _compare_primary(scene, sample_id)
        same = (inst == ref["instance"]) & (tri == ref["tri"])
        rel = np.abs(t[same] - ref["t"][same]) / np.maximum(np.abs(ref["t"][same]), 1e-30)
        bit = (t.view(np.uint32) == ref["t"].view(np.uint32)).mean()
        print(f"  aperture shape {shape} sample {sample_id}: ids equal {same.mean():.5f}, t rel max {rel.max():.2e}, t bit-identical {bit:.4f}")
        assert same.mean() >= 0.999 and rel.max() <= 1e-4
    # the blur is real: with the aperture closed other triangles are hit
    closed = dict(scene.camera, aperture_size=0.0)
    sharp = scenes.example(width=192, height=108, sphere_subdiv=3)
    sharp.camera = closed
    dev = _device_for(sharp)
    inst0, tri0, *_ = dev.trace_primary(7)
    dev.destroy()
    assert ((inst0 != inst) | (tri0 != tri)).mean() > 0.02
    if refdev.available():
        scene.width, scene.height = 192, 108
        rd = refdev.RefDevice(scene, light_tree=None)
        T = 2 * refdev.THREADS_PER_BLOCK * 4
        K = -(-scene.width * scene.height // T)
        rd.configure(T // refdev.THREADS_PER_BLOCK, K)
        rd.set_state(0, 7)
        rd.launch("tasks_create")
        tasks = rd.task_states()[refdev.PRESORT]
        counts = rd.download("trace_counts", np.uint16)[:T]   # the harness' buffers only ever grow
        slot, thread = np.meshgrid(np.arange(K), np.arange(T), indexing="ij")
        live = slot < counts[None, :]
        tt = tasks[live]
        pixel = (thread + slot * T)[live]
        o, d = orc.OracleScene(scene).camera_rays(7)
        eo, ed = np.abs(tt["origin"] - o[pixel]).max(), np.abs(tt["ray"] - d[pixel]).max()
        print(f"  aperture shape {shape}: reference tasks_create vs oracle max |origin| {eo:.3e} |ray| {ed:.3e}")
        assert eo <= 1e-5 and ed <= 1e-5
        assert np.abs(tt["origin"] - np.asarray(scene.camera["pos"], np.float32)).max() > 1e-3   # the lens was sampled
