"""Pins the shading restatement (oracle/orc_shade.c) and the product's LUT generation against the REFERENCE's own CUDA kernels,
compiled from /root/reference for sm_100a (oracle/ref/ref_patch.sh -> oracle/_ref/librefdev.so) and launched unmodified:

  tasks_create              cuda/kernels.cuh:45-193      vs orc_camera_sample / PathID / record / medium initialisation
  geometry_process_tasks    cuda/geometry.cuh:11-180     vs orc_shade_vertices, vertex by vertex, at wavefront iterations 0, 1 and 2
  sky_process_tasks         cuda/sky.cuh:609-633         vs the constant-sky miss term
  accumulation_*            cuda/accumulation.cuh        vs the plane sums of orc_render
  bsdf_generate_*_lut       cuda/bsdf_lut.cuh:20-209     vs lumb200_device_build_bsdf_lut (product) and orc_bsdf_lut_generate

The reference is built with --use_fast_math (CMakeLists.txt:48), the oracle is IEEE + libm, so floating-point outputs are
compared with the tolerances written below and discrete decisions (lobe / light / Russian-roulette choices) must agree on all but
a stated small fraction of vertices (a random number within rounding distance of a decision threshold flips it).
The OptiX programs (closest hit, shadow evaluation) are NOT part of this: they need libnvoptix."""
import numpy as np
import pytest

import orc
import refdev
import refhost
from luminary_b200 import api, scenes

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refdev.available(), reason="oracle/_ref/librefdev.so not built (needs /root/reference)")]

W, H = 192, 108


def parity_scene():
    """Every material class of the path in one small room: diffuse walls, glossy dielectric, metal, smooth and rough glass, two emitters
    (one through a rotated, non-uniformly scaled instance), constant sky visible through an open front."""
    sc = scenes.example_with_light(W, H, sphere_subdiv=3, max_ray_depth=4)
    sc.materials += [
        scenes.default_material(base_substrate=1, albedo=(0.95, 0.97, 1.0, 1.0), roughness=0.02, refraction_index=1.5),   # 6 smooth glass
        scenes.default_material(base_substrate=1, albedo=(0.8, 0.9, 0.7, 1.0), roughness=0.35, refraction_index=1.33),    # 7 rough glass
        scenes.default_material(albedo=(0.2, 0.5, 0.9, 1.0), roughness=0.45),                                              # 8 mid-rough dielectric
        scenes.default_material(albedo=(1.0, 0.6, 0.3, 1.0), emission=(4.0, 2.0, 1.0), emission_active=True, roughness=1.0),  # 9 warm light
    ]
    sc.meshes.append(scenes.icosphere(3, 0.35, (0.0, 0.35, -1.4), 6))
    sc.meshes.append(scenes.icosphere(2, 0.3, (-1.2, 1.6, -3.0), 7))
    sc.meshes.append(scenes.icosphere(2, 0.25, (1.3, 1.9, -3.2), 8))
    sc.meshes.append(scenes.quad((-0.2, 0.0, -0.2), (0.2, 0.0, -0.2), (0.2, 0.0, 0.2), (-0.2, 0.0, 0.2), 9))
    n0 = len(sc.instances)
    sc.instances += [scenes.Instance(n0), scenes.Instance(n0 + 1), scenes.Instance(n0 + 2),
                     scenes.Instance(n0 + 3, translation=(-1.9, 1.2, -2.0), rotation=(0.3, 0.2, 1.4), scale=(1.5, 1.0, 2.5))]
    sc.sky_color = (0.5, 0.6, 0.8)
    sc.name = "parity_mix"
    return sc


@pytest.fixture(scope="module")
def setup():
    sc = parity_scene()
    ref = refdev.RefDevice(sc)  # light tree built by the reference's own light_tree_build
    ref_luts = ref.build_bsdf_lut()
    osc = orc.OracleScene(sc)
    osc.set_light_tree(*ref.light_tree[:3])
    osc.set_bsdf_luts(*ref_luts)
    return sc, ref, osc, ref_luts


def _rel(a, b, floor):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)


def test_bsdf_lut_product_and_oracle_vs_reference_kernels(setup):
    """The product's LUT kernels and the oracle's LUT loop against the reference's three LUT kernels. All three run the same
    65 536-sample quasi-Monte-Carlo sum per texel; the reference accumulates it serially with one FFMA per term, which the product
    replays in sample order (shade.cu: lut_chain) and the oracle writes out with fmaf(). The result is quantised with ceil() to R16
    (bsdf_lut.cuh:56). Bound: product tables BIT-IDENTICAL to the reference's on every texel; the IEEE/libm oracle within 2 R16 steps
    (3e-5 of full scale) on the 2D tables."""
    sc, ref, osc, ref_luts = setup
    dev = api.Device(0)
    dev.build_bsdf_lut()
    prod = dev.get_bsdf_lut()
    dev.destroy()
    for name, a, b in zip(("conductor", "glossy", "dielectric", "dielectric_inv"), prod, ref_luts):
        d = np.abs(a.astype(np.int64).reshape(-1) - b.astype(np.int64).reshape(-1))
        print(f"LUT {name}: product vs reference max |diff| {d.max()} R16 steps, texels differing {np.count_nonzero(d)} / {d.size}")
        assert d.max() == 0, name    # bit-identical tables (measured on B200: 0 of 67 584 texels differ)
    c, g = np.zeros(1024, np.uint16), np.zeros(1024, np.uint16)
    orc.lib().orc_bsdf_lut_generate(c.ctypes.data_as(orc.C.POINTER(orc.C.c_uint16)), g.ctypes.data_as(orc.C.POINTER(orc.C.c_uint16)), None, None, 0x10000, 0, 0)
    for name, a, b in zip(("conductor", "glossy"), (c, g), ref_luts):
        d = np.abs(a.astype(np.int64) - b.astype(np.int64).reshape(-1))
        print(f"LUT {name}: oracle vs reference max |diff| {d.max()} R16 steps, texels differing {np.count_nonzero(d)} / {d.size}")
        assert d.max() <= 2, name     # measured: 1 step on 7 texels


def test_tasks_create_vs_oracle(setup):
    sc, ref, osc, _ = setup
    T = 2 * refdev.THREADS_PER_BLOCK * 4
    K = -(-W * H // T)
    ref.configure(T // refdev.THREADS_PER_BLOCK, K)
    for sample_id in (0, 5):
        ref.set_state(0, sample_id)
        ref.launch("tasks_create")
        tasks = ref.task_states()[refdev.PRESORT]          # [slot][thread]
        counts = ref.download("trace_counts", np.uint16)
        assert int(counts.sum()) == W * H
        slot, thread = np.meshgrid(np.arange(K), np.arange(T), indexing="ij")
        live = slot < counts[None, :]
        t = tasks[live]
        pixel = (thread + slot * T)[live]                  # tasks_create walks pixel ids thread, thread + T, ...
        x, y = pixel % W, pixel // W
        # PathID bit-exact (cuda/utils.cuh:142-178)
        L = orc.lib()
        for k in range(0, t.size, 997):
            pid = L.orc_path_id_get(int(x[k]), int(y[k]), sample_id)
            assert tuple(t["path_id"][k]) == (pid.x, pid.y, pid.z)
        assert np.all(t["state"] == 0x01 | 0x02 | 0x08 | 0x10)
        one = L.orc_record_pack(orc.RGB(1.0, 1.0, 1.0))
        assert np.all(t["record"] == np.array([one.x, one.y], np.uint32))
        assert np.all(t["ior"] == 0)                       # medium_stack_ior_modify({}, 1.0, true) compresses to 0
        o, d = osc.camera_rays(sample_id)
        assert np.abs(t["origin"] - o[pixel]).max() <= 1e-6
        err = np.abs(t["ray"] - d[pixel]).max()
        print(f"tasks_create sample {sample_id}: max |ray - oracle| = {err:.3e}")
        assert err <= 2e-6                                 # fast-math rsqrt / quaternion rotation vs IEEE


def _record_unpack(p):
    """record_unpack, cuda/math.cuh:1595-1607: 3 x 21-bit truncated floats (1 step = 2^-12 relative)"""
    p = np.asarray(p, np.uint32)
    r = p[:, 0] & 0x1FFFFF
    g = (p[:, 0] >> 21) | ((p[:, 1] & 0x3FF) << 11)
    b = p[:, 1] >> 10
    return np.stack([(c << 11).astype(np.uint32).view(np.float32) for c in (r, g, b)], axis=1)


def _ray_unpack(p):
    """ray_unpack, cuda/math.cuh:1621-1635 (octahedral, 2 x 32 bit)"""
    x = np.asarray(p, np.uint32)[:, 0].astype(np.float64) / 0xFFFFFFFF * 2.0 - 1.0
    y = np.asarray(p, np.uint32)[:, 1].astype(np.float64) / 0xFFFFFFFF * 2.0 - 1.0
    z = 1.0 - np.abs(x) - np.abs(y)
    t = np.clip(-z, 0.0, 1.0)
    x = x + np.where(x >= 0, -t, t)
    y = y + np.where(y >= 0, -t, t)
    v = np.stack([x, y, z], axis=1)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


_tasks_from_vertices = refdev.tasks_from_vertices


@pytest.mark.parametrize("iteration", [0, 1, 2, 3])
def test_geometry_process_tasks_vs_oracle(setup, iteration):
    sc, ref, osc, _ = setup
    handles = osc.prim_handles()
    sample_id = 3
    vin, _pix = osc.path_vertices(sample_id, iteration)
    n = vin.size
    print(f"  iter {iteration}: {n} vertices")
    assert n > 1000
    depth = iteration if not (iteration == sc.max_ray_depth and iteration > 0) else iteration - 1
    want = osc.shade_vertices(vin, depth)

    T = 8 * refdev.THREADS_PER_BLOCK
    ref.configure(T // refdev.THREADS_PER_BLOCK, -(-n // T))
    tasks = _tasks_from_vertices(vin, handles)
    dl, res, bounce, trace_counts = ref.shade(tasks, depth)

    stats = {}
    # ---- NEE through the light tree (DeviceTaskDirectLightGeo) ----
    same_light = dl["geo_light_id"] == want["geo_light_id"]
    stats["geo light id equal"] = same_light.mean()
    sel = same_light & (want["geo_light_id"] != 0xFFFFFFFF)
    stats["geo lights valid"] = (want["geo_light_id"] != 0xFFFFFFFF).mean()
    if sel.any():
        stats["geo ray max abs"] = np.abs(dl["geo_ray"][sel] - want["geo_ray"][sel]).max()
        stats["geo dist p99 rel"] = np.percentile(_rel(dl["geo_dist"][sel], want["geo_dist"][sel], 1e-3), 99)
        stats["geo color p99 rel"] = np.percentile(_rel(dl["geo_color"][sel], want["geo_color"][sel], 1e-3).max(axis=1), 99)
        stats["geo color mean ratio"] = dl["geo_color"][sel].sum() / max(want["geo_color"][sel].sum(), 1e-20)
    # ---- BSDF-sampled NEE (DeviceTaskDirectLightBSDF) ----
    both = (dl["bsdf_prob"] != 0) == (want["bsdf_prob"] != 0)
    stats["bsdf rr decision equal"] = both.mean()
    sel = (dl["bsdf_prob"] != 0) & (want["bsdf_prob"] != 0)
    if sel.any():
        ray_ok = np.abs(dl["bsdf_ray"][sel] - want["bsdf_ray"][sel]).max(axis=1) < 1e-3
        stats["bsdf ray equal"] = ray_ok.mean()
        stats["bsdf prob p99 rel"] = np.percentile(_rel(dl["bsdf_prob"][sel][ray_ok], want["bsdf_prob"][sel][ray_ok], 1e-6), 99)
        stats["bsdf weight p99 rel"] = np.percentile(_rel(dl["bsdf_weight"][sel][ray_ok], want["bsdf_weight"][sel][ray_ok], 1e-3).max(axis=1), 99)
        stats["bsdf root_sum p99 rel"] = np.percentile(_rel(dl["bsdf_root_sum"][sel], want["bsdf_root_sum"][sel], 1e-6), 99)
    # ---- ambient NEE (packed record + packed direction) ----
    amb = want["amb_valid"] != 0
    assert amb.all()
    stats["amb ray max abs"] = np.abs(_ray_unpack(dl["amb_ray"]) - _ray_unpack(want["amb_ray"])).max()
    stats["amb color p99 rel"] = np.percentile(_rel(_record_unpack(dl["amb_color"]), _record_unpack(want["amb_color"]), 1e-4).max(axis=1), 99)
    # ---- emission into the result record ----
    stats["emission max abs"] = np.abs(res["color"] - want["emission"]).max()
    # ---- bounce tasks: compacted per thread in input order; match through the path id ----
    alive_want = want["bounce_alive"] != 0
    K = ref.tasks_per_thread
    slot, thread = np.meshgrid(np.arange(K), np.arange(T), indexing="ij")
    live = slot < trace_counts[None, :]
    b = bounce[live]
    key_in = vin["path_id"][:, 0].astype(np.int64) + vin["path_id"][:, 1].astype(np.int64) * 65536
    key_out = b["path_id"][:, 0].astype(np.int64) + b["path_id"][:, 1].astype(np.int64) * 65536
    assert np.unique(key_in).size == n
    order = np.argsort(key_in)
    pos = np.searchsorted(key_in[order], key_out)
    src = order[pos]
    assert np.all(key_in[src] == key_out)
    alive_ref = np.zeros(n, bool)
    alive_ref[src] = True
    stats["rr decision equal"] = (alive_ref == alive_want).mean()
    m = alive_want[src]
    bi, wi = b[m], want[src[m]]
    stats["bounce state equal"] = (bi["state"] == wi["bounce_state"]).mean()
    stats["bounce medium equal"] = (bi["ior"] == wi["bounce_medium_ior"]).mean()
    stats["bounce ignore handle ok"] = float(np.all(bi["instance_id"] == handles[vin["prim"][src[m]], 0]) and np.all(bi["tri_id"] == handles[vin["prim"][src[m]], 1]))
    stats["bounce origin max abs"] = np.abs(bi["origin"] - wi["bounce_origin"]).max()
    ray_ok = np.abs(bi["ray"] - wi["bounce_ray"]).max(axis=1) < 2e-3
    stats["bounce ray equal"] = ray_ok.mean()
    rec_rel = _rel(_record_unpack(bi["record"]), _record_unpack(wi["bounce_record"]), 1e-6).max(axis=1)[ray_ok]
    stats["bounce record p50 rel"] = np.percentile(rec_rel, 50)
    stats["bounce record p99 rel"] = np.percentile(rec_rel, 99)
    stats["bounce record max rel"] = rec_rel.max()
    worst = np.argsort(rec_rel)[-3:]
    for wv in worst:
        k = np.nonzero(ray_ok)[0][wv]
        print("   worst record:", _record_unpack(bi["record"][k:k + 1])[0], _record_unpack(wi["bounce_record"][k:k + 1])[0], "weight", wi["bounce_weight"][k],
              "transp", wi["is_transparent_pass"][k], "prim", vin["prim"][src[m]][k])
    for k, v in stats.items():
        print(f"  iter {iteration}: {k:55s} {v:.6g}")

    assert stats["geo light id equal"] >= 0.99
    if "geo ray max abs" in stats:
        assert stats["geo dist p99 rel"] <= 1e-3 and stats["geo color p99 rel"] <= 2e-2
        assert abs(stats["geo color mean ratio"] - 1.0) <= 1e-3
    assert stats["bsdf rr decision equal"] >= 0.995
    if "bsdf ray equal" in stats:
        assert stats["bsdf ray equal"] >= 0.99 and stats["bsdf prob p99 rel"] <= 2e-2 and stats["bsdf weight p99 rel"] <= 2e-2
    assert stats["amb ray max abs"] <= 1e-4 and stats["amb color p99 rel"] <= 1e-3
    assert stats["emission max abs"] <= 1e-5
    assert stats["rr decision equal"] >= 0.995
    assert stats["bounce state equal"] >= 0.995 and stats["bounce medium equal"] >= 0.995
    assert stats["bounce ignore handle ok"] == 1.0
    assert stats["bounce origin max abs"] <= 1e-4
    assert stats["bounce ray equal"] >= 0.99
    assert stats["bounce record p99 rel"] <= 1e-3


def test_sky_and_accumulation_kernels(setup):
    """sky_process_tasks + accumulation_collect_results(_first_sample) + accumulation_generate_result on hand-made records."""
    sc, ref, osc, _ = setup
    T = 2 * refdev.THREADS_PER_BLOCK
    n = T
    ref.configure(T // refdev.THREADS_PER_BLOCK, 1)
    rng = np.random.default_rng(7)
    L = orc.lib()
    pix = rng.permutation(W * H)[:n].astype(np.uint32)
    pix[1::2] = pix[0::2]                                   # two results per pixel inside one warp: exercises the __match_any merge
    rec = rng.uniform(0.05, 1.0, (n, 3)).astype(np.float32)
    packed = np.array([[p.x, p.y] for p in (L.orc_record_pack(orc.RGB(*map(float, r))) for r in rec)], np.uint32)
    unpacked = np.array([[c.r, c.g, c.b] for c in (L.orc_record_unpack(orc.Uint2(int(a), int(b))) for a, b in packed)], np.float32)
    tasks = np.zeros(n, refdev.TASK_STATE)
    tasks["state"] = np.where(np.arange(n) % 4 == 3, 0x01, 0x11)      # every 4th task lacks STATE_FLAG_ALLOW_AMBIENT
    tasks["path_id"][:, 0] = pix % W
    tasks["path_id"][:, 1] = pix // W
    tasks["ray"] = (0, 1, 0)
    tasks["record"] = packed
    tasks["results_index"] = np.arange(n)
    post = np.zeros((2, 1, T), refdev.TASK_STATE)
    post[refdev.POSTSORT, 0] = tasks
    ref.upload("task_states", refdev.interleave(post, T))
    res = np.zeros((1, 1, T), refdev.RESULT)
    res["index"][0, 0] = pix
    ref.upload("task_results", refdev.interleave(res, T))
    counts = np.zeros((refdev.SHADING_TASK_INDEX_TOTAL, T), np.uint16)
    counts[refdev.SHADING_TASK_INDEX_SKY] = 1
    ref.upload("task_counts", counts)
    ref.upload("task_offsets", np.zeros_like(counts))
    ref.upload("results_counts", np.ones(T, np.uint16))
    ref.set_state(0, 0)
    ref.launch("sky_process_tasks")
    got = ref.results()[0]
    sky = np.array(sc.sky_color, np.float32)
    want = np.where((tasks["state"] & 0x10)[:, None] != 0, sky[None, :] * unpacked, 0).astype(np.float32)
    assert np.abs(got["color"] - want).max() <= 1e-6
    assert np.all(got["index"] == pix)

    for name in ("frame_first_moment_r", "frame_first_moment_g", "frame_first_moment_b", "frame_second_moment_luminance"):
        ref.clear(name)
    ref.launch("accumulation_collect_results")
    ref.launch("accumulation_collect_results")               # two passes -> the sums double
    planes = np.stack([ref.download(nm, np.float32) for nm in
                       ("frame_first_moment_r", "frame_first_moment_g", "frame_first_moment_b", "frame_second_moment_luminance")])
    exp = np.zeros((4, W * H), np.float64)
    lum = want.astype(np.float64) ** 2 @ np.array([0.212655, 0.715158, 0.072187])   # color_luminance, cuda/math.cuh
    for c in range(3):
        np.add.at(exp[c], pix, 2.0 * want[:, c])
    np.add.at(exp[3], pix, 2.0 * lum)
    assert np.abs(planes - exp).max() <= 1e-5
    ref.set_state(0, 0, accumulated_samples=1)                # 2 samples in the planes
    ref.launch("accumulation_generate_result")
    out = np.stack([ref.download(nm, np.float32) for nm in ("frame_result_r", "frame_result_g", "frame_result_b")])
    assert np.abs(out - exp[:3] / 2.0).max() <= 1e-5
