"""GPU tests of the C host layer: the public Luminary API (include/luminary/luminary.h) end to end through the headless
benchmark front end, against the Python mirror of the same C ABI on the same scene, plus API error behaviour."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import host_c
from luminary_b200 import scenes

pytestmark = pytest.mark.gpu


def _scene_files(tmp_path, sc, **lum_kw):
    obj = str(tmp_path / "scene.obj")
    scenes.write_obj(sc, obj)
    lum = str(tmp_path / "scene.lum")
    host_c.write_lum(lum, sc, "scene.obj", **lum_kw)
    return lum, obj


def _python_reference_image(sc, obj, spp, tonemap, exposure, dither, supersampling=0, bloom=0.0, shading_mode=0):
    """Renders what the C host must have rendered: the mesh / materials as the C loader delivers them, one untransformed
    instance, sample ids 0..spp-1, internal resolution = output resolution << supersampling, same output parameters."""
    from luminary_b200 import api

    code, has, v, n, uv, mid, mats, ids = host_c.wavefront_load(obj, bidirectional=True)
    assert code == 0 and has
    scene = scenes.Scene("from_obj", [scenes.Mesh(v, n, uv, mid)], [scenes.Instance(0)], mats, sc.camera, sc.width << supersampling,
                         sc.height << supersampling, sc.max_ray_depth, sc.sky_mode, sc.sky_color)
    scene.sky = getattr(sc, "sky", None)
    scene.textures = list(host_c.last_textures)  # as lum_png_read decoded them (RGBA8 / RGBA16, wrap, linear, gAMA)
    dev = api.Device(0)
    dev.build_bsdf_lut()
    dev.load_bluenoise_1d(api.load_bluenoise_1d())
    if scene.sky_mode != 2:
        dev.load_moon_textures()              # the C host loads the moon's surface with its embedded data
    dev.load_scene(scene, light_tree="auto")  # integrates luminance-textured emitters on the device, as the C host does
    dev.set_shading_mode(shading_mode)
    dev.start_render()
    dev.render_samples(0, spp)
    img = dev.download_output_argb8(spp, exposure=exposure, tonemap=tonemap, dithering=dither, supersampling=supersampling, bloom_blend=bloom)
    st = dev.stats()
    dev.destroy()
    return img, st


def test_benchmark_front_end_matches_python_path(tmp_path):
    sc = scenes.example_with_light(width=128, height=72, sphere_subdiv=2, max_ray_depth=3)
    lum, obj = _scene_files(tmp_path, sc, tonemap=1, dither=1, exposure=1.5)
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([host_c.CLI_PATH, lum, "-b", "3", "run", "-o", str(out), "--device", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    # ladder of mandarin_duck.c:53-98 for N = 3: 1, 2, 4, 3, 8, 6
    names = sorted(os.listdir(out))
    assert names == ["Bench-00001-run.png", "Bench-00002-run.png", "Bench-00003-run.png", "Bench-00004-run.png", "Bench-00006-run.png",
                     "Bench-00008-run.png", "BenchResults-run.txt"]
    rows = [tuple(x.split(",")) for x in open(out / "BenchResults-run.txt").read().split("\n") if x]
    assert sorted(int(a) for a, _ in rows) == [1, 2, 3, 4, 6, 8]
    times = {int(a): float(b) for a, b in rows}
    assert all(times[a] > 0 for a in times) and times[8] > times[1]  # cumulative GPU seconds
    assert "Mrays/s" in r.stdout

    # the host's default settings render at 2 x 2 internal resolution (supersampling 1, reference settings.c:14)
    ref, st = _python_reference_image(sc, obj, 8, tonemap=1, exposure=1.5, dither=True, supersampling=1)
    got = host_c.png_decode_rgba(str(out / "Bench-00008-run.png"))
    assert got.shape == (72, 128, 4)
    # PNG is r, g, b, a; the device image is b, g, r, a
    assert np.array_equal(got[..., 0], ref[..., 2]) and np.array_equal(got[..., 1], ref[..., 1]) and np.array_equal(got[..., 2], ref[..., 0])
    assert (got[..., 3] == 255).all()
    assert ref[..., :3].mean() > 20


@pytest.mark.parametrize("mode", [0, 1])
def test_procedural_sky_through_the_public_api(tmp_path, mode):
    """A scene file under LUMINARY_SKY_MODE_DEFAULT (the reference's default, sky.c:39) with sky settings of its own: the C host
    maps every LuminarySky field onto the device library; an open-top room is lit by sun and sky and equals the Python mirror."""
    sc = scenes.example_with_light(width=96, height=54, sphere_subdiv=2, max_ray_depth=3)
    room = sc.meshes[0]
    keep = np.ones(room.num_tris, bool)
    keep[2:4] = False  # no ceiling
    sc.meshes[0] = scenes.Mesh(room.vertex[keep], room.normal[keep], room.uv[keep], room.material[keep])
    sc.sky_mode = mode   # 1: the HDRI mode, baked from the camera position when the render starts
    sc.sky = dict(azimuth=1.2, altitude=0.9, mie_density=1.5, stars_count=500, stars_seed=4, steps=20, ozone_absorption=0)
    if mode == 1:
        sc.sky.update(hdri_dim=64, hdri_samples=8)
    lum, obj = _scene_files(tmp_path, sc, tonemap=1, dither=0, exposure=1.0)
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([host_c.CLI_PATH, lum, "-b", "2", "sky", "-o", str(out), "--device", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    ref, st = _python_reference_image(sc, obj, 4, tonemap=1, exposure=1.0, dither=False, supersampling=1)
    got = host_c.png_decode_rgba(str(out / "Bench-00004-sky.png"))
    assert np.array_equal(got[..., 0], ref[..., 2]) and np.array_equal(got[..., 1], ref[..., 1]) and np.array_equal(got[..., 2], ref[..., 0])
    assert ref[..., :3].mean() > 40, "sun and sky light the room"


def test_textured_obj_through_the_public_api(tmp_path):
    """*.obj + *.mtl with map_Kd / map_Ke / map_Ns / map_Bump + PNG files -> luminary_host_load_lum_file -> render: the C host
    (PNG reader, texture upload, textured kernel variants) must produce the image of the Python mirror bit for bit."""
    sc = scenes.textured_example(width=96, height=54, sphere_subdiv=2, max_ray_depth=3)
    lum, obj = _scene_files(tmp_path, sc, tonemap=1, dither=1, exposure=1.5, bloom=0.05)  # the camera's bloom travels too
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([host_c.CLI_PATH, lum, "-b", "2", "tex", "-o", str(out), "--device", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    ref, st = _python_reference_image(sc, obj, 4, tonemap=1, exposure=1.5, dither=True, supersampling=1, bloom=0.05)
    got = host_c.png_decode_rgba(str(out / "Bench-00004-tex.png"))
    assert got.shape == (54, 96, 4)
    assert np.array_equal(got[..., 0], ref[..., 2]) and np.array_equal(got[..., 1], ref[..., 1]) and np.array_equal(got[..., 2], ref[..., 0])
    assert ref[..., :3].mean() > 5
    # and the textures matter: the same files without the map_ statements render differently
    mtl = open(str(tmp_path / "scene.mtl")).read()
    open(str(tmp_path / "scene.mtl"), "w").write("\n".join(l for l in mtl.split("\n") if not l.startswith("map_")))
    plain, _ = _python_reference_image(sc, obj, 4, tonemap=1, exposure=1.5, dither=True, supersampling=1)
    assert np.abs(plain[..., :3].astype(np.int32) - ref[..., :3].astype(np.int32)).mean() > 2.0


def _api():
    L = host_c.lib()
    for n in host_c.declared_api_functions():
        if n not in ("luminary_result_to_string", "luminary_b200_last_error", "luminary_init", "luminary_shutdown"):
            getattr(L, n).restype = C.c_uint64
    return L


def test_public_api_error_behaviour(tmp_path):
    L = _api()
    L.luminary_init()
    host = C.c_void_p()
    assert L.luminary_host_create(None, C.c_uint32(1)) == 1  # NULL argument
    assert L.luminary_host_create(C.byref(host), C.c_uint32(1)) == 0
    s = host_c.Settings()
    assert L.luminary_host_get_settings(host, C.byref(s)) == 0
    assert (s.width, s.height, s.max_ray_depth) == (2560, 1440, 4)  # settings.c:9-11
    assert L.luminary_host_get_settings(host, None) == 1
    assert L.luminary_host_get_ocean(host, None) == 1
    # entities outside the path: real layouts, the reference's defaults, readable / writable while inactive; activation is refused
    class Ocean(C.Structure):
        _fields_ = [("active", C.c_bool), ("height", C.c_float), ("amplitude", C.c_float), ("frequency", C.c_float), ("refractive_index", C.c_float),
                    ("water_type", C.c_uint32), ("caustics_active", C.c_bool), ("caustics_ris_sample_count", C.c_uint32),
                    ("caustics_domain_scale", C.c_float), ("multiscattering", C.c_bool), ("triangle_light_contribution", C.c_bool)]
    oc = Ocean()
    assert L.luminary_host_get_ocean(host, C.byref(oc)) == 0
    assert not oc.active and abs(oc.refractive_index - 1.333) < 1e-6 and oc.water_type == 2 and oc.caustics_ris_sample_count == 32  # ocean.c:6-22
    oc.height = 3.5
    assert L.luminary_host_set_ocean(host, C.byref(oc)) == 0
    oc2 = Ocean()
    assert L.luminary_host_get_ocean(host, C.byref(oc2)) == 0 and oc2.height == 3.5
    oc.active = True
    assert (L.luminary_host_set_ocean(host, C.byref(oc)) & 0xFF) == 2
    assert L.luminary_host_request_sky_hdri_build(host) == 0  # the HDRI mode is on the path: bakes at the next render
    m = host_c.Material()
    assert (L.luminary_host_get_material(host, C.c_uint16(0), C.byref(m)) & 0xFF) == 3  # no materials yet
    out = C.c_uint32(0)
    assert (L.luminary_host_try_await_output(host, C.c_uint32(17), C.byref(out)) & 0xFF) == 3 and out.value == 0xFFFFFFFF
    assert s.supersampling == 1  # settings.c:14
    # settings the path does not implement are rejected when a render is started, not silently ignored
    assert s.enable_adaptive_sampling and s.adaptive_sampling_update_interval == 64  # settings.c:15-18
    s.width, s.height, s.adaptive_sampling_output_mode = 64, 36, 7  # not a LuminaryAdaptiveSamplingOutputMode
    assert L.luminary_host_set_settings(host, C.byref(s)) == 0
    assert (L.luminary_host_start_new_render(host) & 0xFF) == 3
    s.adaptive_sampling_output_mode = 0
    s.supersampling = 5
    assert L.luminary_host_set_settings(host, C.byref(s)) == 0
    assert (L.luminary_host_start_new_render(host) & 0xFF) == 3
    s.supersampling = 0
    s.width = 0
    assert (L.luminary_host_set_settings(host, C.byref(s)) & 0xFF) == 3
    # missing scene file
    path = C.c_void_p()
    assert L.luminary_path_create(C.byref(path)) == 0
    assert L.luminary_path_set_from_string(path, str(tmp_path / "nope.lum").encode()) == 0
    assert (L.luminary_host_load_lum_file(host, path) & 0xFF) == 7
    assert L.luminary_path_destroy(C.byref(path)) == 0 and not path.value
    cnt = C.c_uint32(0)
    assert L.luminary_host_get_device_count(host, C.byref(cnt)) == 0 and cnt.value >= 1
    assert L.luminary_host_destroy(C.byref(host)) == 0 and not host.value
    assert L.luminary_host_destroy(C.byref(host)) == 1  # destroyed handles are nulled (structs via T**)


def test_api_render_later_request_continues_accumulating(tmp_path):
    """request_output after the first outputs were delivered keeps accumulating in the same render (no restart)."""
    L = _api()
    sc = scenes.example_with_light(width=64, height=36, sphere_subdiv=1, max_ray_depth=2)
    lum, obj = _scene_files(tmp_path, sc, tonemap=0, dither=0, exposure=1.0)
    L.luminary_init()
    host = C.c_void_p()
    assert L.luminary_host_create(C.byref(host), C.c_uint32(1)) == 0
    path = C.c_void_p()
    L.luminary_path_create(C.byref(path))
    L.luminary_path_set_from_string(path, lum.encode())
    assert L.luminary_host_load_lum_file(host, path) == 0
    L.luminary_path_destroy(C.byref(path))
    st = host_c.Settings()
    assert L.luminary_host_get_settings(host, C.byref(st)) == 0
    st.supersampling = 0  # a v4 scene file cannot express it: drivers set it through the API
    assert L.luminary_host_set_settings(host, C.byref(st)) == 0
    n = C.c_uint32(0)
    assert L.luminary_host_get_num_materials(host, C.byref(n)) == 0 and n.value == 1 + len(sc.materials)
    assert L.luminary_host_get_num_instances(host, C.byref(n)) == 0 and n.value == 1

    class Req(C.Structure):
        _fields_ = [("sample_count", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32)]

    class Meta(C.Structure):
        _fields_ = [("time", C.c_float), ("sample_count", C.c_uint32)]

    class Image(C.Structure):
        _fields_ = [("buffer", C.POINTER(C.c_uint8)), ("width", C.c_uint32), ("height", C.c_uint32), ("ld", C.c_size_t), ("meta_data", Meta)]

    def fetch(promise):
        out = C.c_uint32(0xFFFFFFFF)
        assert L.luminary_b200_host_wait_idle(host) == 0
        assert L.luminary_host_try_await_output(host, promise, C.byref(out)) == 0
        assert out.value != 0xFFFFFFFF
        im = Image()
        assert L.luminary_host_get_image(host, out, C.byref(im)) == 0
        arr = np.ctypeslib.as_array(im.buffer, shape=(im.height, im.ld, 4)).copy()
        meta = (im.meta_data.sample_count, im.meta_data.time)
        assert L.luminary_host_release_output(host, out) == 0
        return arr, meta

    p2, p5 = C.c_uint32(), C.c_uint32()
    assert L.luminary_host_request_output(host, Req(2, 64, 36), C.byref(p2)) == 0
    assert L.luminary_host_start_new_render(host) == 0
    a2, m2 = fetch(p2)
    assert L.luminary_host_request_output(host, Req(5, 64, 36), C.byref(p5)) == 0
    a5, m5 = fetch(p5)
    assert m2[0] == 2 and m5[0] == 5 and m5[1] > m2[1]
    ref, _ = _python_reference_image(sc, obj, 5, tonemap=0, exposure=1.0, dither=False)
    assert np.array_equal(a5, ref)
    rays = C.c_uint64(0)
    assert L.luminary_b200_host_get_ray_count(host, C.byref(rays)) == 0 and rays.value > 5 * 64 * 36

    # debug shading mode through the public settings (LuminaryRendererSettings.shading_mode): the identification image of the one-bounce
    # debug queue, un-tone-mapped; an unknown mode is refused when the render starts
    st.shading_mode = 4  # LUMINARY_SHADING_MODE_IDENTIFICATION
    assert L.luminary_host_set_settings(host, C.byref(st)) == 0
    p1 = C.c_uint32()
    assert L.luminary_host_request_output(host, Req(1, 64, 36), C.byref(p1)) == 0
    assert L.luminary_host_start_new_render(host) == 0
    a1, m1 = fetch(p1)
    ref_id, _ = _python_reference_image(sc, obj, 1, tonemap=0, exposure=1.0, dither=False, shading_mode=4)
    assert m1[0] == 1 and np.array_equal(a1, ref_id) and not np.array_equal(a1, a5)
    assert len(np.unique(a1.reshape(-1, 4), axis=0)) > 20  # one colour per (instance, triangle)
    st.shading_mode = 9
    assert L.luminary_host_set_settings(host, C.byref(st)) == 0
    assert (L.luminary_host_start_new_render(host) & 0xFF) == 3
    assert L.luminary_host_destroy(C.byref(host)) == 0


def test_adaptive_sampling_through_the_public_api(tmp_path):
    """LuminaryRendererSettings.enable_adaptive_sampling through luminary_host_set_settings: 2 + 4 + 1 executions (stage 0, stage 1,
    one of stage 2) on the C host equal the Python mirror driving the same C ABI (bytes within 1: the adaptive executions add
    their results with float atomics, so the summation order differs from run to run)."""
    from luminary_b200 import api

    sc = scenes.example_with_light(width=64, height=36, sphere_subdiv=1, max_ray_depth=2)
    lum, obj = _scene_files(tmp_path, sc, tonemap=4, dither=0, exposure=1.0)
    L = _api()
    L.luminary_init()
    host = C.c_void_p()
    assert L.luminary_host_create(C.byref(host), C.c_uint32(1)) == 0
    path = C.c_void_p()
    L.luminary_path_create(C.byref(path))
    L.luminary_path_set_from_string(path, lum.encode())
    assert L.luminary_host_load_lum_file(host, path) == 0
    L.luminary_path_destroy(C.byref(path))
    st = host_c.Settings()
    assert L.luminary_host_get_settings(host, C.byref(st)) == 0
    st.supersampling = 0
    st.enable_adaptive_sampling = True
    st.adaptive_sampling_update_interval = 2
    st.adaptive_sampling_avg_sampling_rate = 2
    st.adaptive_sampling_max_sampling_rate = 8
    st.adaptive_sampling_exposure_aware = True
    assert L.luminary_host_set_settings(host, C.byref(st)) == 0

    class Req(C.Structure):
        _fields_ = [("sample_count", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32)]

    class Meta(C.Structure):
        _fields_ = [("time", C.c_float), ("sample_count", C.c_uint32)]

    class Image(C.Structure):
        _fields_ = [("buffer", C.POINTER(C.c_uint8)), ("width", C.c_uint32), ("height", C.c_uint32), ("ld", C.c_size_t), ("meta_data", Meta)]

    promise = C.c_uint32()
    assert L.luminary_host_request_output(host, Req(7, 64, 36), C.byref(promise)) == 0
    assert L.luminary_host_start_new_render(host) == 0
    assert L.luminary_b200_host_wait_idle(host) == 0
    out = C.c_uint32(0xFFFFFFFF)
    assert L.luminary_host_try_await_output(host, promise, C.byref(out)) == 0 and out.value != 0xFFFFFFFF
    im = Image()
    assert L.luminary_host_get_image(host, out, C.byref(im)) == 0
    got = np.ctypeslib.as_array(im.buffer, shape=(im.height, im.ld, 4)).copy()
    rays = C.c_uint64(0)
    assert L.luminary_b200_host_get_ray_count(host, C.byref(rays)) == 0
    assert L.luminary_host_destroy(C.byref(host)) == 0

    code, has, v, n, uv, mid, mats, ids = host_c.wavefront_load(obj, bidirectional=True)
    scene = scenes.Scene("from_obj", [scenes.Mesh(v, n, uv, mid)], [scenes.Instance(0)], mats, sc.camera, sc.width, sc.height, sc.max_ray_depth,
                         sc.sky_mode, sc.sky_color)
    dev = api.Device(0)
    dev.build_bsdf_lut()
    dev.load_scene(scene, light_tree="auto")
    dev.update_adaptive_sampling(max_sampling_rate=8, avg_sampling_rate=2, update_interval=2, exposure_aware=True, exposure=1.0, tonemap=4)
    dev.start_render()
    dev.render_executions(7)
    state = dev.adaptive_state()
    ref = dev.download_output_argb8(7, exposure=1.0, tonemap=4, dithering=False)
    stats = dev.stats()
    dev.destroy()
    assert state["stage_id"] == 2 and state["executions"] == [2, 4, 1, 0, 0]
    assert state["paths_traced"] > 7 * 64 * 36  # the adaptive stages spent more than one sample per pixel
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1 and np.count_nonzero(diff) <= 0.01 * diff.size
    assert rays.value == stats["closest_rays"] + stats["shadow_rays"] + stats["light_rays"]


def test_two_devices_in_process_match_one_device(tmp_path):
    from luminary_b200 import api

    if api.device_count() < 2:
        pytest.skip("needs two GPUs")
    sc = scenes.example_with_light(width=96, height=54, sphere_subdiv=2, max_ray_depth=3)
    lum, obj = _scene_files(tmp_path, sc, tonemap=0, dither=0, exposure=1.0)
    imgs = []
    for devices in (["--device", "0"], ["--device", "0", "--device", "1"]):
        out = tmp_path / ("out%d" % len(devices))
        out.mkdir()
        r = subprocess.run([host_c.CLI_PATH, lum, "-b", "3", "run", "-o", str(out), "--supersampling", "0"] + devices, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        imgs.append(host_c.png_decode_rgba(str(out / "Bench-00008-run.png")).astype(np.int32))
    # same sample ids, different float summation order across devices: bytes agree up to rounding at a quantisation step
    assert np.abs(imgs[0] - imgs[1]).max() <= 1


def test_two_devices_share_the_adaptive_sampler(tmp_path):
    """Adaptive sampling over two devices of one process: the executions of a stage are dealt round robin, the planes are combined
    with one NCCL reduce at every stage boundary (lumb200_comm_reduce_planes_all), the main device builds the stage and its counts
    are broadcast (lumb200_comm_broadcast_adaptive_words_all + lumb200_device_adopt_adaptive_stage). Interval 2: stages switch after
    2, 6 and 14 executions, the 16-sample output has seen three stage builds. Same schedule on one device: the stage counts come from
    identically valued planes (float summation order aside), so the images agree to a byte value on >= 99.5 % of the bytes."""
    from luminary_b200 import api

    if api.device_count() < 2:
        pytest.skip("needs two GPUs")
    sc = scenes.example_with_light(width=96, height=54, sphere_subdiv=2, max_ray_depth=3)
    lum, obj = _scene_files(tmp_path, sc, tonemap=0, dither=0, exposure=1.0)
    imgs = []
    for devices in (["--device", "0"], ["--device", "0", "--device", "1"]):
        out = tmp_path / ("as%d" % len(devices))
        out.mkdir()
        r = subprocess.run([host_c.CLI_PATH, lum, "-b", "4", "run", "-o", str(out), "--supersampling", "0", "--adaptive", "1", "--adaptive-interval", "2"]
                           + devices, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "peer copies" not in r.stderr   # the NCCL communicator was available
        imgs.append(host_c.png_decode_rgba(str(out / "Bench-00016-run.png")).astype(np.int32))
    d = np.abs(imgs[0] - imgs[1])
    print(f"  adaptive 1 vs 2 devices: max byte difference {d.max()}, bytes within 1: {(d <= 1).mean():.5f}")
    assert (d <= 1).mean() >= 0.995
    assert imgs[0][..., :3].mean() > 10
