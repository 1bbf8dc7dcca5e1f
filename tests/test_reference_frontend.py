"""The outer boundary as an executable: the REFERENCE's own command line front end (src/mandarin_duck/main.c, argument_parser.c,
mandarin_duck.c - unmodified) compiled against THIS repo's include/luminary/*.h and linked with libluminary_b200.so
(oracle/ref/Makefile: frontend -> oracle/_ref/MandarinDuckRef; only the SDL window is replaced by a stand-in).

  * CPU suite (this container, where /root/reference exists): the front end builds and its argument parser / help / version paths
    run; every struct of the public headers has the size and member offsets of the reference's header (both compiled by gcc); the
    "extra utilities" (array, host_memory, queue, ringbuffer, thread_status, log) behave as documented.
  * GPU suite: the prebuilt reference front end renders a benchmark (`scene.lum -b 2 name -o dir`) through the public API, and its
    PNGs equal the images of this repo's own front end bit for bit."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FRONTEND = os.path.join(ROOT, "oracle", "_ref", "MandarinDuckRef")
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "luminary_b200")

needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "mandarin_duck")), reason="needs /root/reference")


@needs_reference
def test_reference_front_end_builds_and_runs_against_this_api():
    subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle", "ref"), "frontend"])
    assert os.path.exists(FRONTEND)
    r = subprocess.run([FRONTEND, "--help"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    for word in ("OVERVIEW: Mandarin Duck Frontend for Luminary", "--benchmark", "--output", "--device", "Run the benchmark"):
        assert word in r.stdout
    r = subprocess.run([FRONTEND, "--version"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "Mandarin Duck for Luminary" in r.stdout
    # unknown argument: warn_message path of the parser (yellow), then a dry run through --help
    r = subprocess.run([FRONTEND, "--frobnicate", "--help"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "is unknown" in r.stdout


STRUCTS = ["LuminaryVec3", "LuminaryRGBF", "LuminaryRGBAF", "LuminaryARGB8", "LuminaryHostCreateInfo", "LuminaryRendererSettings", "LuminaryDeviceInfo",
           "LuminaryOutputProperties", "LuminaryOutputRequestProperties", "LuminaryPixelQueryResult", "LuminaryImage", "LuminaryCamera", "LuminaryOcean",
           "LuminarySky", "LuminaryCloudLayer", "LuminaryCloud", "LuminaryFog", "LuminaryParticles", "LuminaryMaterial", "LuminaryInstance"]
# every member an application can name, checked by offset
MEMBERS = {
    "LuminaryRendererSettings": ["width", "height", "max_ray_depth", "bridge_max_num_vertices", "undersampling", "supersampling", "enable_adaptive_sampling",
                                 "adaptive_sampling_max_sampling_rate", "adaptive_sampling_avg_sampling_rate", "adaptive_sampling_update_interval",
                                 "adaptive_sampling_exposure_aware", "adaptive_sampling_output_mode", "shading_mode", "region_x", "region_y", "region_width",
                                 "region_height"],
    "LuminaryCamera": ["pos", "rotation", "aperture_shape", "aperture_blade_count", "exposure", "tonemap", "agx_custom_slope", "agx_custom_power",
                       "agx_custom_saturation", "filter", "use_local_error_minimization", "bloom_blend", "dithering", "purkinje", "purkinje_kappa1",
                       "purkinje_kappa2", "wasd_speed", "mouse_speed", "smooth_movement", "smoothing_factor", "russian_roulette_threshold",
                       "use_color_correction", "color_correction", "film_grain", "camera_scale", "object_distance", "use_physical_camera", "thin_lens.fov",
                       "thin_lens.aperture_size", "physical.allow_reflections", "physical.focal_length", "physical.sensor_width"],
    "LuminarySky": ["geometry_offset", "azimuth", "altitude", "moon_azimuth", "moon_altitude", "moon_tex_offset", "sun_strength", "base_density",
                    "ozone_absorption", "steps", "stars_count", "stars_seed", "stars_intensity", "rayleigh_density", "mie_density", "ozone_density",
                    "rayleigh_falloff", "mie_falloff", "mie_diameter", "ground_visibility", "ozone_layer_thickness", "multiscattering_factor", "hdri_dim",
                    "hdri_samples", "aerial_perspective", "constant_color", "mode"],
    "LuminaryMaterial": ["id", "base_substrate", "albedo", "emission", "emission_scale", "roughness", "roughness_clamp", "refraction_index",
                         "emission_active", "thin_walled", "metallic", "colored_transparency", "roughness_as_smoothness", "normal_map_is_compressed",
                         "bidirectional_emission", "albedo_tex", "luminance_tex", "roughness_tex", "metallic_tex", "normal_tex"],
    "LuminaryInstance": ["id", "mesh_id", "position", "rotation", "scale"],
    "LuminaryOcean": ["active", "height", "amplitude", "frequency", "refractive_index", "water_type", "caustics_active", "caustics_ris_sample_count",
                      "caustics_domain_scale", "multiscattering", "triangle_light_contribution"],
    "LuminaryCloud": ["active", "initialized", "atmosphere_scattering", "low", "mid", "top", "offset_x", "offset_z", "density", "seed", "droplet_diameter",
                      "steps", "shadow_steps", "noise_shape_scale", "noise_detail_scale", "noise_weather_scale", "mipmap_bias", "octaves"],
    "LuminaryFog": ["active", "density", "droplet_diameter", "height", "dist"],
    "LuminaryParticles": ["active", "seed", "count", "albedo", "speed", "direction_altitude", "direction_azimuth", "phase_diameter", "scale", "size",
                          "size_variation"],
    "LuminaryPixelQueryResult": ["pixel_query_is_valid", "instance_id", "material_id", "depth", "rel_hit_pos"],
    "LuminaryImage": ["buffer", "width", "height", "ld", "meta_data.time", "meta_data.sample_count"],
    "LuminaryDeviceInfo": ["is_main_device", "is_unavailable", "is_enabled", "name", "memory_size", "allocated_memory_size"],
    "LuminaryOutputRequestProperties": ["sample_count", "width", "height"],
    "LuminaryOutputProperties": ["enabled", "width", "height"],
}
ENUMS = ["LUMINARY_SHADING_MODE_COUNT", "LUMINARY_ADAPTIVE_SAMPLING_OUTPUT_MODE_COUNT", "LUMINARY_FILTER_COUNT", "LUMINARY_TONEMAP_AGX_CUSTOM",
         "LUMINARY_TONEMAP_COUNT", "LUMINARY_APERTURE_COUNT", "LUMINARY_JERLOV_WATER_TYPE_9C", "LUMINARY_SKY_MODE_CONSTANT_COLOR", "LUMINARY_SKY_MODE_COUNT",
         "LUMINARY_MATERIAL_BASE_SUBSTRATE_TRANSLUCENT", "LUMINARY_OUTPUT_HANDLE_INVALID", "LUMINARY_ERROR_INVALID_DEVICE", "LUMINARY_ERROR_MISSING_DATA"]


def _layout_program(tmp_path, include_dir, extra_flags):
    lines = ["#include <stdio.h>", "#include <stddef.h>", "#define LUMINARY_INCLUDE_EXTRA_UTILS", "#include <luminary/luminary.h>", "int main(void) {"]
    for s in STRUCTS:
        lines.append(f'  printf("sizeof {s} %zu\\n", sizeof({s}));')
    for s, members in MEMBERS.items():
        for m in members:
            lines.append(f'  printf("offsetof {s}.{m} %zu\\n", offsetof({s}, {m}));')
    for e in ENUMS:
        lines.append(f'  printf("value {e} %llu\\n", (unsigned long long) {e});')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / ("layout_" + os.path.basename(os.path.dirname(include_dir)))
    subprocess.check_call(["/usr/bin/gcc", "-std=gnu11", "-w"] + extra_flags + ["-I", include_dir, str(src), "-o", str(exe)])
    return subprocess.check_output([str(exe)], text=True)


@needs_reference
def test_public_struct_layouts_equal_the_reference_headers(tmp_path):
    """sizeof / offsetof of every public struct and the enumerator values, compiled once against the reference's include/ and once
    against this repo's include/: identical output. (The reference's structs.h needs <stddef.h> spelled out)"""
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    theirs = _layout_program(tmp_path / "a", os.path.join(REF, "include"), ["-include", "stddef.h"])
    mine = _layout_program(tmp_path / "b", INC, [])
    assert mine == theirs
    assert len(mine.splitlines()) == len(STRUCTS) + sum(len(v) for v in MEMBERS.values()) + len(ENUMS)


@needs_reference
def test_public_api_declares_every_reference_function():
    """every LUMINARY_API function of the reference's public headers is declared here and exported by libluminary_b200.so"""
    import ctypes as C
    import re

    names = set()
    for f in os.listdir(os.path.join(REF, "include", "luminary")):
        text = open(os.path.join(REF, "include", "luminary", f)).read()
        names |= set(re.findall(r"\b((?:luminary_|_array_|array_|_host_|queue_|_queue_|ringbuffer_|_ringbuffer_|thread_status_)[a-z0-9_]+)\s*\(", text))
    names -= {"array_create", "array_resize", "array_push", "array_copy", "array_append", "array_set_num_elements", "array_destroy", "queue_create",
              "queue_destroy", "ringbuffer_create", "ringbuffer_destroy"}  # macros
    assert len(names) >= 85
    mine = "".join(open(os.path.join(INC, "luminary", f)).read() for f in os.listdir(os.path.join(INC, "luminary")))
    lib = C.CDLL(os.path.join(LIBDIR, "libluminary_b200.so"))
    missing = [n for n in sorted(names) if not re.search(r"\b" + n + r"\s*\(", mine) or not hasattr(lib, n)]
    assert not missing, missing


UTILS_TEST = r"""
#define LUMINARY_INCLUDE_EXTRA_UTILS
#include <luminary/luminary.h>
#include <stdio.h>
#include <string.h>
#define CHECK(c) do { if (!(c)) { printf("FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)
#define OK(e) CHECK((e) == LUMINARY_SUCCESS)
static bool eq_u32(void* a, void* b) { return *(uint32_t*) a == *(uint32_t*) b; }
int main(void) {
  luminary_init();
  /* arrays */
  uint32_t* a;
  OK(array_create(&a, sizeof(uint32_t), 2));
  for (uint32_t k = 0; k < 100; k++) OK(array_push(&a, &k));
  uint32_t n; OK(array_get_num_elements(a, &n)); CHECK(n == 100 && a[99] == 99 && a[0] == 0);
  size_t bytes; OK(array_get_size(a, &bytes)); CHECK(bytes == 400);
  uint32_t* b; OK(array_create(&b, sizeof(uint32_t), 1));
  OK(array_copy(&b, a)); OK(array_append(&b, a)); OK(array_get_num_elements(b, &n)); CHECK(n == 200 && b[100] == 0 && b[199] == 99);
  OK(array_set_num_elements(&b, 250)); CHECK(b[249] == 0);
  OK(array_resize(&b, 10)); OK(array_get_num_elements(b, &n)); CHECK(n == 10);
  OK(array_clear(b)); OK(array_get_num_elements(b, &n)); CHECK(n == 0);
  uint64_t* w; OK(array_create(&w, sizeof(uint64_t), 1));
  CHECK(array_append(&w, a) == LUMINARY_ERROR_API_EXCEPTION);          /* element sizes differ */
  uint32_t plain[4] = {0}; CHECK(array_get_num_elements(plain + 2, &n) != LUMINARY_SUCCESS);   /* not an array */
  OK(array_destroy(&a)); CHECK(a == NULL); OK(array_destroy(&b)); OK(array_destroy(&w));
  /* host memory */
  char* p; OK(host_malloc(&p, 16)); strcpy(p, "hello"); OK(host_realloc(&p, 1 << 20)); CHECK(strcmp(p, "hello") == 0); OK(host_free(&p)); CHECK(p == NULL);
  CHECK(host_free(&p) == LUMINARY_ERROR_ARGUMENT_NULL);
  /* queue */
  LuminaryQueue* q; OK(queue_create(&q, sizeof(uint32_t), 4));
  uint32_t v = 7; bool dup, got; OK(queue_push(q, &v)); OK(queue_push_unique(q, &v, eq_u32, &dup)); CHECK(dup);
  v = 8; OK(queue_push_unique(q, &v, eq_u32, &dup)); CHECK(!dup);
  v = 9; OK(queue_push(q, &v)); v = 10; OK(queue_push(q, &v)); v = 11; CHECK(queue_push(q, &v) != LUMINARY_SUCCESS);      /* full */
  uint32_t out; OK(queue_pop(q, &out, &got)); CHECK(got && out == 7); OK(queue_pop_blocking(q, &out, &got)); CHECK(got && out == 8);
  OK(queue_pop(q, &out, &got)); OK(queue_pop(q, &out, &got)); CHECK(out == 10); OK(queue_pop(q, &out, &got)); CHECK(!got);
  OK(queue_set_is_blocking(q, false)); OK(queue_pop_blocking(q, &out, &got)); CHECK(!got);                              /* returns instead of waiting */
  OK(queue_destroy(&q)); CHECK(q == NULL);
  /* ring buffer: FIFO arena with wrap-around */
  LuminaryRingBuffer* r; OK(ringbuffer_create(&r, 100));
  void *e1, *e2, *e3, *e4; OK(ringbuffer_allocate_entry(r, 40, &e1)); OK(ringbuffer_allocate_entry(r, 40, &e2));
  CHECK(ringbuffer_allocate_entry(r, 40, &e3) != LUMINARY_SUCCESS);     /* 20 left at the end, front still in use */
  OK(ringbuffer_release_entry(r, 40)); OK(ringbuffer_allocate_entry(r, 40, &e3)); CHECK(e3 == e1);                     /* wrapped to the front */
  CHECK(ringbuffer_allocate_entry(r, 30, &e4) != LUMINARY_SUCCESS);
  OK(ringbuffer_release_entry(r, 40)); OK(ringbuffer_release_entry(r, 40)); OK(ringbuffer_allocate_entry(r, 100, &e4));
  CHECK(ringbuffer_allocate_entry(r, 101, &e4) != LUMINARY_SUCCESS);
  OK(ringbuffer_destroy(&r));
  /* thread status */
  LuminaryThreadStatus* t; OK(thread_status_create(&t)); OK(thread_status_set_worker_name(t, "worker"));
  const char* s; OK(thread_status_get_worker_name(t, &s)); CHECK(strcmp(s, "worker") == 0);
  OK(thread_status_start(t, "task")); OK(thread_status_get_string(t, &s)); CHECK(strcmp(s, "task") == 0);
  double sec; OK(thread_status_get_time(t, &sec)); CHECK(sec >= 0.0); OK(thread_status_stop(t)); OK(thread_status_get_string(t, &s)); CHECK(s == NULL);
  OK(thread_status_destroy(&t));
  /* names, results */
  CHECK(strcmp(luminary_strings_tonemap[LUMINARY_TONEMAP_AGX_PUNCHY], "Agx Punchy") == 0);
  CHECK(strcmp(luminary_strings_sky_mode[LUMINARY_SKY_MODE_CONSTANT_COLOR], "Constant Color") == 0);
  CHECK(strlen(luminary_result_to_string(LUMINARY_ERROR_INVALID_DEVICE | LUMINARY_ERROR_PROPAGATED)) > 0);
  /* log */
  log_message("only the log sees %d", 1); info_message("info %s", "line"); warn_message("warn %d", 2);
  luminary_write_log();
  luminary_shutdown();
  printf("UTILS OK\n");
  return 0;
}
"""


def test_extra_utilities_behave_as_documented(tmp_path):
    src = tmp_path / "utils_test.c"
    src.write_text(UTILS_TEST)
    exe = tmp_path / "utils_test"
    subprocess.check_call(["/usr/bin/gcc", "-std=gnu11", "-Wall", "-I", INC, str(src), "-o", str(exe), "-L", LIBDIR, "-lluminary_b200", "-llumb200",
                           f"-Wl,-rpath,{LIBDIR}", "-lpthread"])
    r = subprocess.run([str(exe)], capture_output=True, text=True, cwd=tmp_path, timeout=60)
    assert r.returncode == 0 and "UTILS OK" in r.stdout, r.stdout + r.stderr
    log = (tmp_path / "luminary.log").read_text()
    assert "only the log sees 1" in log and "warn 2" in log


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(FRONTEND), reason="oracle/_ref/MandarinDuckRef not built (needs /root/reference)")
def test_reference_front_end_renders_a_benchmark_through_this_library(tmp_path):
    """`MandarinDuck scene.lum -b 2 name -o dir` with the reference's own main / argument parser / benchmark loop: request ladder
    1, 2, 4, 3 -> luminary_host_try_await_output / get_image / save_png / release_output. Same PNG bytes as this repo's front end."""
    import host_c
    from luminary_b200 import scenes

    sc = scenes.example_with_light(width=128, height=72, sphere_subdiv=2, max_ray_depth=3)
    obj = str(tmp_path / "scene.obj")
    scenes.write_obj(sc, obj)
    lum = str(tmp_path / "scene.lum")
    host_c.write_lum(lum, sc, "scene.obj", tonemap=1, dither=1, exposure=1.5)
    out_ref, out_own = tmp_path / "ref", tmp_path / "own"
    out_ref.mkdir()
    out_own.mkdir()
    r = subprocess.run([FRONTEND, lum, "-b", "2", "duck", "-o", str(out_ref), "--device", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Samples" in r.stdout   # "[%07.1fs] %05u Samples" of the reference's benchmark loop
    names = sorted(os.listdir(out_ref))
    assert names == ["Bench-00001-duck.png", "Bench-00002-duck.png", "Bench-00003-duck.png", "Bench-00004-duck.png", "BenchResults-duck.txt"]
    r2 = subprocess.run([host_c.CLI_PATH, lum, "-b", "2", "duck", "-o", str(out_own), "--device", "0"], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    for n in names[:-1]:
        a, b = host_c.png_decode_rgba(str(out_ref / n)), host_c.png_decode_rgba(str(out_own / n))
        assert a.shape == (72, 128, 4) and np.array_equal(a, b), n
        assert a[..., :3].mean() > 10
