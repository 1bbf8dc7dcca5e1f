"""Generates tests/golden/sky_ref.npz: outputs of the REFERENCE's own code for the procedural sky, for the inputs of
tests/sky_common.py. Needs a GPU and oracle/_ref/ (built in the container by oracle/ref/Makefile; travels to the GPU box):

  host C (libref_host.so)    device_struct_sky_convert -> sun / moon positions; sky_stars_update -> star catalogue
  CUDA (librefdev.so)        sky_compute_transmittance_lut, sky_compute_multiscattering_lut (cuda/sky.cuh:144-330) -> LUT samples
                             sky_process_tasks (cuda/sky.cuh:609-633) -> radiance of the miss rays of sky_common.miss_rays
                             sky_compute_hdri (cuda/sky_hdri.cuh:60-158) -> the HDRI mode's baked table, and the miss rays through it
                             sky_process_tasks across the moon's disc with the shipped surface textures
                             sky_process_inscattering_events (cuda/kernels.cuh:356-389) -> aerial perspective of explicit segments

Run on the GPU box:  python tests/golden/make_sky_golden.py gpurun_out/sky_ref.npz   (then copy the file to tests/golden/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refdev  # noqa: E402
import refhost  # noqa: E402
import sky_common  # noqa: E402
from luminary_b200 import scenes  # noqa: E402

W, H = 64, 36
TM_SUB = (slice(None, None, 4), slice(None, None, 8))


def main(out_path):
    out = {}
    for name, variant in sky_common.SKY_VARIANTS.items():
        sc = scenes.example_with_light(width=W, height=H, sphere_subdiv=1, max_ray_depth=2)
        sc.sky_mode, sc.sky = 0, dict(variant)
        ref = refdev.RefDevice(sc, light_tree=None)
        ds = np.frombuffer(ref.device_sky, np.float32)
        out[f"{name}/sun_pos"], out[f"{name}/moon_pos"] = ds[17:20].copy(), ds[20:23].copy()
        tm_low, tm_high, ms_low, ms_high = ref.build_sky_lut()
        out[f"{name}/tm_low"], out[f"{name}/tm_high"] = tm_low[TM_SUB].copy(), tm_high[TM_SUB].copy()
        out[f"{name}/ms_low"], out[f"{name}/ms_high"] = ms_low, ms_high
        stars, offsets = ref.set_stars()
        if stars.shape[0] <= 3000:
            out[f"{name}/stars"] = stars
        out[f"{name}/stars_head"], out[f"{name}/stars_sum"] = stars[:16].copy(), stars.astype(np.float64).sum(axis=0)
        out[f"{name}/stars_offsets"] = offsets
        rays = sky_common.miss_rays(out[f"{name}/sun_pos"], stars, W, H)
        n = rays["ray"].shape[0]
        T = 128 * ((n + 127) // 128)
        ref.configure(T // 128, 1)
        tasks = np.zeros(n, refdev.TASK_STATE)
        tasks["state"] = rays["state"]
        tasks["path_id"][:, 0], tasks["path_id"][:, 1], tasks["path_id"][:, 2] = rays["pixel"][:, 0], rays["pixel"][:, 1], rays["sample"]
        tasks["origin"], tasks["ray"] = rays["origin"], rays["ray"]
        tasks["record"] = sky_common.record_pack(np.ones((n, 3), np.float32))
        for depth in (0, 2):
            out[f"{name}/miss_color_depth{depth}"] = ref.sky(tasks, depth)
        if name in sky_common.HDRI_VARIANTS:
            # HDRI mode: the reference's sky_compute_hdri, then sky_process_tasks reading the baked table
            dim, samples = sky_common.HDRI_VARIANTS[name]
            sc.sky_mode = 1
            ref = refdev.RefDevice(sc, light_tree=None)
            ref.build_sky_lut()
            ref.set_stars()
            out[f"{name}/hdri_color"] = ref.build_sky_hdri(dim, samples, sky_common.HDRI_ORIGIN)
            ref.configure(T // 128, 1)
            out[f"{name}/hdri_miss_color"] = ref.sky(tasks, 0)
            print(name, "hdri mean", out[f"{name}/hdri_color"].mean(axis=(0, 1)), "miss mean", out[f"{name}/hdri_miss_color"].mean(axis=0))
        print(name, "sun", out[f"{name}/sun_pos"], "mean miss radiance", out[f"{name}/miss_color_depth0"].mean(axis=0),
              "max", out[f"{name}/miss_color_depth0"].max())
    # the moon's surface with the shipped textures
    from luminary_b200 import api
    albedo, normal = api.load_moon_textures()
    sc = scenes.example_with_light(width=W, height=H, sphere_subdiv=1, max_ray_depth=2)
    sc.sky_mode, sc.sky = 0, dict(sky_common.MOON_SKY)
    ref = refdev.RefDevice(sc, light_tree=None)
    ref.build_sky_lut()
    ref.set_moon_textures(albedo["data"], normal["data"])
    ds = np.frombuffer(ref.device_sky, np.float32)
    rays = sky_common.moon_rays(ds[20:23], W, H)
    n = rays["ray"].shape[0]
    ref.configure((n + 127) // 128, 1)
    tasks = np.zeros(n, refdev.TASK_STATE)
    tasks["state"] = rays["state"]
    tasks["path_id"][:, 0], tasks["path_id"][:, 1], tasks["path_id"][:, 2] = rays["pixel"][:, 0], rays["pixel"][:, 1], rays["sample"]
    tasks["origin"], tasks["ray"] = rays["origin"], rays["ray"]
    tasks["record"] = sky_common.record_pack(np.ones((n, 3), np.float32))
    out["moon/miss_color"] = ref.sky(tasks, 0)
    print("moon: mean radiance on the disc", out["moon/miss_color"][rays["angle"] < 0.0044].mean(axis=0))

    # aerial perspective: sky_process_inscattering_events on explicit segments, depths 0 (primary: 40 km base range) and 2
    sc = scenes.example_with_light(width=W, height=H, sphere_subdiv=1, max_ray_depth=2)
    sc.sky_mode, sc.sky = 0, dict(sky_common.AERIAL_SKY)
    ref = refdev.RefDevice(sc, light_tree=None)
    ref.build_sky_lut()
    seg = sky_common.aerial_segments(W, H)
    n = seg["ray"].shape[0]
    ref.configure((n + 127) // 128, 1)
    tasks = np.zeros(n, refdev.TASK_STATE)
    tasks["state"] = 0x1B
    tasks["path_id"][:, 0], tasks["path_id"][:, 1], tasks["path_id"][:, 2] = seg["pixel"][:, 0], seg["pixel"][:, 1], seg["sample"]
    tasks["origin"], tasks["ray"], tasks["depth"] = seg["origin"], seg["ray"], seg["t"]
    tasks["instance_id"], tasks["tri_id"] = 0, 0
    tasks["record"] = sky_common.record_pack(seg["record"])
    for depth in (0, 2):
        col, rec = ref.inscatter(tasks, depth)
        out[f"aerial/color_depth{depth}"], out[f"aerial/record_depth{depth}"] = col, rec
        print("aerial depth", depth, "mean in-scattering", col.mean(axis=0), "mean transmittance", (sky_common.record_unpack(rec) / np.maximum(sky_common.record_unpack(tasks["record"]), 1e-9)).mean(axis=0))

    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "sky_ref.npz"))
