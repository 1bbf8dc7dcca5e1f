#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 wavefront path-tracing hot path.

Metric (BASELINE.json): Mrays/s (and spp/s) at 1080p, 5 bounces, on the procedural 1M-triangle atrium (configs[1]).
A "step" is one full-frame sample pass (1 spp): ray generation, 6 closest-hit generations, material sort, shading
with light-tree NEE, shadow rays, accumulation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload atrium1m|atrium100k|example]

* value      whole-job Mrays/s with the scene resident in HBM: rays of all ranks / max-over-ranks device time of the
             K timed steps (CUDA events on the library's stream, barrier + synchronize on both sides).
* e2e        the same metric through the public API with HOST buffers in the timed region: every step uploads the
             camera and settings entities from host structs and downloads the resolved RGB frame to host memory.
* roofline   the kernel stage with the largest share of the step (measured, not assumed). achieved = algorithmic bytes per
             launch / average launch time; the launch times come from a SEPARATE profiled run of the same K steps (CUDA
             events around every stage), so that `value` is timed without any instrumentation. Algorithmic bytes, no
             cache credit: closest-hit ray 48 B + 80 B per BVH8 node visited + 48 B per triangle tested;
             shadow ray 36 B + the same traversal term; shaded vertex 246 B + NEE (16 B + 48 B x root sections + 64 B x
             light-tree nodes descended + 8 x 96 B candidate gathers). Node / triangle / tree-node counts come from one
             instrumented pass of the same kernels on the same rays.
* --spp N    additionally renders N samples per pixel from start_render to the resolved frame in host memory (incl. the
             multi-GPU reduce) and reports the MEASURED wall time as time_to_spp (config 5: --workload atrium4k --spp 1024).
* cpu_baseline  the CPU oracle (plain-C restatement of the reference's device functions, `oracle/`) timed on the host
             cores of this box on a bounded pixel region of the same workload.
* --impl reference  runs that CPU implementation instead (the reference renders GPU-only through OptiX and cannot be
             built here), same metric / config, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from luminary_b200 import scenes, sharding  # noqa: E402

WORKLOADS = {
    "atrium1m": dict(fn=lambda: scenes.atrium(1_000_000, 1920, 1080, 5), desc="procedural 1M-triangle atrium (S1), 1920x1080, 5 bounces"),
    "atrium100k": dict(fn=lambda: scenes.atrium(100_000, 1920, 1080, 5), desc="procedural 100k-triangle atrium, 1920x1080, 5 bounces"),
    "example": dict(fn=lambda: scenes.example_with_light(960, 540, 5, 5), desc="Example box + spheres + light, 960x540, 5 bounces"),
    "divergence": dict(fn=lambda: scenes.divergence(1_000_000, 1920, 1080, 8), desc="divergence stress (S3), 1920x1080, 8 bounces"),
    "terrain10m": dict(fn=lambda: scenes.terrain(2236, 50_000, 1920, 1080, 5),
                       desc="procedural 10M-triangle terrain + 100k emissive triangles (S2), 1920x1080, 5 bounces"),
    "atrium1m_tex": dict(fn=lambda: scenes.atrium_textured(1_000_000, 1920, 1080, 5),
                         desc="procedural 1M-triangle atrium (S1) with material textures (albedo / roughness / normal maps, alpha cut-out columns, "
                              "luminance-textured lights), 1920x1080, 5 bounces"),
    "divergence_sky": dict(fn=lambda: _with_sky(scenes.divergence(1_000_000, 1920, 1080, 8), 0),
                           desc="divergence stress (S3, open ceiling) under the procedural sky (LUMINARY_SKY_MODE_DEFAULT: ray-marched misses, sun NEE), 1920x1080, 8 bounces"),
    "divergence_hdri": dict(fn=lambda: _with_sky(scenes.divergence(1_000_000, 1920, 1080, 8), 1),
                            desc="divergence stress (S3, open ceiling) under the baked sky (LUMINARY_SKY_MODE_HDRI: table look-ups, ambient + sun NEE), 1920x1080, 8 bounces"),
    "atrium4k": dict(fn=lambda: scenes.atrium(1_000_000, 3840, 2160, 5), desc="procedural 1M-triangle atrium (S1), 3840x2160, 5 bounces"),
}


def _with_sky(scene, mode):
    scene.sky_mode = mode
    scene.sky = dict(azimuth=1.2, altitude=1.0)
    return scene


def ncu_traffic(kernel, workload):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t[workload][kernel]
        return float(e["dram_bytes_per_launch"]), e.get("source", "")
    except Exception:
        return None, ""


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def oracle_scene(scene, luts, light_tree):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc

    osc = orc.OracleScene(scene)
    osc.set_bsdf_luts(*luts)
    if light_tree is not None:
        osc.set_light_tree(*light_tree)
    return osc


def cpu_region(scene, target_pixels):
    """Centered pixel rectangle with about target_pixels pixels (same aspect as the frame)."""
    w, h = scene.width, scene.height
    f = min(1.0, (target_pixels / float(w * h)) ** 0.5)
    rw, rh = max(8, int(w * f)), max(8, int(h * f))
    x0, y0 = (w - rw) // 2, (h - rh) // 2
    return x0, y0, x0 + rw, y0 + rh


def cpu_luts():
    """CPU-generated conductor/glossy LUTs (the dielectric tables are not used by the opaque bench scenes)."""
    import ctypes as C

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc

    c, g = np.zeros(1024, np.uint16), np.zeros(1024, np.uint16)
    d, di = np.full(32768, 65535, np.uint16), np.full(32768, 65535, np.uint16)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))
    orc.lib().orc_bsdf_lut_generate(p(c), p(g), p(d), p(di), 4096, 0, 0)
    return c, g, d, di


# The CPU legs use every host thread explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which would otherwise
# silently turn the multi-threaded baseline into a scalar one at N > 1.
CPU_THREADS = os.cpu_count() or 1


def _cpu_rays(info):
    return info["closest_rays"] + info["shadow_rays"] + info["light_enum_rays"]


def plan_cpu_sample(osc, scene, target_seconds, calib_pixels=40000):
    """Bounded CPU sample: a short calibration render fixes the oracle's pixel rate on this box, then the sample is
    sized to about target_seconds of CPU work - a centred region of the frame, or the whole frame at several spp."""
    region = cpu_region(scene, calib_pixels)
    _, info = osc.render(900000, 1, threads=CPU_THREADS, region=region)
    px = (region[2] - region[0]) * (region[3] - region[1])
    rate = px / max(info["seconds"], 1e-6)  # pixel-samples per second
    want = rate * target_seconds
    full = scene.width * scene.height
    if want >= full:
        return (0, 0, scene.width, scene.height), max(1, min(int(want / full), 64))
    return cpu_region(scene, int(want)), 1


def run_cpu(osc, region, spp, first_sample=0):
    planes, info = osc.render(first_sample, spp, threads=CPU_THREADS, region=region)
    return _cpu_rays(info), info["seconds"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="atrium1m", choices=sorted(WORKLOADS))
    ap.add_argument("--no-sort", action="store_true", help="disable the material-keyed queue sort (config 4 comparison)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (parameter sweeps)")
    ap.add_argument("--no-measure", action="store_true", help="skip the instrumented pass (ncu captures: keeps the launch sequence to plain passes)")
    ap.add_argument("--spp", type=int, default=0, help="also measure the wall time of an N-spp render (start_render -> resolved frame on the host)")
    ap.add_argument("--no-reduce-check", action="store_true", help="skip the single-rank re-render that checks the reduced planes (N > 1)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the bounded CPU baseline sample (per step for --impl reference)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1

    config = {"workload": wl["desc"], "spp_per_step": 1, "sort_by_material": not args.no_sort, "scene_resident": True,
              "l2_policy": "inputs larger than L2: the per-step working set (path state and ray queues 830 MB + scene 61 MB at 1080p) exceeds the 126 MB L2; no explicit flush",
              "parallelism": f"sample-id sharding x{world}, scene replicated, one ncclReduce of the 4 accumulation planes behind the C ABI (lumb200_comm_reduce_planes)"}

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return
        scene = wl["fn"]()
        from luminary_b200 import api

        lt = api.build_light_tree(scene)
        luts = cpu_luts()
        osc = oracle_scene(scene, luts, lt)
        # every step is a bounded sample of one pass; the whole run (warmup + steps) is sized to ~2 minutes
        per_step = max(0.5, min(args.cpu_seconds, 120.0 / (steps + min(warmup, 1))))
        region, spp = plan_cpu_sample(osc, scene, per_step)
        for k in range(min(warmup, 1)):
            osc.render(1000 + k, spp, threads=CPU_THREADS, region=region)
        rays = 0
        secs = 0.0
        for k in range(steps):
            _, info = osc.render(k * spp, spp, threads=CPU_THREADS, region=region)
            rays += _cpu_rays(info)
            secs += info["seconds"]
        value = rays / secs / 1e6
        sample = (f"{region[2] - region[0]}x{region[3] - region[1]} pixel region of the frame, {spp} spp per step, "
                  f"{secs / steps:.1f} s of CPU work per step, all host threads (OpenMP)")
        print(json.dumps({
            "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference renders GPU-only via OptiX (no libnvoptix on the box, build needs cmake): CPU oracle port timed instead",
        }))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist

    from luminary_b200 import api

    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- setup, reported separately (part of time-to-N-spp, not of the per-pass metric) ----
    t_setup = time.perf_counter()
    scene = wl["fn"]()
    t_scene = time.perf_counter()
    dev = api.Device(local_rank)
    dev.build_bsdf_lut()
    dev.sync()
    t_lut = time.perf_counter()
    lt = dev.load_scene(scene, light_tree="auto")  # host C tree builder; luminance-textured emitters are integrated on the device first
    dev.sync()
    t_load = time.perf_counter()
    setup = {"scene_generation_s": t_scene - t_setup, "bsdf_lut_s": t_lut - t_scene, "upload_light_tree_bvh_s": t_load - t_lut}
    if args.no_sort:
        dev.update_settings(scene.width, scene.height, scene.max_ray_depth, sort_by_material=False)
    n_pix = scene.width * scene.height

    # accumulation planes live in a torch tensor so that torch.distributed can reduce them in place
    planes = torch.zeros(4 * n_pix, dtype=torch.float32, device=f"cuda:{local_rank}")
    dev.bind_frame_planes(planes.data_ptr(), planes.numel())
    stream = torch.cuda.ExternalStream(dev.stream(), device=f"cuda:{local_rank}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the exchange step of the path lives behind the C ABI (csrc/comm.cu: ncclReduce of the planes on the render stream);
    # torch.distributed only carries the 128-byte communicator id and the timing scalars
    comm = None
    if world > 1:
        uid = torch.tensor(list(api.Comm.unique_id() if rank == 0 else bytes(api.COMM_ID_BYTES)), dtype=torch.uint8, device=f"cuda:{local_rank}")
        dist.broadcast(uid, src=0)
        comm = api.Comm(dev, world, rank, bytes(uid.cpu().tolist()))

    def reduce_to_rank0():
        """one NCCL sum-reduce of the 4 planes onto rank 0, queued on the render stream (lumb200_comm_reduce_planes)"""
        if comm is not None:
            comm.reduce_planes(0)

    def ray_total(st):
        return st["closest_rays"] + st["shadow_rays"] + st["light_rays"]

    # instrumented pass: BVH / light-tree work per ray for the roofline's algorithmic bytes
    dev.start_render()
    if args.no_measure:
        trav = {k: 0 for k in ("closest_rays", "closest_nodes", "closest_tris", "shadow_rays", "shadow_nodes", "shadow_tris", "light_rays",
                               "shaded_vertices", "light_tree_nodes", "light_root_sections")}
    else:
        trav = dev.measure_traversal(0)

    # ---- warm-up: W passes AND the collective (NCCL sets its channels up lazily on the first call of an op) ----
    dev.start_render()
    for k in range(warmup):
        dev.render_samples((1 << 19) + rank + k * world, 1, 1)
    dev.sync()
    reduce_to_rank0()
    barrier()

    # ---- device-resident timing: K un-instrumented steps ----
    first_id, count, stride = sharding.rank_sample_ids(steps * world, rank, world)  # ids rank, rank + world, ...
    dev.start_render()
    stats0 = dev.stats()
    dev.set_profiling(False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    dev.render_samples(first_id, count, stride)
    reduce_to_rank0()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    stats1 = dev.stats()
    rays = ray_total(stats1) - ray_total(stats0)
    launches = stats1["kernel_launches"] - stats0["kernel_launches"]
    assert stats1["stack_overflows"] == 0, "traversal stack overflow: rays lost subtrees"

    # the reduced planes must be what ONE rank renders for the same sample ids (float addition order aside)
    reduce_check = None
    if world > 1 and not args.no_reduce_check:
        reduced = planes.clone() if rank == 0 else None
        barrier()
        if rank == 0:
            dev.start_render()
            dev.render_samples(0, steps * world, 1)
            dev.sync()
            both_finite = torch.isfinite(planes) & torch.isfinite(reduced)
            num = float(torch.linalg.vector_norm(torch.where(both_finite, planes - reduced, torch.zeros_like(planes)).double()))
            den = float(torch.linalg.vector_norm(torch.where(both_finite, planes, torch.zeros_like(planes)).double()))
            reduce_check = {"rel_l2": num / max(den, 1e-30), "sample_ids": steps * world, "ok": bool(num <= 1e-5 * den),
                            "what": "||reduce over ranks - single-rank render of the same ids|| / ||single-rank render||"}
        barrier()

    # ---- the same K steps again with per-stage CUDA events: kernel times for the roofline, outside the timed region ----
    dev.start_render()
    dev.set_profiling(True)
    dev.render_samples(first_id, count, stride)
    dev.sync()
    prof = dev.profile()
    dev.set_profiling(False)

    t = torch.tensor([ms, float(rays), float(launches)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_all, rays_all, launches_all = float(tmax[0]), float(tsum[1]), int(tsum[2])
    else:
        ms_all, rays_all, launches_all = ms, float(rays), int(launches)
    value = rays_all / (ms_all * 1e-3) / 1e6

    # ---- end to end through the public API with host buffers ----
    dev.start_render()
    host_cam = dict(scene.camera)
    # host destinations of the per-step frame read-back (double-buffered: the copy of step k overlaps the passes of step k + 1)
    pinned = [torch.empty(3 * n_pix, dtype=torch.float32).pin_memory() for _ in range(2)]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st_a = dev.stats()
    e0.record(stream)
    t_wall = time.perf_counter()
    for k in range(steps):
        dev.update_settings(scene.width, scene.height, scene.max_ray_depth, sort_by_material=not args.no_sort)  # host struct -> device
        dev.update_camera(host_cam)
        dev.render_samples(rank + k * world, 1, 1)
        dev.wait_download(k & 1)  # the slot's previous copy (step k - 2) has to have landed before it is overwritten
        dev.download_result_async(k + 1, pinned[k & 1].data_ptr(), k & 1)  # resolve + D2H of the RGB frame into pinned host memory
    dev.wait_download(0)
    dev.wait_download(1)  # every frame is in host memory before the clock stops
    dev.sync()
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    st_b = dev.stats()
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)
    e2e_rays = ray_total(st_b) - ray_total(st_a)
    te = torch.tensor([e2e_ms, float(e2e_rays)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        a = te.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone()
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_ms_all, e2e_rays_all = float(a[0]), float(b[1])
    else:
        e2e_ms_all, e2e_rays_all = e2e_ms, float(e2e_rays)
    e2e_value = e2e_rays_all / (e2e_ms_all * 1e-3) / 1e6
    h2d = 16 + 52  # Lumb200Settings + Lumb200Camera host structs per step
    d2h = 3 * n_pix * 4

    # ---- measured time to N spp: start_render -> N / world passes per rank -> reduce -> resolved frame in host memory ----
    time_to_spp = None
    if args.spp > 0:
        total = args.spp
        f_id, f_count, f_stride = sharding.rank_sample_ids(total, rank, world)
        host_frame = torch.empty(3 * n_pix, dtype=torch.float32).pin_memory()
        barrier()
        t0 = time.perf_counter()
        dev.start_render()
        done = 0
        while done < f_count:  # passes are queued in chunks so that the launch queue never blocks the host for long
            chunk = min(64, f_count - done)
            dev.render_samples(f_id + done * f_stride, chunk, f_stride)
            done += chunk
        reduce_to_rank0()
        if rank == 0:
            dev.download_result_async(total, host_frame.data_ptr(), 0)
            dev.wait_download(0)
        dev.sync()
        barrier()
        wall = time.perf_counter() - t0
        tt = torch.tensor([wall], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        time_to_spp = {"spp": total, "seconds": float(tt[0]), "passes_per_rank": f_count, "includes": "start_render, all sample passes, NCCL reduce, "
                       "resolve + D2H of the RGB frame", "excludes": "scene generation / upload / BVH build / LUTs (see setup_s)",
                       "frame_mean": float(host_frame.mean()) if rank == 0 else None}

    if rank == 0:
        # roofline of the dominant stage, chosen by its measured share of the step
        peak, peak_src = measured_peak_gbs()
        total_prof_ms = sum(v["ms"] for v in prof.values())
        share = {k: (v["ms"] / total_prof_ms if total_prof_ms else 0.0) for k, v in prof.items()}
        sections = trav["light_root_sections"]
        nee_bytes = (16.0 + 48.0 * sections + 8.0 * 96.0) * trav["shaded_vertices"] + 64.0 * trav["light_tree_nodes"] if lt is not None else 0.0
        stage_bytes = {  # algorithmic bytes of ONE sample pass, no cache credit
            "trace_closest": 48.0 * trav["closest_rays"] + 80.0 * trav["closest_nodes"] + 48.0 * trav["closest_tris"],
            "trace_shadow": 36.0 * trav["shadow_rays"] + 80.0 * trav["shadow_nodes"] + 48.0 * trav["shadow_tris"],
            "shade": 246.0 * trav["shaded_vertices"] + nee_bytes,
        }
        kernel_names = {"trace_closest": "k_trace_closest", "trace_shadow": "k_trace_shadow", "shade": "k_shade"}
        dominant = max(stage_bytes, key=lambda k: share.get(k, 0.0))
        launches_per_pass = scene.max_ray_depth + 1
        stages = {}
        for k, b in stage_bytes.items():
            avg_ms = prof[k]["ms"] / max(prof[k]["launches"], 1)
            ach = (b / launches_per_pass) / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
            tr, tr_src = ncu_traffic(kernel_names[k], args.workload)
            stages[k] = {"kernel": kernel_names[k], "share_of_step": share.get(k, 0.0), "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": b / launches_per_pass,
                         "achieved": ach, "frac": ach / peak, "traffic": tr, "traffic_source": tr_src}
        dk = stages[dominant]
        roofline = {
            "bound": "hbm", "kernel": dk["kernel"], "achieved": dk["achieved"], "peak": peak, "unit": "GB/s", "frac": dk["frac"], "traffic": dk["traffic"],
            "traffic_source": dk["traffic_source"], "peak_source": peak_src,
            "algorithmic_bytes_per_launch": dk["algorithmic_bytes_per_launch"], "avg_launch_ms": dk["avg_launch_ms"],
            "launch": "one wavefront iteration's launch(es) of the stage (k_shade: its material-class kernels together)",
            "per_ray": {"nodes_visited": trav["closest_nodes"] / max(trav["closest_rays"], 1), "tris_tested": trav["closest_tris"] / max(trav["closest_rays"], 1),
                        "shadow_nodes_visited": trav["shadow_nodes"] / max(trav["shadow_rays"], 1),
                        "shadow_tris_tested": trav["shadow_tris"] / max(trav["shadow_rays"], 1)},
            "per_vertex": {"light_root_sections": sections, "light_tree_nodes": trav["light_tree_nodes"] / max(trav["shaded_vertices"], 1),
                           "bytes": stage_bytes["shade"] / max(trav["shaded_vertices"], 1)},
            "share_of_step": share, "stages": stages,
            "note": "kernel times from a separate profiled run of the same steps; scenes that fit the 126 MB L2 are served from it, so the "
                    "no-cache-credit algorithmic rate may exceed what DRAM delivers - `traffic` is the ncu-measured DRAM bytes per launch",
        }
        # CPU baseline on a bounded sample of the same workload (oracle port, all host threads)
        cpu = None  # reported at N = 1 only (the host cores are shared by all ranks at N > 1)
        if world == 1 and not args.no_cpu:
            luts = dev.get_bsdf_lut()
            osc = oracle_scene(scene, luts, lt)
            region, cpu_spp = plan_cpu_sample(osc, scene, args.cpu_seconds)
            cpu_rays, cpu_secs = run_cpu(osc, region, cpu_spp)
            cpu = {"value": cpu_rays / cpu_secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                   "sample": f"{region[2] - region[0]}x{region[3] - region[1]} pixel region, {cpu_spp} spp, {cpu_secs:.1f} s of CPU time, "
                             "OpenMP over all host threads"}
        out = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_all / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "spp_per_s": world * steps / (ms_all * 1e-3), "rays_per_step": rays_all / steps / world,
            "time_to_1024spp_extrapolated_s": 1024.0 / (world * steps / (ms_all * 1e-3)), "time_to_spp": time_to_spp, "setup_s": setup,
            "reduce_check": reduce_check,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_all / steps},
            "gpu_launches": launches_all, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "bvh": {"nodes": stats1["bvh_nodes"], "tris": stats1["bvh_tris"], "build_ms": stats1["accel_build_seconds"] * 1e3,
                    "depth": stats1["bvh_depth"], "sah_cost": stats1["bvh_sah_cost"], "ploc_radius": stats1["bvh_ploc_radius"],
                    "stack_overflows": stats1["stack_overflows"]},
            "nonfinite_samples": int(dev.stats()["nonfinite_samples"]),
            "kernel_ms_per_step": {k: v["ms"] / steps for k, v in prof.items()},
        }
        print(json.dumps(out))
    if comm is not None:
        comm.destroy()
    dev.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
