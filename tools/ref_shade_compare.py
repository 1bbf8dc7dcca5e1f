"""Reference shading kernel vs product k_shade on the same wavefront, on one B200.

The reference's closest-hit / shadow stages are OptiX programs and cannot run here (no libnvoptix), but its
shading kernel `geometry_process_tasks` (cuda/geometry.cuh:11-180) is plain CUDA: oracle/ref builds it UNMODIFIED from
/root/reference for sm_100a (oracle/_ref/librefdev.so). This tool times it with CUDA events on the path vertices of every
wavefront iteration of one sample pass of a bench workload (vertices produced by the CPU oracle, packed into the reference's
warp-interleaved task records with the reference's own launch geometry), and prints the product's k_shade time for the same pass
beside it. k_shade additionally enumerates the emitter BVH for the BSDF-sampled light (the reference does that in its OptiX shadow
stage), so the comparison is conservative for the product.

Test / measurement infrastructure only (imports tests/ helpers and oracle/); prints one JSON line, also written to
gpurun_out/ref_shade_compare.json.

    python tools/ref_shade_compare.py [--workload atrium1m] [--sample-id 0] [--iterations 6]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="atrium1m")
    ap.add_argument("--sample-id", type=int, default=0)
    ap.add_argument("--iterations", type=int, default=0, help="wavefront iterations to compare (0 = max_ray_depth + 1)")
    ap.add_argument("--repeats", type=int, default=5)
    args = ap.parse_args()

    import bench
    import orc
    import refdev
    from luminary_b200 import api

    if not refdev.available():
        print(json.dumps({"unavailable": "oracle/_ref/librefdev.so not built (needs /root/reference at build time)"}))
        return
    scene = bench.WORKLOADS[args.workload]["fn"]()
    iters = args.iterations or scene.max_ray_depth + 1

    # ---- reference device: its own packers, its own light tree, its own LUT kernels
    ref = refdev.RefDevice(scene)
    ref_luts = ref.build_bsdf_lut()
    lt = ref.light_tree

    # ---- product: one profiled pass of the same sample id with the SAME light tree
    dev = api.Device(0)
    dev.build_bsdf_lut()
    dev.load_scene(scene, light_tree=lt[:3])
    dev.start_render()
    for k in range(3):
        dev.render_samples(1000 + k, 1, 1)
    dev.sync()
    dev.start_render()
    s0 = dev.stats()
    dev.set_profiling(True)
    dev.render_samples(args.sample_id, 1, 1)
    dev.sync()
    prof = dev.profile()
    dev.set_profiling(False)
    s1 = dev.stats()
    ours_ms = prof["shade"]["ms"]
    ours_launches = prof["shade"]["launches"]
    ours_vertices = s1["closest_rays"] - s0["closest_rays"]  # every traced path is shaded (hit or miss) once
    dev.destroy()

    # ---- path vertices of each iteration from the CPU oracle, shaded by the reference kernel
    osc = orc.OracleScene(scene)
    osc.set_light_tree(*lt[:3])
    osc.set_bsdf_luts(*ref_luts)
    handles = osc.prim_handles()
    # the reference's launch geometry (device.c:422-488): 148 SMs x 8 blocks x 256 threads on B200
    T_blocks = 148 * 2048 // refdev.THREADS_PER_BLOCK
    T = T_blocks * refdev.THREADS_PER_BLOCK
    per_iter = []
    ref_total = 0.0
    ref_vertices = 0
    for it in range(iters):
        t0 = time.perf_counter()
        vin, _pix = osc.path_vertices(args.sample_id, it)
        cpu_s = time.perf_counter() - t0
        n = int(vin.size)
        if n == 0:
            break
        depth = it if not (it == scene.max_ray_depth and it > 0) else it - 1  # device_renderer.c:126-130 quirk
        ref.configure(T_blocks, -(-n // T))
        ms = ref.time_shade(refdev.tasks_from_vertices(vin, handles), depth, args.repeats)
        per_iter.append({"iteration": it, "vertices": n, "ref_ms": ms, "oracle_cpu_s": cpu_s})
        ref_total += ms
        ref_vertices += n
        print(f"iteration {it}: {n} geometry vertices, reference geometry_process_tasks {ms:.3f} ms", file=sys.stderr, flush=True)

    out = {
        "workload": bench.WORKLOADS[args.workload]["desc"], "sample_id": args.sample_id,
        "reference_kernel": "geometry_process_tasks (cuda/geometry.cuh:11-180), unmodified, nvcc 12.9 sm_100a --use_fast_math",
        "reference_ms_per_pass": ref_total, "reference_geometry_vertices": ref_vertices,
        "reference_ns_per_vertex": 1e6 * ref_total / max(ref_vertices, 1),
        "product_kernel": "k_shade (geometry + miss shading + emitter-BVH light enumeration)",
        "product_ms_per_pass": ours_ms, "product_launches": ours_launches, "product_vertices_incl_misses": ours_vertices,
        "product_ns_per_vertex": 1e6 * ours_ms / max(ours_vertices, 1),
        "speedup_shade_stage": ref_total / ours_ms if ours_ms > 0 else None,
        "per_iteration": per_iter,
    }
    line = json.dumps(out)
    print(line)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_shade_compare.json"), "w") as f:
        f.write(line + "\n")


if __name__ == "__main__":
    main()
