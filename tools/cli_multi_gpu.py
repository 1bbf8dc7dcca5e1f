"""The C host's headless front end (LuminaryB200 -b) on all GPUs of the box, on the headline workload:
writes the atrium as one world-space *.obj + *.lum, renders 2^N samples over `--gpus` devices through the public Luminary API
(NCCL reduce behind the C ABI at every output) and prints the front end's own Mrays/s line next to bench.py's number for the same N.

usage: python tools/cli_multi_gpu.py --gpus 8 --log2-samples 10 [--tris 1000000]"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from luminary_b200 import scenes  # noqa: E402

CLI = os.path.join(ROOT, "luminary_b200", "LuminaryB200")


def quat_from_euler(rot):  # host_math.c:6-21 -> (x, y, z, w)
    cr, sr = np.cos(rot[0] * 0.5), np.sin(rot[0] * 0.5)
    cp, sp = np.cos(rot[1] * 0.5), np.sin(rot[1] * 0.5)
    cy, sy = np.cos(rot[2] * 0.5), np.sin(rot[2] * 0.5)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])


def rotate(q, v):
    u, w = q[:3], q[3]
    t = 2.0 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


def bake_world(sc):
    """all active instances as ONE world-space mesh (the *.obj loader of the host delivers one mesh and one identity instance)"""
    vs, ns, ts, ms = [], [], [], []
    for ins in sc.instances:
        if not ins.active:
            continue
        m = sc.meshes[ins.mesh_id]
        q = quat_from_euler(np.asarray(ins.rotation, np.float64))
        s = np.asarray(ins.scale, np.float64)
        v = rotate(q, m.vertex.reshape(-1, 3).astype(np.float64) * s) + np.asarray(ins.translation, np.float64)
        n = rotate(q, m.normal.reshape(-1, 3).astype(np.float64) / s)
        n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
        vs.append(v.reshape(-1, 3, 3)), ns.append(n.reshape(-1, 3, 3)), ts.append(m.uv), ms.append(m.material)
    mesh = scenes.Mesh(np.concatenate(vs).astype(np.float32), np.concatenate(ns).astype(np.float32), np.concatenate(ts), np.concatenate(ms))
    out = scenes.Scene(sc.name + "_baked", [mesh], [scenes.Instance(0)], sc.materials, sc.camera, sc.width, sc.height, sc.max_ray_depth,
                       sc.sky_mode, sc.sky_color)
    return out


def write_lum(path, sc, obj_name):
    c = sc.camera
    with open(path, "w") as f:
        f.write("Luminary\nVERSION 4\n")
        f.write(f"GENERAL WIDTH___ {sc.width}\nGENERAL HEIGHT__ {sc.height}\nGENERAL BOUNCES_ {sc.max_ray_depth}\nGENERAL MESHFILE {obj_name}\n")
        f.write("CAMERA POSITION %.9g %.9g %.9g\nCAMERA ROTATION %.9g %.9g %.9g\n" % (tuple(c["pos"]) + tuple(c["rotation"])))
        f.write("CAMERA FOV_____ %.9g\nCAMERA FOCALLEN %.9g\nCAMERA APERTURE %.9g\n" % (c["fov"], c["object_distance"], c["aperture_size"]))
        f.write("CAMERA EXPOSURE 1\nCAMERA TONEMAP_ 0\nCAMERA DITHER__ 0\nCAMERA PURKINJE 0\nCAMERA BLOOMBLE 0\nCAMERA RUSSIANR %.9g\n" % c["russian_roulette_threshold"])
        f.write("SKY MODE____ %d\nSKY COLORCON %.9g %.9g %.9g\n" % ((sc.sky_mode,) + tuple(sc.sky_color)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--log2-samples", type=int, default=8)
    ap.add_argument("--tris", type=int, default=1_000_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cli_multi_gpu.json"))
    args = ap.parse_args()
    t0 = time.time()
    sc = bake_world(scenes.atrium(target_tris=args.tris))
    tmp = tempfile.mkdtemp(prefix="lumb200_cli_")
    scenes.write_obj(sc, os.path.join(tmp, "atrium.obj"))
    write_lum(os.path.join(tmp, "atrium.lum"), sc, "atrium.obj")
    print(f"scene files written in {time.time() - t0:.1f} s ({sc.num_tris} triangles)", flush=True)
    outdir = os.path.join(tmp, "out")
    os.makedirs(outdir)
    cmd = [CLI, os.path.join(tmp, "atrium.lum"), "-b", str(args.log2_samples), "cli", "-o", outdir, "--supersampling", "0", "--adaptive", "0"]
    for g in range(args.gpus):
        cmd += ["--device", str(g)]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=3000)
    wall = time.time() - t0
    print(r.stdout[-1500:])
    print(r.stderr[-1500:], file=sys.stderr)
    m = re.search(r"(\d+) rays in ([0-9.]+) GPU seconds: ([0-9.]+) Mrays/s, ([0-9.]+) samples/s", r.stdout)
    res = dict(gpus=args.gpus, samples=1 << args.log2_samples, returncode=r.returncode, wall_s=wall)
    if m:
        res.update(rays=int(m.group(1)), gpu_seconds=float(m.group(2)), mrays_per_s=float(m.group(3)), samples_per_s=float(m.group(4)))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
