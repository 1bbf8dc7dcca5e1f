"""Throughput of adaptive-sampling executions next to uniform passes on the headline scene (atrium-1M, 1080p, 5 bounces).
Prints one JSON line; run on the GPU box: python tools/adaptive_bench.py [--interval 8] [--avg 2] [--max 256]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from luminary_b200 import api, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--interval", type=int, default=8)
ap.add_argument("--avg", type=int, default=2)
ap.add_argument("--max", type=int, default=256)
ap.add_argument("--tris", type=int, default=1_000_000)
args = ap.parse_args()

scene = scenes.atrium(args.tris, 1920, 1080, 5)
dev = api.Device(0)
dev.build_bsdf_lut()
dev.load_scene(scene, light_tree="auto")


def rays(st):
    return st["closest_rays"] + st["shadow_rays"] + st["light_rays"]


out = {"scene": scene.name, "interval": args.interval, "avg_sampling_rate": args.avg, "max_sampling_rate": args.max}
# uniform reference
dev.start_render()
dev.render_samples(0, 4)
dev.sync()
dev.start_render()
dev.render_samples(0, 8)
dev.sync()
st = dev.stats()
out["uniform"] = {"ms_per_pass": 1e3 * st["render_seconds"] / 8, "mrays_s": rays(st) / st["render_seconds"] / 1e6}
# adaptive: stage 0 (uniform, table sampler) then stage 1 (adaptive task creation, per-call Sobol sampler, atomic accumulation)
dev.update_adaptive_sampling(max_sampling_rate=args.max, avg_sampling_rate=args.avg, update_interval=args.interval, exposure_aware=True, exposure=1.0,
                             tonemap=4)
dev.start_render()
dev.render_executions(args.interval)
dev.sync()
s0 = dev.stats()
a0 = dev.adaptive_state()
n1 = min(4, 2 * args.interval - 1)
dev.render_executions(n1)
dev.sync()
s1 = dev.stats()
a1 = dev.adaptive_state()
words = dev.adaptive_words()
c = (words & 0xFF) + 1
dt = s1["render_seconds"] - s0["render_seconds"]
out["stage0"] = {"ms_per_execution": 1e3 * s0["render_seconds"] / args.interval, "mrays_s": rays(s0) / s0["render_seconds"] / 1e6}
out["stage1"] = {"executions": n1, "tasks_per_execution": a1["tasks_per_execution"], "ms_per_execution": 1e3 * dt / n1,
                 "mrays_s": (rays(s1) - rays(s0)) / dt / 1e6, "counts_mean": float(c.mean()), "counts_max": int(c.max()),
                 "counts_hist": {str(k): int((c == k).sum()) for k in sorted(set(c.reshape(-1).tolist()))[:12]}}
assert a0["stage_id"] == 1 and a1["stage_id"] == 1
print(json.dumps(out))
dev.destroy()
