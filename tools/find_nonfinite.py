"""Hunts the path samples whose radiance is NaN / Inf (Lumb200Stats.nonfinite_samples): renders sample ids one chunk at a time, isolates the
sample id and pixel of every event, then lets the CPU oracle shade the same (pixel, sample id) to tell a product bug from behaviour the
reference arithmetic has too. usage: python tools/find_nonfinite.py [--workload atrium4k] [--spp 256] -> gpurun_out/nonfinite.json"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from luminary_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="atrium4k")
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=8)
    ap.add_argument("--oracle", type=int, default=1)
    args = ap.parse_args()
    sc = bench.WORKLOADS[args.workload]["fn"]()
    lt = api.build_light_tree(sc)
    dev = api.Device(0)
    dev.build_bsdf_lut()
    dev.load_scene(sc, light_tree=lt)
    events = []
    for first in range(0, args.spp, args.chunk):
        dev.start_render()
        dev.render_samples(first, args.chunk)
        if dev.stats()["nonfinite_samples"] == 0:
            continue
        for s in range(first, first + args.chunk):
            dev.start_render()
            dev.render_samples(s, 1)
            st = dev.stats()
            if st["nonfinite_samples"]:
                pix = int(st["nonfinite_pixel"])
                events.append(dict(sample_id=s, count=int(st["nonfinite_samples"]), x=pix % sc.width, y=pix // sc.width))
                print("non-finite sample:", events[-1], flush=True)
    print(f"{len(events)} sample passes with non-finite radiance in {args.spp} spp of {args.workload} ({sc.width}x{sc.height})")
    if events and args.oracle:
        import orc
        from test_shade_vertices_gpu import product_vertices
        osc = orc.OracleScene(sc)
        osc.set_light_tree(*lt)
        osc.set_bsdf_luts(*dev.get_bsdf_lut())
        for e in events[:args.oracle]:
            img, info = osc.render(e["sample_id"], 1, region=(e["x"], e["y"], e["x"] + 1, e["y"] + 1))
            e["oracle_rgb"] = [float(img[c, e["y"], e["x"]]) for c in range(3)]
            print("oracle at", e["x"], e["y"], "sample", e["sample_id"], "->", e["oracle_rgb"], flush=True)
            # walk the oracle's path of that pixel and shade every vertex with the product: which output turns non-finite first?
            for it in range(sc.max_ray_depth + 1):
                vin, _ = osc.path_vertices(e["sample_id"], it)
                sel = ((vin["path_id"][:, 0] & 0x3FFF) == e["x"]) & ((vin["path_id"][:, 1] & 0x3FFF) == e["y"])
                if not sel.any():
                    print("  iteration", it, ": the oracle's path has ended")
                    break
                v = vin[sel]
                depth = it if not (it == sc.max_ray_depth and it > 0) else it - 1
                got = dev.shade_vertices(product_vertices(v), e["sample_id"], depth, it == sc.max_ray_depth)
                want = osc.shade_vertices(v, depth)
                seg = osc.nee_segments(v, depth)
                g = got[0]
                rec = {k: g[k].tolist() for k in ("emission", "alive", "state", "origin", "ray", "record")}
                rec["nee"] = [dict(valid=int(n["valid"]), color=n["color"].tolist(), visible=n["visible"].tolist(), dist=float(n["dist"]),
                                   target=int(n["target_prim"])) for n in g["nee"]]
                rec["oracle_nee"] = [dict(valid=int(n["valid"]), color=n["color"].tolist(), vis=n["visibility"].tolist()) for n in seg[0]]
                rec["vin"] = {k: v[0][k].tolist() for k in ("state", "origin", "ray", "prim", "t", "record")}
                rec["oracle_emission"] = want[0]["emission"].tolist()
                e.setdefault("vertices", []).append(rec)
                print("  iteration", it, json.dumps(rec), flush=True)
    dev.destroy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(workload=args.workload, spp=args.spp, events=events), open(os.path.join(ROOT, "gpurun_out", "nonfinite.json"), "w"))


if __name__ == "__main__":
    main()
