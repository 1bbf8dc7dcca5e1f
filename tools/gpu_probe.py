"""Quick GPU probe used during development: BVH build + primary-ray throughput on the procedural scenes."""
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from luminary_b200 import api, scenes


def probe(scene, repeats=10):
    t0 = time.time()
    dev = api.Device(0)
    dev.load_scene(scene)
    t1 = time.time()
    ms = dev.time_primary_trace(0, repeats)
    st = dev.stats()
    n = scene.width * scene.height
    inst, tri, t, u, v = dev.trace_primary(0)
    print(f"{scene.name}: tris={st['bvh_tris']} nodes={st['bvh_nodes']} build={st['accel_build_seconds']*1e3:.1f} ms upload+build wall={t1-t0:.2f}s "
          f"primary {ms:.3f} ms -> {n/ms/1e3:.1f} Mrays/s  hit%={100*np.mean(inst!=0xFFFFFFFE):.1f} bytes={st['device_bytes']/1e6:.0f}MB", flush=True)
    dev.destroy()


if __name__ == "__main__":
    probe(scenes.example())
    probe(scenes.atrium(target_tris=100_000))
    t0 = time.time()
    sc = scenes.atrium()
    print("atrium gen", time.time() - t0, "s", sc.num_tris)
    probe(sc)
