"""Bit-exactness of the closest hit on the terrain scene (tiny nodes far from the ray origins: the stress case for the quantisation
margin of the BVH8 child boxes) against the oracle: primary rays + random long rays. usage: python tools/parity_terrain.py [grid]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
from luminary_b200 import api, scenes  # noqa: E402

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
scene = scenes.terrain(grid, 5000, 480, 270, 2)
dev = api.Device(0)
dev.load_scene(scene)
osc = orc.OracleScene(scene)
inst, tri, t, u, v = dev.trace_primary(0)
ref = osc.trace_primary(0)
same = (inst == ref["instance"]) & (tri == ref["tri"])
print("primary: ids equal", float(same.mean()), "t bit-identical", bool(np.array_equal(t.view(np.uint32), ref["t"].view(np.uint32))), "hits", int((inst != 0xFFFFFFFE).sum()))
rng = np.random.default_rng(5)
n = 400_000
o = np.stack([rng.uniform(-500, 500, n), rng.uniform(5, 120, n), rng.uniform(-500, 500, n)], axis=1).astype(np.float32)
d = rng.normal(size=(n, 3)).astype(np.float32)
d[:, 1] = -np.abs(d[:, 1]) * 0.3
d /= np.linalg.norm(d, axis=1, keepdims=True)
gi, gt_, gtt, gu, gv = dev.trace_rays(o, d.astype(np.float32))
r = osc.trace_rays(o, d.astype(np.float32))
handles = np.stack([gi, gt_], axis=1)
prim_inst = np.array([0xFFFFFFFE], np.uint32)
ok_t = np.array_equal(gtt.view(np.uint32), r["t"].view(np.uint32))
hit = r["prim"] != 0xFFFFFFFE
print("random rays:", n, "hits", int(hit.sum()), "t bit-identical", ok_t, "misses agree", bool(np.array_equal(gi == 0xFFFFFFFE, ~hit)))
assert same.all() and ok_t
print("OK")
dev.destroy()
