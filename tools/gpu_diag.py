"""Development diagnostic: compare GPU closest hits with the oracle on the small atrium and report where they differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import orc
from luminary_b200 import api, scenes

SKY = 0xFFFFFFFE

def run(sc, label):
    dev = api.Device(0)
    dev.load_scene(sc)
    inst, tri, t, u, v = dev.trace_primary(0)
    osc = orc.OracleScene(sc)
    ref = osc.trace_primary(0)
    bad = (inst != ref["instance"]) | (tri != ref["tri"])
    print(label, "gpu hit%", 100 * np.mean(inst != SKY), "oracle hit%", 100 * np.mean(ref["instance"] != SKY), "id mismatches", bad.sum(), "of", bad.size)
    tb = t.view(np.uint32) != ref["t"].view(np.uint32)
    print("  t bit mismatches", tb.sum(), "u", (u.view(np.uint32) != ref["u"].view(np.uint32)).sum(), "v", (v.view(np.uint32) != ref["v"].view(np.uint32)).sum())
    if tb.sum():
        idx = np.nonzero(tb & ~bad)[0][:5]
        for i in idx:
            print("   ", i, t[i], ref["t"][i], int(t.view(np.uint32)[i]) - int(ref["t"].view(np.uint32)[i]), u[i], ref["u"][i], v[i], ref["v"][i])
    # rays: feed oracle rays to the GPU to isolate the camera from the traversal
    o, d = osc.camera_rays(0) if sc.width * sc.height <= 200000 else (None, None)
    if o is not None:
        i2, t2, tt2, u2, v2 = dev.trace_rays(o, d)
        r2 = osc.trace_rays(o, d)
        offs = np.cumsum([0] + [sc.meshes[i.mesh_id].num_tris for i in sc.instances])
        hit = r2["prim"] != SKY
        ri = np.searchsorted(offs, r2["prim"][hit], side="right") - 1
        rt = r2["prim"][hit] - offs[ri]
        bad2 = np.zeros(o.shape[0], bool)
        bad2[hit] = (i2[hit] != ri) | (t2[hit] != rt)
        bad2[~hit] = i2[~hit] != SKY
        print("  explicit rays: id mismatches", bad2.sum(), "t bit mismatches", (tt2.view(np.uint32) != r2["t"].view(np.uint32)).sum())
        if bad2.sum():
            j = np.nonzero(bad2)[0]
            print("   mismatch dirs sign pattern:", np.unique((d[j] > 0).astype(int) @ np.array([1, 2, 4]), return_counts=True))
            print("   all dirs sign pattern:", np.unique((d > 0).astype(int) @ np.array([1, 2, 4]), return_counts=True))
            print("   gpu says", i2[j[:8]], t2[j[:8]], tt2[j[:8]], "oracle", r2["prim"][j[:8]], r2["t"][j[:8]])
    st = dev.stats()
    print("  stats", st)
    dev.destroy()

run(scenes.example(width=320, height=180, sphere_subdiv=3), "example-small")
run(scenes.atrium(target_tris=20000, width=320, height=180), "atrium-20k")
sc = scenes.atrium(target_tris=20000, width=320, height=180)
sc.instances = sc.instances[:1]
run(sc, "atrium-static-only")
