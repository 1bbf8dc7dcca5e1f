"""Generates tests/golden/config1_hits.json: digests of the closest-hit records of BASELINE config 1 (Example scene,
960x540 primary rays of sample 0) as computed by the CPU oracle, plus 64 spot values.

The reference cannot run here (GPU-only, closed-source OptiX traversal) and ships no golden vectors, so the fixture
pins the ORACLE: tests/test_oracle_core.py re-derives the digests on the CPU (a change of the oracle's arithmetic
shows up immediately) and tests/test_trace_gpu.py requires the CUDA path to hit the same digests through the C ABI.
The oracle's BVH2 traversal is cross-checked against its own brute-force loop over all triangles on a strided subset
of the rays before anything is written. Run from the repo root: python tests/golden/make_config1_hits.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import orc  # noqa: E402
from luminary_b200 import scenes  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    scene = scenes.example()
    osc = orc.OracleScene(scene)
    ref = osc.trace_primary(0)
    # brute-force cross-check on every 997th ray: same primitive, bit-identical t from both traversal strategies
    import ctypes as C

    L = orc.lib()
    w = scene.width
    idx = np.arange(0, ref["tri"].size, 997)
    o = np.empty((idx.size, 3), np.float32)
    d = np.empty((idx.size, 3), np.float32)
    oo, dd = orc.Vec3(), orc.Vec3()
    for k, i in enumerate(idx):
        L.orc_camera_sample(C.byref(osc.camera), C.byref(osc.settings), L.orc_path_id_get(int(i % w), int(i // w), 0), C.byref(oo), C.byref(dd))
        o[k] = (oo.x, oo.y, oo.z)
        d[k] = (dd.x, dd.y, dd.z)
    bvh = osc.trace_rays(o, d)
    for k, i in enumerate(idx):
        h = L.orc_closest_hit_bruteforce(osc.handle, orc.vec3(o[k]), orc.vec3(d[k]), 0.0, 3.402823466e38, 0xFFFFFFFF, 0)
        assert h.prim == bvh["prim"][k], (i, h.prim, bvh["prim"][k])
        assert np.float32(h.t).view(np.uint32) == bvh["t"][k].view(np.uint32) == ref["t"][i].view(np.uint32)
    spots = [int(i) for i in np.linspace(0, ref["tri"].size - 1, 64).astype(np.int64)]
    out = {
        "scene": "scenes.example() seed 0xB200E0, 40972 triangles, 960x540, sample 0",
        "count": int(ref["tri"].size),
        "sha256": {k: digest(ref[k]) for k in ("instance", "tri")} | {k: digest(ref[k].view(np.uint32)) for k in ("t", "u", "v")},
        "spots": [[i, int(ref["instance"][i]), int(ref["tri"][i]), int(ref["t"][i].view(np.uint32))] for i in spots],
    }
    with open(os.path.join(HERE, "config1_hits.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["sha256"], indent=1))


if __name__ == "__main__":
    main()
