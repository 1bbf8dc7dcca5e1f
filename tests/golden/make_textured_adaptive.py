"""Generates tests/golden/textured_adaptive.json: digests that pin the oracle's restatement of the two SURVEY 8(f) rows built in
round 1 - (1) closest hits with alpha cut-outs on scenes.textured_example (384 x 216, sample 3), (2) the adaptive sampler's stage
sample counts after 2 + 4 executions on scenes.example_with_light (48 x 28, interval 2) and the planes' checksum.

The reference cannot run here and ships no golden vectors; the fixture pins the ORACLE (tests/test_texture_oracle.py and
tests/test_adaptive_oracle.py re-derive it on the CPU) and the CUDA path is checked against the same digests on the B200
(tests/test_texture_gpu.py). The texel fetch itself is pinned against values measured on the B200 texture unit
(tests/golden/tex_probe_b200.npz, made by tools/tex_probe.py). Run from the repo root: python tests/golden/make_textured_adaptive.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import orc  # noqa: E402
from luminary_b200 import api, scenes  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def textured_hits():
    scene = scenes.textured_example(width=384, height=216)
    ref = orc.OracleScene(scene).trace_primary(3)
    return {"scene": "scenes.textured_example(384, 216) seed 0xB200F1, sample 3", "count": int(ref["tri"].size),
            "sha256": {k: digest(ref[k]) for k in ("instance", "tri")} | {"t": digest(ref["t"].view(np.uint32))},
            "screen_hits": int((ref["instance"] == 1).sum())}


def adaptive_words():
    sc = scenes.example_with_light(width=48, height=28, sphere_subdiv=2, max_ray_depth=2)
    osc = orc.OracleScene(sc)
    osc.set_light_tree(*api.build_light_tree(sc))
    p = orc.adaptive_params(max_sampling_rate=8, avg_sampling_rate=2, update_interval=2, exposure=1.0, tonemap=4)
    st = osc.render_adaptive(p, 2 + 4, threads=1)
    return {"scene": "scenes.example_with_light(48, 28, subdiv 2, depth 2), no LUTs set (albedo terms = 1), max 8 / avg 2 / interval 2 / AgX",
            "executions": st["executions"], "stage": st["stage"], "paths": st["paths"], "words_sha256": digest(st["words"]),
            "count_histogram_stage1": np.bincount(((st["words"] & 0xFF) + 1).reshape(-1), minlength=9).tolist(),
            "count_histogram_stage2": np.bincount((((st["words"] >> 8) & 0xFF) + 1).reshape(-1), minlength=9).tolist()}


if __name__ == "__main__":
    out = {"textured_hits": textured_hits(), "adaptive": adaptive_words()}
    with open(os.path.join(HERE, "textured_adaptive.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1)[:1200])
