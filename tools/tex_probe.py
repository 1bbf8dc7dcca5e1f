"""Probes the B200 texture unit's bilinear weight rule: a 2x1 fp32 texture {0, 1} sampled at fine u steps returns the
weight itself. Output: gpurun_out/tex_probe.npz (u, value) for widths 2, 64 and 4096."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from luminary_b200 import api

dev = api.Device(0, load_embedded_data=False)
out = {}
texs = []
for w in (2, 64, 4096):
    row = np.zeros((1, w, 1), np.float32)
    row[0, 1::2, 0] = 1.0
    texs.append(dict(data=row, wrap_u=1, wrap_v=1, filter=1, gamma=1.0))
# a 2x2 for the 2D rule
texs.append(dict(data=np.array([[[0.0], [1.0]], [[2.0], [4.0]]], np.float32), wrap_u=1, wrap_v=1, filter=1, gamma=1.0))
# u8 texture to see unorm conversion + interpolation precision
texs.append(dict(data=np.array([[[0], [255]]], np.uint8), wrap_u=1, wrap_v=1, filter=1, gamma=1.0))
texs.append(dict(data=np.array([[[51], [102]]], np.uint8), wrap_u=1, wrap_v=1, filter=1, gamma=1.0))
dev.add_textures(texs)
for k, w in enumerate((2, 64, 4096)):
    # between texel 0 and texel 1 centres: u in [0.5/w, 1.5/w]
    n = 1 << 16
    xb = np.arange(n + 1, dtype=np.float64) / n  # xB in [0, 1]
    u = ((xb + 0.5) / w).astype(np.float32)
    val = dev.sample_texture(k, np.stack([u, np.full_like(u, 0.5)], axis=1))[:, 0]
    out[f"u_{w}"] = u
    out[f"v_{w}"] = val
rng = np.random.default_rng(1)
uv = rng.random((20000, 2)).astype(np.float32)
out["uv2d"] = uv
out["v2d"] = dev.sample_texture(3, uv)[:, 0]
u = ((np.arange(4097, dtype=np.float64) / 4096 + 0.5) / 2).astype(np.float32)
uvl = np.stack([u, np.full_like(u, 0.5)], axis=1)
out["u_u8"] = u
out["v_u8a"] = dev.sample_texture(4, uvl)[:, 0]
out["v_u8b"] = dev.sample_texture(5, uvl)[:, 0]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "tex_probe.npz"), **out)
print("ok")
