import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import orc
from luminary_b200 import api, scenes
sc = scenes.atrium(target_tris=20000, width=320, height=180)
sc.instances = sc.instances[:1]
dev = api.Device(0)
dev.load_scene(sc)
nodes, tris = dev.download_bvh(0)
inst, tri, t, u, v = dev.trace_primary(0)
osc = orc.OracleScene(sc)
o, d = osc.camera_rays(0)
ref = osc.trace_rays(o, d)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/bvh_dump.npz", nodes=nodes, tris=tris, o=o, d=d, gpu_inst=inst, gpu_tri=tri, gpu_t=t, ref_prim=ref["prim"], ref_t=ref["t"], world=osc.world_tris())
print("dumped", nodes.shape, tris.shape)
