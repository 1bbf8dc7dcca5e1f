#!/bin/bash
# Round 2, GPU visit 25: tree rotations on the PLOC hierarchy before the collapse (LUMB200_BVH_ROTATIONS = passes)
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]; r=d["roofline"]["per_ray"]; b=d["bvh"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} nodes/ray {r["nodes_visited"]:.2f} tris/ray {r["tris_tested"]:.2f} shadow nodes {r["shadow_nodes_visited"]:.2f} SAH {b["sah_cost"]:.3f} radius {b["ploc_radius"]} depth {b["depth"]} build {b["build_ms"]:.1f} ms ovf {b["stack_overflows"]}')
PY
}
for wl in atrium1m terrain10m divergence; do
for rot in 0 1 2 4; do
  LUMB200_BVH_ROTATIONS=$rot timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2y_tmp.json 2> gpurun_out/r2y_tmp.err
  echo "$wl rotations $rot: $(line gpurun_out/r2y_tmp.json)" | tee -a gpurun_out/r2y_rotations.txt
  grep -c "rotation pass" gpurun_out/r2y_tmp.err > /dev/null
done
done
LUMB200_BVH_ROTATIONS=2 LUMB200_BVH_VERBOSE=1 timeout 900 python -m pytest tests/test_trace_gpu.py tests/test_configs_gpu.py -q -x 2>&1 | grep -E "passed|failed|rotation pass 0" | sort | uniq -c | tail -5
