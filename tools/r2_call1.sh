#!/bin/bash
# Round 2, GPU visit 1: the whole GPU suite incl. the new per-vertex / per-ray / BASELINE-config parity tests on the round-1 kernels,
# then the PLOC-radius sweep (SAH-predicted vs measured node visits) and the L2 persisting-window A/B.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]; r=d["roofline"]["per_ray"]; b=d["bvh"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} nodes/ray {r["nodes_visited"]:.2f} tris/ray {r["tris_tested"]:.2f} bvh nodes {b["nodes"]} depth {b["depth"]} sah {b["sah_cost"]:.3f} radius {b["ploc_radius"]} build {b["build_ms"]:.1f} ms ovf {b["stack_overflows"]}')
PY
}
for wl in atrium1m terrain10m; do
  rs="8 16 32 64 auto"; [ $wl = terrain10m ] && rs="16 32 64 auto"
  for r in $rs; do
    LUMB200_BVH_VERBOSE=1 LUMB200_PLOC_RADIUS=$r timeout 600 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu > gpurun_out/r2a_${wl}_r$r.json 2> gpurun_out/r2a_${wl}_r$r.err
    echo "$wl radius $r: $(line gpurun_out/r2a_${wl}_r$r.json)" | tee -a gpurun_out/r2a_sweep.txt
    grep "PLOC radius" gpurun_out/r2a_${wl}_r$r.err | head -5 >> gpurun_out/r2a_sweep.txt
  done
  LUMB200_NO_L2_WINDOW=1 timeout 600 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu > gpurun_out/r2a_${wl}_nol2.json 2> gpurun_out/r2a_${wl}_nol2.err
  echo "$wl no L2 window: $(line gpurun_out/r2a_${wl}_nol2.json)" | tee -a gpurun_out/r2a_sweep.txt
done
