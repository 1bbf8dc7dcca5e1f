#!/bin/bash
# Sweeps the BVH build knobs (collapse mode, SAH triangle cost, PLOC radius) on a bench workload; one line per variant.
# usage (under gpurun): bash tools/sweep_bvh.sh [workload]
wl=${1:-atrium1m}
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py --workload $wl --steps 8 --warmup 3 --cpu-seconds 0.3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; r=d['roofline']['per_ray']
print('$label: %.1f Mrays/s ms=%.3f closest=%.3f shadow=%.3f shade=%.3f nodes/ray=%.2f tris/ray=%.2f bvh_nodes=%d build_ms=%.1f'%(d['value'],d['ms_per_step'],k['trace_closest'],k['trace_shadow'],k['shade'],r['nodes_visited'],r['tris_tested'],d['bvh']['nodes'],d['bvh']['build_ms']))"
}
run "greedy r16" LUMB200_COLLAPSE=greedy
for c in ${CPRIM:-0.3 0.5 0.8 1.2}; do run "dp cprim=$c r16" LUMB200_SAH_CPRIM=$c; done
for r in ${RADIUS:-32 64}; do run "dp cprim=0.5 r$r" LUMB200_PLOC_RADIUS=$r; done
