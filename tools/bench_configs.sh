#!/bin/bash
# Runs the BASELINE.json configurations other than the headline one and prints one summary line each.
summ() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['kernel_ms_per_step']
        print('$1: %.1f Mrays/s  %.2f ms/pass  %.1f spp/s  rays/pass %.0f  closest %.2f shadow %.2f shade %.2f sort %.2f  nodes/ray %.1f tris/ray %.1f bvh_ms %.0f' % (d['value'], d['ms_per_step'], d['spp_per_s'], d['rays_per_step'], k['trace_closest'], k['trace_shadow'], k['shade'], k['sort'], d['roofline']['per_ray']['nodes_visited'], d['roofline']['per_ray']['tris_tested'], d['bvh']['build_ms']))
"; }
python bench.py --workload divergence --steps 8 --cpu-seconds 1 2>&1 | summ "config4 divergence sorted"
python bench.py --workload divergence --steps 8 --cpu-seconds 1 --no-sort 2>&1 | summ "config4 divergence unsorted"
python bench.py --workload atrium4k --steps 4 --cpu-seconds 1 2>&1 | summ "config5 atrium 4K (1 GPU)"
python bench.py --workload terrain10m --steps 4 --cpu-seconds 1 2>&1 | tee gpurun_out/terrain.log | summ "config3 terrain 10M"
tail -3 gpurun_out/terrain.log | cut -c1-600
