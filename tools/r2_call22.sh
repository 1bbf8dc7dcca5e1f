#!/bin/bash
# Round 2, GPU visit 22: ncu launch list of the bench command restricted to the per-bounce kernels (the unfiltered -c 800 list of visit 13
# ended inside the BVH build: the SAH-guided PLOC radius search alone launches ~790 kernels)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"^k_(raygen|rng_table|trace_closest|trace_shadow|trace_enum|enum_finish|sort_count|sort_scan|sort_scatter|shade|shade_miss|next_bounce|reset_fetch|accumulate|generate_result|resolve)" \
  -c 900 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-measure > gpurun_out/r2v_ncu_launch_run.log 2>&1
grep -c "k_shade" gpurun_out/r2v_launches.csv; tail -2 gpurun_out/r2v_ncu_launch_run.log | cut -c1-300
