#!/bin/bash
# Round 2, GPU visit 28: ncu --set full of the sky kernels (LUT builders, HDRI bake, ray-marched miss shader, sun-enabled k_shade) on divergence_sky / _hdri
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size"
timeout 900 ncu --set full --clock-control none -k regex:"k_sky_|k_shade_miss_sky|k_shade" -c 24 -f -o /tmp/r2ab_sky \
  python bench.py --workload divergence_sky --steps 1 --warmup 1 --no-cpu --no-measure > gpurun_out/r2ab_ncu_sky.log 2>&1
ncu -i /tmp/r2ab_sky.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2ab_sky_raw.csv 2>> gpurun_out/r2ab_ncu_sky.log
timeout 900 ncu --set full --clock-control none -k regex:"k_sky_hdri" -c 2 -f -o /tmp/r2ab_hdri \
  python bench.py --workload divergence_hdri --steps 1 --warmup 1 --no-cpu --no-measure > gpurun_out/r2ab_ncu_hdri.log 2>&1
ncu -i /tmp/r2ab_hdri.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2ab_hdri_raw.csv 2>> gpurun_out/r2ab_ncu_hdri.log
wc -l gpurun_out/r2ab_sky_raw.csv gpurun_out/r2ab_hdri_raw.csv
