#!/bin/bash
# Round 2, GPU visit 42 (final state): whole suite, smoke, headline bench + reference arm, the other two workloads, ncu launch list of the
# per-bounce kernels and ncu --set full captures of the final kernels (64-blocks-per-SM k_shade grid, tiled pixel order, 69 launches per pass)
mkdir -p gpurun_out
T=r2ar
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; grep "^{" gpurun_out/${T}_bench.log > gpurun_out/${T}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 2>> gpurun_out/${T}_bench.err | grep "^{" > gpurun_out/${T}_bench_reference_arm.json
for wl in terrain10m divergence; do
  timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu 2> gpurun_out/${T}_$wl.err | grep "^{" > gpurun_out/${T}_bench_$wl.json
done
python - <<'PY'
import json
for n in ("n1","terrain10m","divergence"):
    d=json.loads(open(f"gpurun_out/r2ar_bench_{n}.json").readline())
    print(n, "%.1f Mrays/s %.3f ms e2e %.1f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]), {k: round(v,3) for k,v in d["kernel_ms_per_step"].items()}, "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"^k_(raygen|rng_table|trace_closest|trace_shadow|trace_enum|enum_finish|sort_count|sort_scan|sort_scatter|shade|shade_miss|next_bounce|accumulate|generate_result|resolve)" \
  -c 900 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-measure > gpurun_out/${T}_ncu_launch_run.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,launch__grid_size,launch__block_size,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
cap() { # workload skip count
  timeout 700 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_shade|k_trace_shadow|k_trace_enum" -s $2 -c $3 -f -o /tmp/${T}_full_$1 \
    python bench.py --workload $1 --steps 1 --warmup 1 --no-cpu --no-measure > gpurun_out/${T}_ncu_full_$1.log 2>&1
  ncu -i /tmp/${T}_full_$1.ncu-rep --page raw --csv --metrics $M > gpurun_out/${T}_full_$1_raw.csv 2>> gpurun_out/${T}_ncu_full_$1.log
}
cap atrium1m 36 12
cap terrain10m 30 10
cap divergence 54 12
du -sh gpurun_out; ls gpurun_out | grep ${T} | tail -30
