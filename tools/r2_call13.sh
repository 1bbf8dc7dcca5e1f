#!/bin/bash
# Round 2, GPU visit 13: whole suite after the color_any fix, non-finite hunt (expect none), default bench, ncu launch list + full captures of the
# three workloads with the final kernels (packed node test, pipelined root loop, sky-capable k_shade)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest.log
tail -5 gpurun_out/r2m_pytest.log
timeout 600 python tools/find_nonfinite.py --workload atrium1m --spp 16 --chunk 8 --oracle 0 > gpurun_out/r2m_nonfinite.log 2>&1; tail -2 gpurun_out/r2m_nonfinite.log
timeout 600 python bench.py --cpu-seconds 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -2 gpurun_out/r2m_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r2m_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("bench:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["kernel_ms_per_step"], "nonfinite", d["nonfinite_samples"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
PY
for wl in divergence terrain10m; do
  timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2m_$wl.json 2> gpurun_out/r2m_$wl.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 800 --csv --log-file gpurun_out/r2m_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-measure > gpurun_out/r2m_ncu_launch_run.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,launch__grid_size,launch__block_size,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
cap() { # workload skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_shade|k_trace_shadow|k_trace_enum" -s $2 -c $3 -f -o /tmp/r2m_full_$1 \
    python bench.py --workload $1 --steps 1 --warmup 1 --no-cpu --no-measure > gpurun_out/r2m_ncu_full_$1.log 2>&1
  ncu -i /tmp/r2m_full_$1.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2m_full_$1_raw.csv 2>> gpurun_out/r2m_ncu_full_$1.log
  ls -la /tmp/r2m_full_$1.ncu-rep
}
cap atrium1m 36 12
cap terrain10m 30 10
cap divergence 54 12
ncu -i /tmp/r2m_full_atrium1m.ncu-rep --page source --csv -k regex:"k_shade" > gpurun_out/r2m_full_atrium1m_shade_source.csv 2>/dev/null
ncu -i /tmp/r2m_full_atrium1m.ncu-rep --page source --csv -k regex:"k_trace_closest" > gpurun_out/r2m_full_atrium1m_closest_source.csv 2>/dev/null
du -sh gpurun_out; ls gpurun_out | grep r2m | tail -30
