#!/bin/bash
# Round 2, GPU visit 35: A/B of the hit-mask assembly of lb_node_hits (LB_HITMASK_V2, stock = 1, variant hm0 = 0) on two workloads,
# then the whole suite, smoke, the headline bench, the reference arm, the launch list and ncu --set full captures with the final kernels
# (the r2m captures predate the reinsertion pass of the builder).
mkdir -p gpurun_out
T=r2ai
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} enum {k.get("trace_enum",0):.3f} ovf {d["bvh"]["stack_overflows"]}')
PY
}
for rep in 1 2; do
for wl in atrium1m terrain10m; do
[ $rep = 2 ] && [ $wl = terrain10m ] && continue
for v in stock hm0; do
    lib=$PWD/luminary_b200/liblumb200.so; [ $v = stock ] || lib=$PWD/luminary_b200/liblumb200_$v.so
    LUMB200_LIBRARY=$lib timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/${T}_tmp.json 2> gpurun_out/${T}_tmp.err
    echo "$wl $v run $rep: $(line gpurun_out/${T}_tmp.json)" | tee -a gpurun_out/${T}_variants_hitmask.txt
done; done; done
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err; echo "bench: $(line gpurun_out/${T}_bench.json)"
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 800 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-measure > gpurun_out/${T}_ncu_launch_run.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,launch__grid_size,launch__block_size,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
cap() { # workload skip count
  timeout 700 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_shade|k_trace_shadow|k_trace_enum" -s $2 -c $3 -f -o /tmp/${T}_full_$1 \
    python bench.py --workload $1 --steps 1 --warmup 1 --no-cpu --no-measure > gpurun_out/${T}_ncu_full_$1.log 2>&1
  ncu -i /tmp/${T}_full_$1.ncu-rep --page raw --csv --metrics $M > gpurun_out/${T}_full_$1_raw.csv 2>> gpurun_out/${T}_ncu_full_$1.log
}
cap atrium1m 36 12
cap terrain10m 30 10
ncu -i /tmp/${T}_full_atrium1m.ncu-rep --page source --csv -k regex:"k_trace_closest" > gpurun_out/${T}_full_atrium1m_closest_source.csv 2>/dev/null
ncu -i /tmp/${T}_full_atrium1m.ncu-rep --page source --csv -k regex:"k_trace_shadow" > gpurun_out/${T}_full_atrium1m_shadow_source.csv 2>/dev/null
du -sh gpurun_out; ls gpurun_out | grep ${T} | tail -30
