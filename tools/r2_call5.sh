#!/bin/bash
# Round 2, GPU visit 5: whole suite after the output-chain/host fixes, root-children staging A/B, default bench, ncu launch list + full captures
# (raw pages exported on the box; the .ncu-rep files stay there unless small, gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
tail -6 gpurun_out/r2e_pytest.log
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]; r=d["roofline"]["per_ray"]; b=d["bvh"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} enum {k.get("trace_enum",0):.3f} sort {k["sort"]:.3f} nodes/ray {r["nodes_visited"]:.2f} ovf {b["stack_overflows"]} roofline {d["roofline"]["kernel"]} frac {d["roofline"]["frac"]:.3f}')
PY
}
for rep in 1 2; do
for v in stock nostage; do
  for wl in atrium1m divergence; do
    lib=$PWD/luminary_b200/liblumb200_$v.so; [ $v = stock ] && lib=$PWD/luminary_b200/liblumb200.so
    LUMB200_LIBRARY=$lib timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2e_${wl}_$v.json 2> gpurun_out/r2e_${wl}_$v.err
    echo "$wl variant $v run $rep: $(line gpurun_out/r2e_${wl}_$v.json)" | tee -a gpurun_out/r2e_variants.txt
  done
done
done
timeout 600 python bench.py --workload terrain10m --steps 16 --warmup 3 --no-cpu > gpurun_out/r2e_terrain10m.json 2> gpurun_out/r2e_terrain10m.err
echo "terrain10m stock: $(line gpurun_out/r2e_terrain10m.json)" | tee -a gpurun_out/r2e_variants.txt
timeout 600 python bench.py --cpu-seconds 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 1800 gpurun_out/r2e_bench.json; tail -3 gpurun_out/r2e_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 800 --csv --log-file gpurun_out/r2e_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-measure > gpurun_out/r2e_ncu_launch_run.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,launch__grid_size,launch__block_size"
cap() { # workload skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_shade|k_trace_shadow|k_trace_enum" -s $2 -c $3 -f -o /tmp/r2e_full_$1 \
    python bench.py --workload $1 --steps 1 --warmup 1 --no-cpu --no-measure > gpurun_out/r2e_ncu_full_$1.log 2>&1
  ncu -i /tmp/r2e_full_$1.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2e_full_$1_raw.csv 2>> gpurun_out/r2e_ncu_full_$1.log
  ls -la /tmp/r2e_full_$1.ncu-rep
}
# one warm-up pass = 6 depths x (closest, shade dielectric, shade metal, shade miss, enum, shadow) = 36 matching launches on the atrium
cap atrium1m 36 12
cap terrain10m 30 10
cap divergence 54 12
ncu -i /tmp/r2e_full_atrium1m.ncu-rep --page source --csv -k regex:"k_shade" > gpurun_out/r2e_full_atrium1m_shade_source.csv 2>/dev/null
sz=$(stat -c %s /tmp/r2e_full_atrium1m.ncu-rep); [ "$sz" -lt 30000000 ] && cp /tmp/r2e_full_atrium1m.ncu-rep gpurun_out/r2e_full_atrium1m.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | grep r2e | tail -30
