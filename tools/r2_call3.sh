#!/bin/bash
# Round 2, GPU visit 3: suite with measurement-based thresholds, new bench.py, shared-memory stack variants, ncu launch list + full captures.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -6 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --cpu-seconds 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 1500 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]; r=d["roofline"]["per_ray"]; b=d["bvh"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} enum {k.get("trace_enum",0):.3f} sort {k["sort"]:.3f} nodes/ray {r["nodes_visited"]:.2f} ovf {b["stack_overflows"]} roofline {d["roofline"]["kernel"]} frac {d["roofline"]["frac"]:.3f}')
PY
}
for v in ${VARIANTS}; do
  for wl in atrium1m terrain10m; do
    LUMB200_LIBRARY=$PWD/luminary_b200/liblumb200_$v.so timeout 600 python bench.py --workload $wl --steps 12 --warmup 3 --no-cpu > gpurun_out/r2c_${wl}_$v.json 2> gpurun_out/r2c_${wl}_$v.err
    echo "$wl variant $v: $(line gpurun_out/r2c_${wl}_$v.json)" | tee -a gpurun_out/r2c_variants.txt
  done
done
for wl in atrium1m terrain10m divergence; do
  timeout 600 python bench.py --workload $wl --steps 12 --warmup 3 --no-cpu > gpurun_out/r2c_${wl}.json 2> gpurun_out/r2c_${wl}.err
  echo "$wl stock: $(line gpurun_out/r2c_${wl}.json)" | tee -a gpurun_out/r2c_variants.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 800 --csv --log-file gpurun_out/r2c_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-measure > gpurun_out/r2c_ncu_launch_run.log 2>&1
# one warm-up pass = 6 depths x (closest, shade dielectric, shade metal, shade miss, enum, shadow) = 36 matching launches on the atrium
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_shade|k_trace_shadow|k_trace_enum" -s 36 -c 12 -f -o gpurun_out/r2c_full_atrium1m \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-measure > gpurun_out/r2c_ncu_full_atrium.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_trace_closest|k_shade|k_trace_shadow|k_trace_enum" -s 30 -c 10 -f -o gpurun_out/r2c_full_terrain10m \
  python bench.py --workload terrain10m --steps 1 --warmup 1 --no-cpu --no-measure > gpurun_out/r2c_ncu_full_terrain.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_trace_closest|k_shade|k_trace_shadow|k_trace_enum" -s 54 -c 12 -f -o gpurun_out/r2c_full_divergence \
  python bench.py --workload divergence --steps 1 --warmup 1 --no-cpu --no-measure > gpurun_out/r2c_ncu_full_divergence.log 2>&1
ls -la gpurun_out | grep r2c | tail -30
