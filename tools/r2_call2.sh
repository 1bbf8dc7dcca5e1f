#!/bin/bash
# Round 2, GPU visit 2: class-specialised shading kernels + queued emitter enumeration + per-slot shadow queues, L2 window off.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -12 gpurun_out/r2b_pytest.log
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]; r=d["roofline"]["per_ray"]; b=d["bvh"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} enum {k.get("trace_enum",0):.3f} sort {k["sort"]:.3f} nodes/ray {r["nodes_visited"]:.2f} ovf {b["stack_overflows"]} rays/step {d["rays_per_step"]:.0f}')
PY
}
for wl in atrium1m terrain10m divergence; do
  timeout 600 python bench.py --workload $wl --steps 12 --warmup 3 --no-cpu > gpurun_out/r2b_${wl}.json 2> gpurun_out/r2b_${wl}.err
  echo "$wl stock: $(line gpurun_out/r2b_${wl}.json)" | tee -a gpurun_out/r2b_variants.txt
done
for v in ${VARIANTS}; do
  for wl in atrium1m divergence; do
    LUMB200_LIBRARY=$PWD/luminary_b200/liblumb200_$v.so timeout 600 python bench.py --workload $wl --steps 12 --warmup 3 --no-cpu > gpurun_out/r2b_${wl}_$v.json 2> gpurun_out/r2b_${wl}_$v.err
    echo "$wl variant $v: $(line gpurun_out/r2b_${wl}_$v.json)" | tee -a gpurun_out/r2b_variants.txt
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_shade|k_trace_shadow|k_trace_enum" -s 24 -c 8 -f -o gpurun_out/r2b_full \
  python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2b_ncu_full_run.log 2>&1
ls -la gpurun_out | grep r2b | tail -30
