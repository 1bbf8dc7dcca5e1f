#!/bin/bash
# Round 2, GPU visit 8: packed node test (HADD2.F32 + FFMA2) and pipelined root-children loop, A/B on three workloads; parity of the trace kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trace_gpu.py tests/test_configs_gpu.py -q -x > gpurun_out/r2h_pytest_trace.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest_trace.log
tail -4 gpurun_out/r2h_pytest_trace.log
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]; r=d["roofline"]["per_ray"]; b=d["bvh"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} enum {k.get("trace_enum",0):.3f} sort {k["sort"]:.3f} nodes/ray {r["nodes_visited"]:.2f} ovf {b["stack_overflows"]}')
PY
}
for rep in 1 2; do
for v in stock unpacked pipe; do
  for wl in atrium1m divergence terrain10m; do
    lib=$PWD/luminary_b200/liblumb200_$v.so; [ $v = stock ] && lib=$PWD/luminary_b200/liblumb200.so
    LUMB200_LIBRARY=$lib timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2h_${wl}_$v.json 2> gpurun_out/r2h_${wl}_$v.err
    echo "$wl variant $v run $rep: $(line gpurun_out/r2h_${wl}_$v.json)" | tee -a gpurun_out/r2h_variants.txt
  done
done
done
LUMB200_LIBRARY=$PWD/luminary_b200/liblumb200_pipe.so timeout 900 python -m pytest tests/test_shade_vertices_gpu.py tests/test_render_gpu.py -q -x > gpurun_out/r2h_pytest_pipe.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest_pipe.log
tail -4 gpurun_out/r2h_pytest_pipe.log
