#!/bin/bash
# Round 2, GPU visit 29: L2 prefetch of the next batch's ray records in the persistent trace loop
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f}')
PY
}
for rep in 1 2; do
for v in stock tpf; do
  for wl in atrium1m terrain10m; do
    lib=$PWD/luminary_b200/liblumb200_$v.so; [ $v = stock ] && lib=$PWD/luminary_b200/liblumb200.so
    LUMB200_LIBRARY=$lib timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2ac_tmp.json 2> gpurun_out/r2ac_tmp.err
    echo "$wl variant $v run $rep: $(line gpurun_out/r2ac_tmp.json)" | tee -a gpurun_out/r2ac_variants.txt
  done
done
done
LUMB200_LIBRARY=$PWD/luminary_b200/liblumb200_tpf.so timeout 900 python -m pytest tests/test_trace_gpu.py tests/test_render_gpu.py -q -x 2>&1 | tail -2
