#!/bin/bash
# Round 2, GPU visit 41: compile-time variants (VARIANTS="name ...", luminary_b200/liblumb200_<name>.so) against the stock library
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
for wl in ${WORKLOADS:-atrium1m}; do
for v in stock ${VARIANTS} stock; do
    lib=$PWD/luminary_b200/liblumb200.so; [ $v = stock ] || lib=$PWD/luminary_b200/liblumb200_$v.so
    LUMB200_LIBRARY=$lib timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2ao_tmp.json 2> gpurun_out/r2ao_tmp.err
    echo "$wl $v: $(line gpurun_out/r2ao_tmp.json)" | tee -a gpurun_out/${OUT:-r2ao_variants.txt}
done; done
