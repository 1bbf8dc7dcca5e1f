#!/bin/bash
# Round 2, GPU visit 39: k_shade grid size. The class kernels keep 5 (opaque, 96 registers) / 4 (generic, 123) blocks of 128 threads resident per
# SM but were launched with 8 per SM (1.6 waves of statically divided work): sweep LUMB200_SHADE_BLOCKS_PER_SM.
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
for wl in ${WORKLOADS:-atrium1m divergence}; do
for b in ${BLOCKS:-8 4 5 10 15 20 40}; do
    LUMB200_SHADE_BLOCKS_PER_SM=$b timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2am_tmp.json 2> gpurun_out/r2am_tmp.err
    echo "$wl shade blocks/SM $b: $(line gpurun_out/r2am_tmp.json)" | tee -a gpurun_out/r2am_shade_grid.txt
done; done
