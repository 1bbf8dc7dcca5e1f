#!/bin/bash
# Round 2, GPU visit 19: three compile-time variants (traversal stack of 24 entries; root-children loop unrolled by 2; packed FMA in the
# reservoir lane update) against stock on two workloads, twice; golden regeneration check after the harness change
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} ovf {d["bvh"]["stack_overflows"]}')
PY
}
for rep in 1 2; do
for v in stock stack24 unroll2 ffma2; do
  for wl in atrium1m terrain10m; do
    lib=$PWD/luminary_b200/liblumb200_$v.so; [ $v = stock ] && lib=$PWD/luminary_b200/liblumb200.so
    LUMB200_LIBRARY=$lib timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2s_tmp.json 2> gpurun_out/r2s_tmp.err
    echo "$wl variant $v run $rep: $(line gpurun_out/r2s_tmp.json)" | tee -a gpurun_out/r2s_variants.txt
  done
done
done
for v in unroll2 ffma2; do
LUMB200_LIBRARY=$PWD/luminary_b200/liblumb200_$v.so timeout 900 python -m pytest tests/test_shade_vertices_gpu.py tests/test_render_gpu.py -q -x 2>&1 | tail -2
done
timeout 600 python tests/golden/make_sky_golden.py gpurun_out/sky_ref_check.npz > gpurun_out/r2s_sky_golden.log 2>&1
python - <<'PY'
import numpy as np
a=np.load("tests/golden/sky_ref.npz"); b=np.load("gpurun_out/sky_ref_check.npz")
print("golden reproduced bit for bit after the harness change:", all(np.array_equal(a[k], b[k]) for k in a.files), sorted(set(b.files)^set(a.files)))
PY
rm -f gpurun_out/sky_ref_check.npz
