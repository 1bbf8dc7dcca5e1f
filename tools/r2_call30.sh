#!/bin/bash
# Round 2, GPU visit 30: parallel reinsertion on the PLOC hierarchy (LUMB200_BVH_REINSERT passes): SAH, node visits, stage times
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]; b=d.get("bvh",{})
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} | bvh {json.dumps(b)}')
PY
}
# correctness first: a small tree with many passes, hits must stay bit-identical
LUMB200_BVH_REINSERT=8 timeout 900 python -m pytest tests/test_trace_gpu.py -q -x 2>&1 | tail -3 | tee gpurun_out/r2ad_pytest_trace.txt
for wl in atrium1m terrain10m divergence; do
for p in 0 2 8 32; do
    [ $wl = divergence ] && [ $p != 0 ] && [ $p != 8 ] && continue
    LUMB200_BVH_VERBOSE=1 LUMB200_BVH_REINSERT=$p timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2ad_tmp.json 2> gpurun_out/r2ad_tmp_${wl}_$p.err
    echo "$wl reinsert $p: $(line gpurun_out/r2ad_tmp.json)" | tee -a gpurun_out/r2ad_reinsert.txt
    grep "reinsertion\|SAH" gpurun_out/r2ad_tmp_${wl}_$p.err | tail -40 >> gpurun_out/r2ad_reinsert_log.txt
done
done
LUMB200_BVH_REINSERT=8 timeout 900 python -m pytest tests/test_trace_gpu.py tests/test_render_gpu.py tests/test_configs_gpu.py -q -x 2>&1 | tail -3 | tee -a gpurun_out/r2ad_pytest_trace.txt
