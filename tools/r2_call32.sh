#!/bin/bash
# Round 2, GPU visit 32: reinsertion pass budget with the early stop (area-sum gain per pass < 0.3 %), per-pass cost
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]; b=d.get("bvh",{})
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} | sah {b["sah_cost"]:.3f} depth {b["depth"]} build_ms {b["build_ms"]:.1f} setup {json.dumps(d.get("setup_s"))}')
PY
}
for cfg in "atrium1m 0" "atrium1m 16" "atrium1m 32" "terrain10m 0" "terrain10m 32" "terrain10m 64" "divergence 32"; do
    set -- $cfg; wl=$1; p=$2
    LUMB200_BVH_VERBOSE=1 LUMB200_BVH_REINSERT=$p timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2af_tmp.json 2> gpurun_out/r2af_tmp_${wl}_$p.err
    echo "$wl reinsert $p: $(line gpurun_out/r2af_tmp.json)" | tee -a gpurun_out/r2af_reinsert.txt
done
