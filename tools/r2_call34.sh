#!/bin/bash
# Round 2, GPU visit 34: does the PLOC radius that is best before reinsertion stay the best after it?
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]; r=d["roofline"]["per_ray"]; b=d["bvh"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} nodes/ray {r["nodes_visited"]:.2f} tris/ray {r["tris_tested"]:.2f} shadow nodes {r["shadow_nodes_visited"]:.2f} SAH {b["sah_cost"]:.3f} radius {b["ploc_radius"]} depth {b["depth"]}')
PY
}
for wl in atrium1m terrain10m; do
for r in 8 16 32 64 128; do
    LUMB200_PLOC_RADIUS=$r timeout 600 python bench.py --workload $wl --steps 12 --warmup 3 --no-cpu > gpurun_out/r2ah_tmp.json 2> gpurun_out/r2ah_tmp.err
    echo "$wl radius $r + reinsertion: $(line gpurun_out/r2ah_tmp.json)" | tee -a gpurun_out/r2ah_radius_after_reinsertion.txt
done
done
for cp in 0.35 0.7; do
    LUMB200_SAH_CPRIM=$cp timeout 600 python bench.py --workload atrium1m --steps 12 --warmup 3 --no-cpu > gpurun_out/r2ah_tmp.json 2> gpurun_out/r2ah_tmp.err
    echo "atrium1m c_prim $cp + reinsertion: $(line gpurun_out/r2ah_tmp.json)" | tee -a gpurun_out/r2ah_radius_after_reinsertion.txt
done
