#!/bin/bash
# Round 2, GPU visit 33: reinsertion on by default (32 passes max, early stop), two-step conflict resolution: full suite, smoke, node visits
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]; r=d["roofline"]["per_ray"]; b=d["bvh"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} nodes/ray {r["nodes_visited"]:.2f} tris/ray {r["tris_tested"]:.2f} shadow nodes {r["shadow_nodes_visited"]:.2f} SAH {b["sah_cost"]:.3f} radius {b["ploc_radius"]} depth {b["depth"]} build {b["build_ms"]:.1f} ms ovf {b["stack_overflows"]}')
PY
}
for cfg in "atrium1m 0" "atrium1m 32" "terrain10m 0" "terrain10m 32" "divergence 0" "divergence 32"; do
    set -- $cfg; wl=$1; p=$2
    LUMB200_BVH_VERBOSE=1 LUMB200_BVH_REINSERT=$p timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2ag_tmp.json 2> gpurun_out/r2ag_tmp_${wl}_$p.err
    echo "$wl reinsert $p: $(line gpurun_out/r2ag_tmp.json)" | tee -a gpurun_out/r2ag_reinsert.txt
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2ag_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ag_pytest.log; tail -3 gpurun_out/r2ag_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ag_smoke.log 2>&1; tail -2 gpurun_out/r2ag_smoke.log
