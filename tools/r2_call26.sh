#!/bin/bash
# Round 2, GPU visit 26: cost of the sky modes on the open-ceiling divergence workload (constant colour vs procedural vs HDRI)
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); k=d["kernel_ms_per_step"]
        print(f'{d["value"]:.1f} Mrays/s {d["ms_per_step"]:.3f} ms rays/step {d["rays_per_step"]:.0f} closest {k["trace_closest"]:.3f} shadow {k["trace_shadow"]:.3f} shade {k["shade"]:.3f} enum {k.get("trace_enum",0):.3f} nonfinite {d["nonfinite_samples"]}')
PY
}
for wl in divergence divergence_sky divergence_hdri; do
  timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2z_$wl.json 2> gpurun_out/r2z_$wl.err
  echo "$wl: $(line gpurun_out/r2z_$wl.json)" | tee -a gpurun_out/r2z_sky_modes.txt
  tail -2 gpurun_out/r2z_$wl.err | cut -c1-200
done
