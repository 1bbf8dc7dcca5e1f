#!/bin/bash
# Round 2, GPU visit 14: fetch / triangle-postponing threshold sweep of the persistent trace loop after the packed node test
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
for wl in atrium1m terrain10m; do
for cfg in "22 8" "22 4" "22 12" "22 16" "16 8" "26 8" "28 12" "24 10"; do
  set -- $cfg
  LUMB200_FETCH_THRESHOLD=$1 LUMB200_TRI_THRESHOLD=$2 timeout 600 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu > gpurun_out/r2n_tmp.json 2> gpurun_out/r2n_tmp.err
  echo "$wl fetch<=$1 tri>=$2: $(line gpurun_out/r2n_tmp.json)" | tee -a gpurun_out/r2n_threshold_sweep.txt
done
done
