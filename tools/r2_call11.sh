#!/bin/bash
# Round 2, GPU visit 11: N=1 4K 1024-spp measured render (pair of the N=8 run of visit 10), hunt for non-finite samples, quick suite
mkdir -p gpurun_out
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu --workload atrium4k --spp 1024 > gpurun_out/r2k_bench_4k_n1.json 2> gpurun_out/r2k_bench_4k_n1.err
python - <<'PY'
import json
for l in open("gpurun_out/r2k_bench_4k_n1.json"):
    if l.startswith("{"):
        d=json.loads(l); print("4K N=1:", d["value"], d["ms_per_step"], "time_to_spp", d.get("time_to_spp"), "nonfinite", d.get("nonfinite_samples"))
PY
timeout 900 python tools/find_nonfinite.py --workload atrium4k --spp 256 > gpurun_out/r2k_nonfinite.log 2>&1; tail -15 gpurun_out/r2k_nonfinite.log
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_adaptive_gpu.py tests/test_abi.py -q > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
