#!/bin/bash
# Round 2, GPU visit 6: procedural sky + sun NEE. Golden fixture from the reference's kernels, whole GPU suite, default bench.
mkdir -p gpurun_out
timeout 600 python tests/golden/make_sky_golden.py gpurun_out/sky_ref.npz > gpurun_out/r2f_sky_golden.log 2>&1; echo "golden exit $?" >> gpurun_out/r2f_sky_golden.log
tail -8 gpurun_out/r2f_sky_golden.log
mkdir -p tests/golden; cp gpurun_out/sky_ref.npz tests/golden/sky_ref.npz 2>/dev/null
timeout 900 python -m pytest tests/test_sky_gpu.py tests/test_sky_oracle.py -q -s > gpurun_out/r2f_pytest_sky.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest_sky.log
grep -E "passed|failed|error|Error|assert" gpurun_out/r2f_pytest_sky.log | tail -30
timeout 1800 python -m pytest tests -m gpu -q -s --deselect tests/test_sky_gpu.py > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --cpu-seconds 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 600 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r2f_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("bench:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernel_ms_per_step"])
PY
