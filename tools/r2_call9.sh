#!/bin/bash
# Round 2, GPU visit 9: HDRI sky mode. Golden regenerated (same inputs + HDRI tables), sky tests, host test, whole suite.
mkdir -p gpurun_out
timeout 600 python tests/golden/make_sky_golden.py gpurun_out/sky_ref.npz > gpurun_out/r2i_sky_golden.log 2>&1; echo "golden exit $?" >> gpurun_out/r2i_sky_golden.log
tail -6 gpurun_out/r2i_sky_golden.log
python - <<'PY'
import numpy as np
a=np.load("tests/golden/sky_ref.npz"); b=np.load("gpurun_out/sky_ref.npz")
same=all(np.array_equal(a[k], b[k]) for k in a.files)
print("golden entries of the previous fixture reproduced bit for bit:", same, "new keys:", sorted(set(b.files)-set(a.files)))
PY
cp gpurun_out/sky_ref.npz tests/golden/sky_ref.npz
timeout 900 python -m pytest tests/test_sky_gpu.py tests/test_sky_oracle.py -q -s > gpurun_out/r2i_pytest_sky.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest_sky.log
grep -E "passed|failed|error|Error|assert|HDRI" gpurun_out/r2i_pytest_sky.log | tail -40
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_sky_gpu.py > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
