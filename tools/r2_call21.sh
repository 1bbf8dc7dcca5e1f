#!/bin/bash
# Round 2, GPU visit 21: golden fixture extended by the moon's disc and aerial perspective (earlier entries must reproduce bit for bit)
mkdir -p gpurun_out
timeout 600 python tests/golden/make_sky_golden.py gpurun_out/sky_ref.npz > gpurun_out/r2u_sky_golden.log 2>&1; echo "golden exit $?" >> gpurun_out/r2u_sky_golden.log
tail -6 gpurun_out/r2u_sky_golden.log
python - <<'PY'
import numpy as np
a=np.load("tests/golden/sky_ref.npz"); b=np.load("gpurun_out/sky_ref.npz")
print("earlier golden entries reproduced bit for bit:", all(np.array_equal(a[k], b[k]) for k in a.files), "new keys:", sorted(set(b.files)-set(a.files)))
PY
cp gpurun_out/sky_ref.npz tests/golden/sky_ref.npz
timeout 900 python -m pytest tests/test_sky_oracle.py -q -s 2>&1 | grep -E "moon|aerial|passed|failed|assert" | head -12
