#!/bin/bash
# Round 2, GPU visit 7: sky tests after the fixes (double-precision celestial positions, LUT floors)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sky_gpu.py tests/test_sky_oracle.py "tests/test_host_api_gpu.py::test_procedural_sky_through_the_public_api" -q -s > gpurun_out/r2g_pytest_sky.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest_sky.log
grep -E "passed|failed|error|Error|assert" gpurun_out/r2g_pytest_sky.log | tail -30
