#!/bin/bash
# Round 2, GPU visit 15: HDRI-mode ambient NEE, aperture test, whole suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sky_gpu.py tests/test_trace_gpu.py "tests/test_host_api_gpu.py::test_procedural_sky_through_the_public_api" -q -s > gpurun_out/r2o_pytest_sky.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest_sky.log
grep -E "passed|failed|rror|assert|HDRI|aperture" gpurun_out/r2o_pytest_sky.log | tail -40
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_sky_gpu.py --deselect tests/test_trace_gpu.py > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest.log
tail -5 gpurun_out/r2o_pytest.log
