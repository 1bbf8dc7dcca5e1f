#!/bin/bash
# Round 2, GPU visit 16: sky suite after the HDRI-mode ambient NEE (test fixed: the oracle needs the BSDF LUTs and the light tree)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sky_gpu.py tests/test_trace_gpu.py "tests/test_host_api_gpu.py::test_procedural_sky_through_the_public_api" -q -s > gpurun_out/r2p_pytest_sky.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2p_pytest_sky.log
grep -E "passed|failed|rror|assert|ambient|HDRI-lit|aperture" gpurun_out/r2p_pytest_sky.log | tail -30
