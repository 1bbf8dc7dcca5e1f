#!/bin/bash
# Round 2, GPU visit 20: the moon's surface (textures through the C ABI, oracle, reference kernel), sky suite, host tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sky_gpu.py -q -s > gpurun_out/r2t_pytest_sky.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2t_pytest_sky.log
grep -E "passed|failed|rror|assert|moon" gpurun_out/r2t_pytest_sky.log | tail -20
timeout 900 python -m pytest tests/test_host_api_gpu.py tests/test_reference_frontend.py -q > gpurun_out/r2t_pytest_host.log 2>&1; tail -4 gpurun_out/r2t_pytest_host.log
