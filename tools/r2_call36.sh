#!/bin/bash
# Round 2, GPU visit 36: debug shading modes (product vs oracle, public API) on the GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_debug_modes.py tests/test_host_api_gpu.py -m gpu -q -s > gpurun_out/r2aj_pytest_debug_modes.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2aj_pytest_debug_modes.log
grep -E "debug mode|passed|failed|Error|error|exit" gpurun_out/r2aj_pytest_debug_modes.log | tail -30
