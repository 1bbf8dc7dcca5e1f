#!/bin/bash
# Round 2, GPU visit 18: aerial perspective (sky_process_inscattering_events) - parity vs oracle and vs the reference kernel; sky suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sky_gpu.py -q -s > gpurun_out/r2r_pytest_sky.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2r_pytest_sky.log
grep -E "passed|failed|rror|assert|aerial" gpurun_out/r2r_pytest_sky.log | tail -30
