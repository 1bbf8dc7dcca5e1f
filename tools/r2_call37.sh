#!/bin/bash
# Round 2, GPU visit 37: compute-sanitizer memcheck on the debug shading kernel and on the traversal kernels with the rewritten hit mask
mkdir -p gpurun_out
run() { # name tool tests...
  name=$1; tool=$2; shift 2
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 17 python -m pytest "$@" -q -x > gpurun_out/r2ak_sanitizer_$name.log 2>&1
  echo "$name ($tool): exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r2ak_sanitizer_$name.log | tr '\n' ' ')" | tee -a gpurun_out/r2ak_sanitizer_summary.txt
}
run debug_mem memcheck tests/test_debug_modes.py -m gpu
run trace_mem memcheck tests/test_trace_gpu.py
