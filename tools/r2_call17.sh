#!/bin/bash
# Round 2, GPU visit 17: compute-sanitizer on the round-2 kernels (sky / HDRI, material-class shading with staged root children,
# queued emitter enumeration, packed node test, non-finite drop, NCCL-free single device paths)
mkdir -p gpurun_out
run() { # name tool tests...
  name=$1; tool=$2; shift 2
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 17 python -m pytest "$@" -q -x > gpurun_out/r2q_sanitizer_$name.log 2>&1
  echo "$name ($tool): exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r2q_sanitizer_$name.log | tr '\n' ' ')" | tee -a gpurun_out/r2q_sanitizer_summary.txt
}
run sky_mem memcheck tests/test_sky_gpu.py
run shade_mem memcheck tests/test_shade_vertices_gpu.py tests/test_render_gpu.py
run trace_mem memcheck tests/test_trace_gpu.py
run sky_race racecheck tests/test_sky_gpu.py -k "hdri_mode or tables or image_under_the_procedural"
run shade_race racecheck tests/test_render_gpu.py -k "lit_room or translucent"
