#!/bin/bash
# One GPU-box visit: parity tests, headline bench, ncu launch list, one ncu --set full capture of the top kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 3000 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; tail -c 1200 gpurun_out/${tag}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --cpu-seconds 1 > gpurun_out/${tag}_ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_shade|k_trace_shadow" -s 18 -c 6 -f -o gpurun_out/${tag}_full \
  python bench.py --steps 1 --warmup 1 --cpu-seconds 1 > gpurun_out/${tag}_ncu_full_run.log 2>&1
ls -la gpurun_out | tail -20
