#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/find_nonfinite.py --workload atrium1m --spp 24 --chunk 4 --oracle 3 > gpurun_out/r2l_nonfinite.log 2>&1; tail -60 gpurun_out/r2l_nonfinite.log | cut -c1-2500
