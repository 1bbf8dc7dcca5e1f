#!/bin/bash
# Round 2, GPU visit 38 (2 GPUs): the in-process 2-device tests of the C host and bench.py at N = 2 on the final kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_api_gpu.py tests/test_sharding.py -q -k "two_devices or gpu or nccl" > gpurun_out/r2al_pytest_two_devices.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2al_pytest_two_devices.log
tail -4 gpurun_out/r2al_pytest_two_devices.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 16 --warmup 3 \
  > gpurun_out/r2al_bench_n2.log 2> gpurun_out/r2al_bench_n2.err
grep "^{" gpurun_out/r2al_bench_n2.log > gpurun_out/r2al_bench_n2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2al_bench_n2.json").readline())
print("N=2:", d["value"], "Mrays/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], "reduce_check", d.get("reduce_check"))
PY
