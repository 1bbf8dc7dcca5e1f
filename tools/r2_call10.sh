#!/bin/bash
# Round 2, GPU visit 10 (8 GPUs): headline bench at N=8, MEASURED 4K 1024-spp render at N=8, the C host's front end on 8 devices,
# the in-process 2-device tests of the C host. Everything bounded by timeouts; logs under gpurun_out/r2j_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2j_smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 16 --warmup 3 --no-cpu > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err
timeout 900 $TR --master-port 29522 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu --workload atrium4k --spp 1024 > gpurun_out/r2j_bench_4k_n8.json 2> gpurun_out/r2j_bench_4k_n8.err
python - <<'PY'
import json
for f in ("r2j_bench_n8", "r2j_bench_4k_n8"):
    for l in open(f"gpurun_out/{f}.json"):
        if l.startswith("{"):
            d=json.loads(l); print(f, d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "time_to_spp", d.get("time_to_spp"), "reduce_check", d.get("reduce_check"))
PY
tail -3 gpurun_out/r2j_bench_4k_n8.err
timeout 900 python tools/cli_multi_gpu.py --gpus 8 --log2-samples 10 --out gpurun_out/r2j_cli_n8.json > gpurun_out/r2j_cli_n8.log 2>&1; tail -4 gpurun_out/r2j_cli_n8.log
timeout 600 python -m pytest tests/test_host_api_gpu.py tests/test_sharding.py -q -s -k "two_devices or sharding or reduce" > gpurun_out/r2j_pytest_two_devices.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest_two_devices.log
tail -5 gpurun_out/r2j_pytest_two_devices.log
