#!/bin/bash
# Round 2, GPU visit 4 (2 GPUs): whole suite incl. the 2-device tests and the reference front end, N=2 bench with the C-ABI NCCL reduce
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2d_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -6 gpurun_out/r2d_pytest.log
grep -n "64 lights\|adaptive 1 vs 2\|sorted (class" gpurun_out/r2d_pytest.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
tail -c 1800 gpurun_out/r2d_bench_n2.json; tail -5 gpurun_out/r2d_bench_n2.err
timeout 600 python bench.py --steps 16 --warmup 3 --cpu-seconds 5 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
python - <<'PY'
import json
for n in (1,2):
    for l in open(f"gpurun_out/r2d_bench_n{n}.json"):
        if l.startswith("{"):
            d=json.loads(l); print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("reduce_check"), d["roofline"]["kernel"], d["roofline"]["frac"])
PY
