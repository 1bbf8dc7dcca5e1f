#!/bin/bash
# Round 2, GPU visit 24 (2 GPUs): final state - whole GPU suite incl. the 2-device host tests, smoke(), default bench, N=2 bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2x_pytest.log
tail -5 gpurun_out/r2x_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; tail -3 gpurun_out/r2x_smoke.log
timeout 600 python bench.py > gpurun_out/r2x_bench_n1.json 2> gpurun_out/r2x_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_ref.json 2> gpurun_out/r2x_bench_ref.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 16 --warmup 3 > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err
python - <<'PY'
import json
for f in ("r2x_bench_n1", "r2x_bench_ref", "r2x_bench_n2"):
    for l in open(f"gpurun_out/{f}.json"):
        if l.startswith("{"):
            d=json.loads(l); print(f, d.get("impl","ours"), d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("nonfinite_samples"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("traffic"), d.get("clocks"))
PY
