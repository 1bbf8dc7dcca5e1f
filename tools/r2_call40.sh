#!/bin/bash
# Round 2, GPU visit 40: whole suite + smoke + headline bench with the 64-blocks-per-SM k_shade grid
mkdir -p gpurun_out
T=r2an
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; grep "^{" gpurun_out/${T}_bench.log > gpurun_out/${T}_bench_n1.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2an_bench_n1.json").readline())
print("bench:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["kernel_ms_per_step"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
PY
