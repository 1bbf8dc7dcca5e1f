#!/bin/bash
# 1 -> 8 GPU scaling of config 5 (4K, sample-id sharding + NCCL reduce of the planes) and the headline config at N = 8.
# usage (under gpurun --gpus 8): bash tools/scale_run.sh <tag>
tag=${1:-scale}
mkdir -p gpurun_out
run() { # N workload steps
  if [ "$1" = 1 ]; then
    python bench.py --gpus 1 --workload $2 --steps $3 --warmup 3 --cpu-seconds 1
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --workload $2 --steps $3 --warmup 3 --cpu-seconds 1
  fi
}
for n in 1 2 4 8; do
  run $n atrium4k 16 2>gpurun_out/${tag}_4k_n$n.err | grep '^{' > gpurun_out/${tag}_4k_n$n.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_4k_n$n.json").read().strip().splitlines()[-1])
print("4K N=$n: %.1f Mrays/s  %.2f ms/step  spp/s %.2f  time-to-1024spp %.2f s  e2e %.1f" % (d["value"], d["ms_per_step"], d["spp_per_s"], d["time_to_1024spp_s"], d["e2e"]["value"]))
PY
done
run 8 atrium1m 16 2>gpurun_out/${tag}_1080p_n8.err | grep '^{' > gpurun_out/${tag}_1080p_n8.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_1080p_n8.json").read().strip().splitlines()[-1])
print("1080p N=8: %.1f Mrays/s  %.2f ms/step  spp/s %.2f  e2e %.1f" % (d["value"], d["ms_per_step"], d["spp_per_s"], d["e2e"]["value"]))
PY
