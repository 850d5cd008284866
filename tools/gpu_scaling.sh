#!/bin/bash
# Strong-scaling table on ONE box: bench.py at N = 8, 4, 2, 1 back to back (as the round-end driver does).  Usage: bash tools/gpu_scaling.sh <tag>
TAG=$1
mkdir -p gpurun_out
: > gpurun_out/${TAG}_scaling.jsonl
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/${TAG}_n$n.err | grep '^{' >> gpurun_out/${TAG}_scaling.jsonl
  echo "n=$n exit $?"
done
timeout 400 python bench.py --gpus 1 --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/${TAG}_n1.err | grep '^{' >> gpurun_out/${TAG}_scaling.jsonl
python - <<PY
import json
rows=[json.loads(l) for l in open("gpurun_out/${TAG}_scaling.jsonl")]
one=[r for r in rows if r["n_gpus"]==1]
for r in rows:
    eff = r["value"]/(one[0]["value"]*r["n_gpus"]) if one else float("nan")
    print(r["n_gpus"], round(r["value"],1), "steps/s  eff", round(eff,3), " e2e", round(r["e2e"]["value"],1), {k:v["ms"] for k,v in r["roofline"]["per_kernel"].items()})
PY
