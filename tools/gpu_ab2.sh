#!/bin/bash
# A/B of environment-selected variants on the headline bench only.  Usage: bash tools/gpu_ab2.sh <tag> "VAR=a" "VAR=b" ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for V in "$@"; do
  env $V timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-context --no-extra > gpurun_out/${TAG}_ab$i.json 2> gpurun_out/${TAG}_ab$i.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_ab$i.json").read().strip().splitlines()[-1])
print("$V", round(d["value"],1), "steps/s", {k:v["ms"] for k,v in d["roofline"]["per_kernel"].items()}, "parity", d["parity"] and d["parity"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
  i=$((i+1))
done
