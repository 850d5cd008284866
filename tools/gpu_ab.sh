#!/bin/bash
# A/B of environment-selected kernel variants in one gpurun call.
# Usage: bash tools/gpu_ab.sh <tag> "VAR=a" "VAR=b" ...   (each argument is an env assignment list for one bench run)
TAG=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "round_trip or compiled_reference or bitwise or unit_tests" > gpurun_out/${TAG}_pytest.log 2>&1
tail -2 gpurun_out/${TAG}_pytest.log
i=0
for V in "$@"; do
  env $V timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ab$i.json 2> gpurun_out/${TAG}_ab$i.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_ab$i.json").read().strip().splitlines()[-1])
print("$V", round(d["value"],1), "steps/s", {k:v["ms"] for k,v in d["roofline"]["per_kernel"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
  i=$((i+1))
done
