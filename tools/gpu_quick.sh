#!/bin/bash
# Quick GPU check: a few parity tests + bench line (+ optional per-kernel ncu).  Usage: bash tools/gpu_quick.sh <tag> [ncu]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "round_trip or compiled_reference or bitwise or unit_tests" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("${TAG}", round(d["value"],1), "steps/s")
for k,v in d["roofline"]["per_kernel"].items(): print("   ",k,v, round(v["GBps"]/6556.5,3))
PY
if [ "${2:-}" = "ncu" ]; then
CUPSS_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'xpass|axis_' -s 11 -c 4 \
  -f -o gpurun_out/${TAG}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
