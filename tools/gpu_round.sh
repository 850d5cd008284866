#!/bin/bash
# One gpurun call: GPU test suite, bench line, ncu launch list and one full capture of the step kernels.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh <tag> [tests|notests]
set -u
TAG=${1:-r01}
WHAT=${2:-tests}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ "$WHAT" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json
# launch list (cold-cache, serialised: shares only)
CUPSS_B200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3 -c 60 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launches exit $?"
# full capture of one step's kernels (skip upload's 3 + two steps)
CUPSS_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'xpass|axis_' -s 11 -c 4 \
  -f -o gpurun_out/${TAG}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | tail -20
