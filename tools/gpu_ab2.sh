#!/bin/bash
# A/B of env-selected variants, no test suite.  Usage: bash tools/gpu_ab2.sh <tag> "VAR=a" "VAR=b" ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for V in "$@"; do
  env $V timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-context --no-extra > gpurun_out/${TAG}_ab$i.json 2> gpurun_out/${TAG}_ab$i.err
  echo "$V"; python tools/show_extras.py gpurun_out/${TAG}_ab$i.json || tail -5 gpurun_out/${TAG}_ab$i.err
  i=$((i+1))
done
