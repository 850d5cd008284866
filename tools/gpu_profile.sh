#!/bin/bash
# One gpurun call (1 GPU): bench line, ncu launch list, full capture of the four step kernels (CH-3D 512^3) and of the
# cluster k stage (CH-2D 4096^2).  Usage: bash tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r2p}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; python tools/show_extras.py gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err
echo "reference arm exit $?"; tail -c 600 gpurun_out/${TAG}_reference_arm.json
P="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-context --no-extra"
CUPSS_B200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3 -c 60 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py $P > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launches exit $?"
CUPSS_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'xpass|axis_' -s 11 -c 4 \
  -f -o gpurun_out/${TAG}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-context --no-extra > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
CUPSS_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cluster' -s 4 -c 2 \
  -f -o gpurun_out/${TAG}_cluster_prof python tools/bench_configs.py --only ch2d --steps 5 > gpurun_out/${TAG}_ncu_cluster.log 2>&1
echo "ncu cluster exit $?"
ls -la gpurun_out | grep ${TAG}
