#!/bin/bash
# N GPUs: cyclic vs block ky distribution.  Usage: bash tools/gpu_r2m.sh <tag> <N>
TAG=${1:-r2m}; N=${2:-2}
mkdir -p gpurun_out
B="--steps 200 --warmup 20 --no-cpu-baseline --no-context --no-extra --no-e2e"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/${TAG}_mg.log 2>&1
echo "multi_gpu_check N=$N exit $?"; grep -c "bitwise=True" gpurun_out/${TAG}_mg.log; grep "bitwise=False" gpurun_out/${TAG}_mg.log | head -5; tail -2 gpurun_out/${TAG}_mg.log
for V in "CUPSS_B200_KY_BLOCK=0" "CUPSS_B200_KY_BLOCK=1" "CUPSS_B200_KY_BLOCK=0"; do
  env $V timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N $B 2> gpurun_out/${TAG}_$V.err | grep '^{' > gpurun_out/${TAG}_$V.json
  echo "$V"; python tools/show_extras.py gpurun_out/${TAG}_$V.json || tail -5 gpurun_out/${TAG}_$V.err
done
