#!/bin/bash
# Multi-GPU round: bit-equality test vs 1 GPU, then the bench at N ranks.  Usage: bash tools/gpu_multi.sh <tag> <N> [extra bench args]
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k multi_gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
for n in $N; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 200 --warmup 20 --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
  echo "n=$n exit $?"; tail -c 2500 gpurun_out/${TAG}_bench_n$n.json
done
