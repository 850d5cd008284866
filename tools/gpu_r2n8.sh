#!/bin/bash
# 8 GPUs: bit-equality check at 8 ranks, bench at N = 8 (with the KPZ 1024^3 extra) and N = 4
TAG=${1:-r2n8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/${TAG}_mg8.log 2>&1
echo "multi_gpu_check N=8 exit $?"; grep -c "bitwise=True" gpurun_out/${TAG}_mg8.log; grep "bitwise=False" gpurun_out/${TAG}_mg8.log | head -5; tail -2 gpurun_out/${TAG}_mg8.log
: > gpurun_out/${TAG}_scaling.jsonl
for n in 8 4; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/${TAG}_n$n.err | grep '^{' > gpurun_out/${TAG}_n$n.json
  echo "n=$n exit $?"; cat gpurun_out/${TAG}_n$n.json >> gpurun_out/${TAG}_scaling.jsonl
  python tools/show_extras.py gpurun_out/${TAG}_n$n.json; tail -3 gpurun_out/${TAG}_n$n.err
done
