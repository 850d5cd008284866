#!/bin/bash
# 2 GPUs: partitioned correctness (incl. noisy KPZ with the lean evaluator), bench at N = 2
TAG=${1:-r2mg}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/${TAG}_mg.log 2>&1
echo "multi_gpu_check exit $?"; grep -c "bitwise=True" gpurun_out/${TAG}_mg.log; grep "bitwise=False\|Error\|error" gpurun_out/${TAG}_mg.log | head -5; tail -2 gpurun_out/${TAG}_mg.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --no-context > gpurun_out/${TAG}_n2.json 2> gpurun_out/${TAG}_n2.err
python tools/show_extras.py gpurun_out/${TAG}_n2.json
