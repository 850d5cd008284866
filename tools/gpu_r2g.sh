#!/bin/bash
# 2 GPUs: partitioned correctness with the TMA prologue / DIF inverse, bench at N=2 and N=1, full GPU suite
TAG=${1:-r2g}
mkdir -p gpurun_out
B="--steps 200 --warmup 20 --no-cpu-baseline --no-context --no-extra"
for X in 1 4; do
  CUPSS_B200_XCHUNKS=$X timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/${TAG}_mg_x$X.log 2>&1
  echo "multi_gpu_check XCHUNKS=$X exit $?"; grep -c "bitwise=True" gpurun_out/${TAG}_mg_x$X.log; grep "bitwise=False" gpurun_out/${TAG}_mg_x$X.log | head -5; tail -2 gpurun_out/${TAG}_mg_x$X.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 $B > gpurun_out/${TAG}_n2.json 2> gpurun_out/${TAG}_n2.err
python tools/show_extras.py gpurun_out/${TAG}_n2.json
timeout 300 python bench.py $B > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
python tools/show_extras.py gpurun_out/${TAG}_n1.json
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
