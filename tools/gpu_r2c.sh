#!/bin/bash
# Round-2 experiment call (2 GPUs): z-chunked x->y pipeline sweep on 1 GPU, column-chunked exchange pipeline on 2 GPUs.
TAG=${1:-r2c}
mkdir -p gpurun_out
B="--steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-context --no-extra"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"],1), "steps/s", {k:v["ms"] for k,v in d["roofline"]["per_kernel"].items()}, "parity", d["parity"] and d["parity"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
for Z in 0 8 16 32 64; do
  CUPSS_B200_ZCHUNK=$Z timeout 300 python bench.py $B > gpurun_out/${TAG}_z$Z.json 2> gpurun_out/${TAG}_z$Z.err
  show gpurun_out/${TAG}_z$Z.json "ZCHUNK=$Z"
done
CUPSS_B200_ZCHUNK=4 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_z4.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_z4.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
for X in 1 4; do
  CUPSS_B200_XCHUNKS=$X timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/${TAG}_mg_x$X.log 2>&1
  echo "multi_gpu_check XCHUNKS=$X exit $?"; tail -4 gpurun_out/${TAG}_mg_x$X.log
done
for X in 1 2 4 8; do
  CUPSS_B200_XCHUNKS=$X timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 $B > gpurun_out/${TAG}_n2_x$X.json 2> gpurun_out/${TAG}_n2_x$X.err
  show gpurun_out/${TAG}_n2_x$X.json "N=2 XCHUNKS=$X"
done
