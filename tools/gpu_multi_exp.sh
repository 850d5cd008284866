#!/bin/bash
# Multi-GPU overlap experiments at N ranks: z-chunked x / pushed-y overlap (side lane at low stream priority).
N=${1:-2}
run() { echo "== $*"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1),'steps/s', {k:v['ms'] for k,v in d['roofline']['per_kernel'].items()})"; }
run CUPSS_B200_XCHUNKS=1
run CUPSS_B200_XCHUNKS=2
run CUPSS_B200_XCHUNKS=4
run CUPSS_B200_XCHUNKS=8
