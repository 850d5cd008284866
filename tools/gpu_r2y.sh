#!/bin/bash
# Lean evaluator for noisy fields: noise tests, then KPZ-3D 512^3 A/B
TAG=${1:-r2y}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "noise or noisy or stochastic or bitwise or kpz" > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
for V in "A=1" "CUPSS_B200_LEAN_NOISE_MINB=2" "CUPSS_B200_NO_LEAN_NOISE=1"; do
  echo "== $V"
  env $V timeout 300 python tools/bench_configs.py --only kpz --steps 100 2> gpurun_out/${TAG}_cfg.err | tee -a gpurun_out/${TAG}_kpz.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(round(d['steps_per_s'],1), {k:(v['ms'],v['launches']) for k,v in d['per_kernel'].items()})"
done
