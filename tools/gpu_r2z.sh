#!/bin/bash
# Staged source rows for sweeps without a forward transform: parity, then KPZ-3D 512^3 A/B
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kpz or ops or random_systems or modelh_32 or environment_variants or bc_ or noisy or plan_specialised" > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
for V in "A=1" "CUPSS_B200_NO_STAGE0=1"; do
  echo "== $V"
  env $V timeout 300 python tools/bench_configs.py --only kpz,modelh --steps 100 2> gpurun_out/${TAG}_cfg.err | tee -a gpurun_out/${TAG}_cfg.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(round(d['steps_per_s'],1), {k:(v['ms'],v['launches']) for k,v in d['per_kernel'].items()})"
done
