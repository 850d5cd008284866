#!/bin/bash
# Run-time compiled k stage variants: parity, then KPZ-3D 512^3 / Model H 2048^2
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kpz or ops or random_systems or modelh or environment_variants or bc_ or noisy or plan_specialised or mixed" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/bench_configs.py --only kpz,modelh --steps 100 2> gpurun_out/${TAG}_cfg.err | tee -a gpurun_out/${TAG}_cfg.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(round(d['steps_per_s'],1), {k:(v['ms'],v['launches']) for k,v in d['per_kernel'].items()})"
