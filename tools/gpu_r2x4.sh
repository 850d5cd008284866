#!/bin/bash
# Pruned strided level of the three-level x pass: parity of the long-line Cahn-Hilliard cases, then CH-2D 4096^2 A/B
TAG=${1:-r2x4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ch2d or ch3d or quartic or sh3d or sh2d or environment_variants or burgers" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
for V in "A=1" "CUPSS_B200_X4_NOPRUNE=1" "A=2" "CUPSS_B200_X4_NOPRUNE=1"; do
  echo "== $V"
  env $V timeout 300 python tools/bench_configs.py --only ch2d --steps 200 2> gpurun_out/${TAG}_cfg.err | tee -a gpurun_out/${TAG}_ch2d.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(round(d['steps_per_s'],1), {k:(v['ms'],v['launches']) for k,v in d['per_kernel'].items()})"
done
