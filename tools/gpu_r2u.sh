#!/bin/bash
# Round 2, stash x pass on long lines + lane fan-out: parity of the new cases, then Model H 2048^2 A/B.
TAG=${1:-r2u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed or modelh or environment_variants or kpz2d or bc_even or random_systems" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
for V in "A=1" "CUPSS_B200_NO_XS1=1" "CUPSS_B200_NO_FANOUT=1" "CUPSS_B200_NO_XS1=1 CUPSS_B200_NO_FANOUT=1"; do
  echo "== $V"
  env $V timeout 300 python tools/bench_configs.py --only modelh --steps 200 2> gpurun_out/${TAG}_cfg.err | tee -a gpurun_out/${TAG}_modelh.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(round(d['steps_per_s'],1), {k:(v['ms'],v['launches']) for k,v in d['per_kernel'].items()})"
done
