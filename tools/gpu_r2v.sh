#!/bin/bash
# Model H 2048^2 A/B of the one-job stash x pass variants (threads per CTA, twiddles from global memory, line buffers)
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed or modelh_2048 or environment_variants" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
for V in "A=1" "CUPSS_B200_XS_OB=1" "CUPSS_B200_XS_TWG=0" "CUPSS_B200_XS_TWG=1" "CUPSS_B200_XS_NT=128" "CUPSS_B200_XS_OB=1 CUPSS_B200_XS_TWG=1"; do
  echo "== $V"
  env $V timeout 300 python tools/bench_configs.py --only modelh --steps 200 2> gpurun_out/${TAG}_cfg.err | tee -a gpurun_out/${TAG}_modelh.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(round(d['steps_per_s'],1), {k:(v['ms'],v['launches']) for k,v in d['per_kernel'].items()})"
done
