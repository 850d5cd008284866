#!/bin/bash
# Round-2 experiment call (1 GPU): TMA tile prologue + DIF inverse: tests, A/B, serial z-chunks.
TAG=${1:-r2d}
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
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
i=0
for V in "CUPSS_B200_TMA=1" "CUPSS_B200_TMA=0" "CUPSS_B200_TMA=1" "CUPSS_B200_TMA=0" "CUPSS_B200_ZCHUNK=32 CUPSS_B200_ZCHUNK_LANES=1" "CUPSS_B200_ZCHUNK=64 CUPSS_B200_ZCHUNK_LANES=1" "CUPSS_B200_ZCHUNK=128 CUPSS_B200_ZCHUNK_LANES=1"; do
  env $V timeout 300 python bench.py $B > gpurun_out/${TAG}_ab$i.json 2> gpurun_out/${TAG}_ab$i.err
  show gpurun_out/${TAG}_ab$i.json "$V"
  i=$((i+1))
done
