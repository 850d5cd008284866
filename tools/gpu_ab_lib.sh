#!/bin/bash
# A/B of two builds of the engine library on the headline bench: the in-tree build, then <variant .so> copied over it.
TAG=$1; VAR=$2
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-context --no-extra > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["value"],1), "steps/s", {k:v["ms"] for k,v in d["roofline"]["per_kernel"].items()}, "parity", d["parity"] and d["parity"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/${TAG}_$1.err").read()[-1500:])
PY
}
run base0
cp cupss_b200/lib/libcupss_b200.so /tmp/base.so
cp $VAR cupss_b200/lib/libcupss_b200.so
run var0
cp /tmp/base.so cupss_b200/lib/libcupss_b200.so
run base1
cp $VAR cupss_b200/lib/libcupss_b200.so
run var1
