#!/bin/bash
# 1 GPU: full GPU suite + bench line with the secondary configurations (cluster kernels for long axes)
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${TAG}_pytest.log 2>&1; tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-context > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "steps/s", {k:v["ms"] for k,v in d["roofline"]["per_kernel"].items()}, "parity", d["parity"]["ok"], "e2e", round(d["e2e"]["value"],1))
for k,v in (d.get("extra_configs") or {}).items():
    print(k, round(v.get("steps_per_s",0),1), {kk:(vv["ms"],vv["frac"]) for kk,vv in v.get("per_kernel",{}).items()}, v.get("parity"), v.get("error"))
PY
tail -5 gpurun_out/${TAG}_bench.err
