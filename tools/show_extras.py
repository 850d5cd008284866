import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"], 1), "steps/s", {k: v["ms"] for k, v in d["roofline"]["per_kernel"].items()}, "parity", d["parity"] and d["parity"]["ok"], "e2e", d["e2e"] and round(d["e2e"]["value"], 1))
for k, v in (d.get("extra_configs") or {}).items():
    print(k, round(v.get("steps_per_s", 0), 1), {kk: (vv["ms"], vv["frac"]) for kk, vv in v.get("per_kernel", {}).items()}, (v.get("parity") or {}).get("ok"), (v.get("parity") or {}).get("rel_l2"), v.get("error"))
