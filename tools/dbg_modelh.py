"""GPU debug: Model H (and a two-group system) on long-axis shapes against ORACLE-F, per field."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from cases import CASES, ORACLE_F, rel_l2
two = dict(dt=0.01, fields=[("u", 1), ("v", 0)], params={}, eqs=["dt u + 0.5*q^2*u = -0.5*iqx*u^2 - v*u", "v = iqy*u"], ic=dict(u=("smooth", (0.5, 0.05))), steps=3, threads=0)
for name, base, shape in [("modelh", CASES["modelh_32"], (256, 2048, 1)), ("modelh", CASES["modelh_32"], (2048, 2048, 1))]:
    case = dict(base); case["shape"] = shape; case["steps"] = 3; case["threads"] = 0
    got = cases.run_case(case)
    want = cases.run_case(case, lib=ORACLE_F, device=0)
    print(name, shape, os.environ.get("CUPSS_B200_JIT"), {f: "%.1e" % rel_l2(got[f], want[f]) for f, _ in case["fields"]}, flush=True)
