"""Float32 floor of the Model H 256^2 x 100-step parity case (tests/cases.py: modelh_256["tol"]).

Runs, on the CPU, the compiled reference (ORACLE-F: float32 arithmetic, FFT shim with double precision inside) and the
numpy restatement in float64 and in float32 (complex64 FFTs) from the same initial condition and prints the relative L2
distance per field.  Derived fields that are differences of nearly equal terms (vx, vy, w) cannot agree to 1e-5 between ANY
two float32 pipelines; the GPU parity test allows 1.5 x (float32 restatement vs float64) for them.

    python tests/golden/f32_floor.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from cases import CASES, ORACLE_F  # noqa: E402
from oracle.restatement import from_plan_dump  # noqa: E402


def rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main(name="modelh_256"):
    case = dict(CASES[name])
    ev = cases.build_system(case, lib=ORACLE_F, device=0)
    ics = {n: ev.real(n) for n, _ in case["fields"]}
    ev.prepareProblem()
    out = {}
    for dt in (np.float64, np.float32):
        s = from_plan_dump(ev.dumpPlan(), case["shape"], (1.0, 1.0, 1.0), case["dt"], dt, "gpu")
        for n, a in ics.items():
            s.real[n] = a.astype(dt)
        s.prepare()
        s.step(case["steps"])
        out[dt] = s
    ev.advanceTime(case["steps"])
    print("%-8s %-10s %-22s %-22s %-22s" % ("field", "rms", "reference f32 vs f64", "numpy f32 vs f64", "reference vs numpy f32"))
    for n, _ in case["fields"]:
        r = ev.real(n)
        print("%-8s %-10.3e %-22.2e %-22.2e %-22.2e" % (n, np.linalg.norm(r) / r.size ** 0.5, rel(r, out[np.float64].real[n]),
                                                      rel(out[np.float32].real[n], out[np.float64].real[n]), rel(r, out[np.float32].real[n])))
    ev.close()


if __name__ == "__main__":
    main(*sys.argv[1:])
