"""Regenerates tests/golden/ from the reference tree (run in the build container, where /root/reference exists).

* base_truths/  -- the reference's own golden vectors, copied verbatim from /root/reference/tests/base_truths
                   (data, not source; rows "i[, j[, k]], value", written by the reference with %.6f).
* ref_runs.npz  -- outputs of the compiled, unmodified reference CPU path (oracle/_ref/libcupss_ref_{u,f}.so) on
                   small seeded inputs, so that GPU-box tests can check the oracle build they carry and the product
                   against numbers that were produced HERE from the reference itself.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    src = "/root/reference/tests/base_truths"
    dst = os.path.join(HERE, "base_truths")
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(src)):
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    from cases import CASES, run_case, ORACLE_F, ORACLE_U
    out = {}
    for name, case in CASES.items():
        if int(np.prod(case["shape"])) > 65536:
            continue   # the long-line cases are checked against the oracle build directly; the fixture stays small
        lib = ORACLE_U if case.get("oracle") == "U" else ORACLE_F
        res = run_case(case, lib=lib, device=0)
        for field, arr in res.items():
            out[f"{name}/{field}"] = arr
        print(name, {k: float(np.linalg.norm(v)) for k, v in res.items()})
    np.savez_compressed(os.path.join(HERE, "ref_runs.npz"), **out)


if __name__ == "__main__":
    main()
