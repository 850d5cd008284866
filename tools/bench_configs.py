#!/usr/bin/env python
"""Secondary measurements on one B200 (the headline metric lives in bench.py):

  * the other BASELINE.json configurations that fit one GPU -- CH-2D 4096^2 (cfg 02), Model H 2048^2 (cfg 04) and the 3-D KPZ
    system of cfg 06 at 512^3 with noise on -- steps/s, per-launch breakdown, algorithmic GB/s;
  * for context, the reference's OWN GPU path (unmodified sources, cuFFT + cuRAND, oracle/_ref/libcupss_ref_gpu.so) on
    CH-3D 512^3 and CH-2D 4096^2, timed with the same harness.

One JSON line per measurement.  Usage: python tools/bench_configs.py [--only ch2d,modelh,kpz,ref3d,ref2d] [--steps K]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from cupss_b200.capi import Evolver, RUN_GPU  # noqa: E402

REF_GPU = os.path.join(ROOT, "oracle", "_ref", "libcupss_ref_gpu.so")
PEAK = 6556.5
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", PEAK))


def white(shape, amp, seed=1324):
    rng = np.random.default_rng(seed)
    return (amp * (2.0 * rng.random(shape, dtype=np.float32) - 1.0)).astype(np.float32)


def system(name, lib=None):
    if name == "ch3d":
        n = 512
        ev = Evolver(RUN_GPU, n, n, n, 1.0, 1.0, 1.0, 0.01, lib=lib)
        ev.createField("phi", True)
        for k, v in cases.CH_PARAMS.items():
            ev.addParameter(k, v)
        ev.addEquation("dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 ")
        ev.setReal("phi", white((n, n, n), 0.01))
        return ev, n ** 3, "examples/03_cahn_hilliard_3d 512^3"
    if name == "ch2d":
        n = 4096
        ev = Evolver(RUN_GPU, n, n, 1, 1.0, 1.0, 1.0, 0.1, lib=lib)
        ev.createField("phi", True)
        for k, v in cases.CH_PARAMS.items():
            ev.addParameter(k, v)
        ev.addEquation("dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3")
        ev.setReal("phi", white((1, n, n), 0.1))
        return ev, n * n, "examples/02_cahn_hilliard 4096^2 deterministic"
    if name == "modelh":
        n = 2048
        ev = Evolver(RUN_GPU, n, n, 1, 1.0, 1.0, 1.0, 0.1, lib=lib)
        for f, d in cases.MODELH_FIELDS:
            ev.createField(f, d)
        for k, v in cases.MODELH_PARAMS.items():
            ev.addParameter(k, v)
        for e in cases.MODELH_EQS:
            ev.addEquation(e)
        ev.setReal("phi", white((1, n, n), 0.1))
        return ev, n * n, "examples/04_model_h 2048^2 (9 fields)"
    if name == "kpz":
        n = 512
        ev = Evolver(RUN_GPU, n, n, n, 1.0, 1.0, 1.0, 0.01, lib=lib)
        for f, d in [("h", 1), ("iqxh", 0), ("iqyh", 0), ("iqzh", 0)]:
            ev.createField(f, d)
        ev.addParameter("D", 0.5)
        ev.addParameter("l", 0.5)
        for e in ["dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2 + l*iqzh^2", "iqxh = iqx*h", "iqyh = iqy*h", "iqzh = iqz*h"]:
            ev.addEquation(e)
        ev.addNoise("h", "2*D")
        if lib is None:
            ev.setNoiseSeed(1234)
        return ev, n ** 3, "examples/06_kpz as a 3-D system, 512^3, noise on (1 GPU; cfg 06 is 1024^3 on 8)"
    raise ValueError(name)


def run_product(name, steps):
    ev, npts, what = system(name)
    ev.prepareProblem()
    ev.advanceTime(10)
    ev.sync()
    ms = ev.timeSteps(steps)
    value = steps / (ms * 1e-3)
    prof = {}
    reps = 3
    for _ in range(reps):
        for nm, t_ms, by in ev.profileStep():
            a = prof.setdefault(nm, [0.0, 0.0, 0])
            a[0] += t_ms; a[1] += by; a[2] += 1
    kern = {k: {"ms": round(v[0] / reps, 4), "launches": v[2] // reps, "GBps": round(v[1] / max(v[0], 1e-9) / 1e6, 1)} for k, v in prof.items() if k != "bump"}
    sb = ev.bytesPerStep()
    line = {"what": what, "impl": "b200", "steps_per_s": value, "ms_per_step": ms / steps, "grid_point_steps_per_s": value * npts,
            "algorithmic_bytes_per_step": sb, "bytes_per_point_step": sb / npts, "step_GBps": sb * value / 1e9, "step_frac_of_measured_hbm": sb * value / 1e9 / PEAK,
            "launches_per_step": ev.launchesPerStep(), "per_kernel": kern}
    ev.close()
    print(json.dumps(line), flush=True)


def run_reference_gpu(name, steps):
    if not os.path.exists(REF_GPU):
        print(json.dumps({"what": name, "impl": "reference-cufft", "unavailable": "oracle/_ref/libcupss_ref_gpu.so not built"}), flush=True)
        return
    import torch
    ev, npts, what = system(name, lib=REF_GPU)
    ev.prepareProblem()
    ev.advanceTime(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev.advanceTime(steps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    ev.close()
    print(json.dumps({"what": what, "impl": "reference-cufft (unmodified reference GPU path, cuFFT 11 / cuRAND, sm_100a build)", "steps_per_s": 1.0 / dt,
                      "ms_per_step": 1e3 * dt, "grid_point_steps_per_s": npts / dt, "steps_timed": steps}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="ch2d,modelh,kpz,ref3d,ref2d")
    ap.add_argument("--steps", type=int, default=100)
    args = ap.parse_args()
    import torch
    assert torch.cuda.is_available(), "needs a GPU"
    torch.cuda.set_device(0)
    for item in args.only.split(","):
        if item in ("ch2d", "modelh", "kpz", "ch3d"):
            run_product(item, args.steps)
        elif item == "ref3d":
            run_reference_gpu("ch3d", 10)
        elif item == "ref2d":
            run_reference_gpu("ch2d", 20)


if __name__ == "__main__":
    main()
