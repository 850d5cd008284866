#!/usr/bin/env python
"""bench.py -- timesteps/s of Cahn-Hilliard 3-D 512^3 (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            own arm (CUDA engine through the evolver API)
  python bench.py --impl reference ...                     the reference's own CPU implementation (oracle/_ref), SAME grid

One JSON line on rank 0.
  value         steps/s with the state resident in HBM (CUDA events on the engine's stream, max over ranks)
  parity        BEFORE the timed loop, at every N: the seeded 12-step job of tests/golden/make_golden_fullsize.py on the
                512^3 grid, every rank checking its own z-slab against tests/golden/ch3d_512_ref.npz (the reference's CPU
                path run in the build container); relative L2 > 1e-5 fails the run (exit code 1)
  e2e           the same job through the public API from HOST buffers (upload of the initial condition, K steps,
                download of the result, all inside the timed region)
  roofline      the dominant kernel's algorithmic bytes / its event-timed duration against MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference CPU path at the SAME 512^3 grid (a bounded number of steps), FFT shim threaded
  context       the reference's own cuFFT/cuRAND GPU build timed on the same box (N = 1)
  extra_configs the other BASELINE.json configurations (02 at 4096^2, 04 at 2048^2, 06 at 512^3 on 1 GPU / 1024^3 on 8):
                steps/s, per-kernel GB/s, a parity witness
Strong scaling: the 512^3 grid is fixed and slab-partitioned over the ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EQ = "dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "       # examples/03_cahn_hilliard_3d/03_cahn_hilliard_3d.cpp:13-23
PARAMS = (("a", -1.0), ("b", 1.0), ("k", 4.0))
DT = 0.01
FALLBACK_HBM = 6650.0   # GB/s, /opt/skills/guides/B200_PROFILING.md
NVLINK_NOMINAL, NVLINK_MEASURED = 900.0, 706.0   # GB/s per direction per GPU: nominal / peer STORES measured on this pool (tools/ubench/p2p_push.cu, profiles/r2b_p2p_push.txt)
PARITY_TOL = 1e-5       # BASELINE.json north_star
GOLDEN_512 = os.path.join(ROOT, "tests", "golden", "ch3d_512_ref.npz")
MODELH_FIELDS = [("phi", 1), ("iqxphi", 0), ("iqyphi", 0), ("sigxx", 0), ("sigxy", 0), ("vx", 0), ("vy", 0), ("w", 0), ("P", 0)]
MODELH_PARAMS = dict(a=-1, b=1, k=4, eta=1, friction=0, ka=4)
MODELH_EQS = [   # examples/04_model_h/modelh_base.cpp:33-43
    "dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 -vx*iqxphi - vy*iqyphi", "iqxphi = iqx*phi", "iqyphi = iqy*phi",
    "sigxx = - 0.5*ka *iqxphi * iqxphi + 0.5*ka*iqyphi*iqyphi", "sigxy = - ka *iqxphi * iqyphi",
    "-q^2*P = (iqx*iqx-iqy*iqy)*sigxx + 2.0 * iqx*iqy*sigxy", "vx * (friction + eta*q^2) = -iqx*P + iqx*sigxx + iqy*sigxy",
    "vy * (friction + eta*q^2) = -iqy*P + iqx*sigxy - iqy*sigxx", "w = 0.5*iqx * vy - 0.5*iqy*vx "]
KPZ_FIELDS = [("h", 1), ("iqxh", 0), ("iqyh", 0), ("iqzh", 0)]
KPZ_EQS = ["dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2 + l*iqzh^2", "iqxh = iqx*h", "iqyh = iqy*h", "iqzh = iqz*h"]

CONFIG = "ch3d"   # --config: "ch3d" (the headline, BASELINE.json configs[2]) or "kpz3d" (configs[4]: 3-D KPZ with noise)


def make_named_system(Evolver, dev, name, n, lib=None, noise=True):
    """The systems of BASELINE.json's configs.  name: ch3d (03), ch2d (02), modelh (04), kpz3d (06 lifted to 3-D)."""
    if name == "ch3d":
        ev = Evolver(dev, n, n, n, 1.0, 1.0, 1.0, DT, lib=lib)
        ev.createField("phi", True)
        for k, v in PARAMS:
            ev.addParameter(k, v)
        ev.addEquation(EQ)
        return ev
    if name == "ch2d":   # examples/02_cahn_hilliard/02_cahn-hilliard.cpp:17-35
        ev = Evolver(dev, n, n, 1, 1.0, 1.0, 1.0, 0.1, lib=lib)
        ev.createField("phi", True)
        for k, v in PARAMS:
            ev.addParameter(k, v)
        ev.addEquation("dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3")
        return ev
    if name == "modelh":   # examples/04_model_h/modelh_base.cpp:16-54
        ev = Evolver(dev, n, n, 1, 1.0, 1.0, 1.0, 0.1, lib=lib)
        for f, d in MODELH_FIELDS:
            ev.createField(f, d)
        for k, v in MODELH_PARAMS.items():
            ev.addParameter(k, v)
        for e in MODELH_EQS:
            ev.addEquation(e)
        return ev
    if name == "kpz3d":   # examples/06_kpz lifted to 3-D: h + three gradient constraint fields, white noise on h
        ev = Evolver(dev, n, n, n, 1.0, 1.0, 1.0, DT, lib=lib)
        for f, d in KPZ_FIELDS:
            ev.createField(f, d)
        ev.addParameter("D", 0.5)
        ev.addParameter("l", 0.5)
        for e in KPZ_EQS:
            ev.addEquation(e)
        if noise:
            ev.addNoise("h", "2*D")
            if lib is None:
                ev.setNoiseSeed(1234)
        return ev
    raise ValueError(name)


def make_system(Evolver, dev, n, lib=None):
    return make_named_system(Evolver, dev, CONFIG, n, lib=lib)


def main_field():
    return "h" if CONFIG == "kpz3d" else "phi"


def synthetic_ic(n, seed=1324):
    rng = np.random.default_rng(seed)
    return (0.01 * (2.0 * rng.random((n, n, n), dtype=np.float32) - 1.0)).astype(np.float32)


def smooth_ic(sx, sy, sz, amp=0.5, noise=0.05, seed=1, kmult=1):
    """tests/cases.py::smooth_ic evaluated by broadcasting instead of three full meshgrids (same IEEE operations per
    element, hence the same bits -- tests/test_host.py checks it): the parity initial condition at 512^3 without 3 GiB of
    index arrays."""
    rng = np.random.default_rng(seed)
    x = np.arange(sx).reshape(1, 1, sx)
    y = np.arange(sy).reshape(1, sy, 1)
    z = np.arange(sz).reshape(sz, 1, 1)
    f = amp * np.sin(2 * np.pi * (2 * kmult) * x / sx)
    if sy > 1:
        f = f * np.cos(2 * np.pi * (3 * kmult) * y / sy)
    if sz > 1:
        f = f * np.cos(2 * np.pi * kmult * z / sz)
    return (f + noise * (2 * rng.random((sz, sy, sx)) - 1)).astype(np.float32)


class ClockSampler:
    """SM clock + throttle reasons WHILE the timed region runs (B200_PROFILING.md clocks line).

    NVML is polled from a thread of this process every few ms (the engine's C calls release the GIL), so even a
    0.2 s timed region gets tens of samples; `nvidia-smi -lms` (whose start-up alone can outlast a short region) is
    only the fallback when the NVML binding is missing."""

    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index, period_s=0.025):
        self.index, self.proc, self.lines, self.period = index, None, [], period_s
        self.nv, self.handle, self.thread, self.run = None, None, None, False
        self.sm, self.mx, self.reasons = [], [], set()
        self.samples = []   # (perf_counter, sm_mhz, reason bits): filtered to the timed window in stop()

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip() != ""]
        if ids and all(v.strip().isdigit() for v in ids) and self.index < len(ids):
            return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.handle = nv.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)))
            self.nv, self.run = nv, True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = ((nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
        while self.run:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.samples.append((time.perf_counter(), mhz, tuple(nm for b, nm in bits if r & b)))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, t_begin=None, t_end=None):
        if self.nv is not None:
            self.run = False
            self.thread.join(timeout=1.0)
            inside = [x for x in self.samples if (t_begin is None or x[0] >= t_begin) and (t_end is None or x[0] <= t_end)]
            for _, mhz, rs in (inside or self.samples):
                self.sm.append(mhz)
                self.reasons.update(rs)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": f"nvml, polled every {self.period * 1e3:.0f} ms inside the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------- reference CPU path
def host_ram_gb():
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2 ** 30
    except (ValueError, OSError):
        return 0.0


def _shim_threads(lib_path, threads):
    import ctypes as C
    shim = C.CDLL(lib_path)
    if hasattr(shim, "cupss_shim_set_threads"):
        shim.cupss_shim_set_threads(int(threads))
        return True
    return False


def cpu_reference_steps(n, warm, steps, threads, extra_one_core_steps=0):
    """The UNMODIFIED reference CPU path (oracle/_ref/libcupss_ref_u.so = its sources where they lie + the FFTW-API shim) on the
    benchmark's own n^3 Cahn-Hilliard grid: `warm` untimed steps, then `steps` steps timed one by one, then optionally steps
    with the shim's FFT on ONE thread (the reference CPU path itself is serial).  Returns (list of seconds, list of 1-core seconds)."""
    from cupss_b200.capi import Evolver, RUN_CPU
    lib = os.path.join(ROOT, "oracle", "_ref", "libcupss_ref_u.so")
    if not os.path.exists(lib):
        raise FileNotFoundError("oracle/_ref/libcupss_ref_u.so missing: run __graft_entry__.build() where /root/reference exists")
    _shim_threads(lib, threads)
    ev = make_named_system(Evolver, RUN_CPU, "ch3d", n, lib=lib)
    ev.setReal("phi", synthetic_ic(n))
    ev.prepareProblem()
    if warm > 0:
        ev.advanceTime(warm)
    per = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ev.advanceTime(1)
        per.append(time.perf_counter() - t0)
    one = []
    if extra_one_core_steps > 0:
        _shim_threads(lib, 1)
        for _ in range(extra_one_core_steps):
            t0 = time.perf_counter()
            ev.advanceTime(1)
            one.append(time.perf_counter() - t0)
        _shim_threads(lib, threads)
    ev.close()
    return per, one


def reference_grid(n):
    """The reference arm runs the NAMED grid.  Only a host that cannot hold it (the reference keeps ~150 B per grid point
    between its host arrays and the host-memory stand-ins of its device arrays: ~20 GB at 512^3) falls back, and says so."""
    need_gb = 170.0 * n ** 3 / 2 ** 30
    ram = host_ram_gb()
    if ram and ram < need_gb * 1.15:
        m = n
        while m > 64 and 170.0 * m ** 3 / 2 ** 30 * 1.15 > ram:
            m //= 2
        return m, f"host RAM {ram:.0f} GiB < {need_gb:.0f} GiB needed at {n}^3: reduced to {m}^3"
    return n, None


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the step on the SAME grid as the product arm, all host
    threads the FFT shim can use, W warm-up steps and exactly K timed steps (one advanceTime each)."""
    if rank != 0:
        return
    n, why = reference_grid(args.size)
    cores = os.cpu_count() or 1
    per, one = cpu_reference_steps(n, args.warmup, args.steps, cores, extra_one_core_steps=1)
    sec = float(np.mean(per))
    value = 1.0 / sec
    sample = (f"CH-3D {n}^3 (the named grid), {args.steps} timed advanceTime calls after {args.warmup} warm-up steps, unmodified reference sources, "
              f"FFTW-API shim threaded over {cores} host threads (pointwise passes are serial, as in the reference)")
    if why:
        sample += "; " + why
    workload = f"examples/03_cahn_hilliard_3d {n}^3, dt={DT}, a=-1 b=1 k=4, deterministic, IC 0.01*(2u-1)"
    line = {"impl": "reference", "metric": "timesteps/s, Cahn-Hilliard 3D 512^3", "value": value, "unit": "steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "partition": "host cores of the box (reference CPU path, FFT shim threaded)", "l2": "n/a (CPU)",
                       "grid_point_steps_per_s": value * n ** 3},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "reference", "sample": sample,
                             "one_core": {"value": 1.0 / float(np.mean(one)), "unit": "steps/s", "cores": 1,
                                          "sample": f"one more step of the same run with the shim on 1 thread (the reference CPU path as shipped is serial)"} if one else None,
                             "seconds_per_step_min_max": [float(np.min(per)), float(np.max(per))]},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- parity witness
def parity_witness(ev, rank, world, dist, n):
    """The seeded 12-step job of tests/golden/make_golden_fullsize.py on the benchmark grid, through the same evolver object
    (and partition) that is timed afterwards.  Every rank compares its z-slab with the committed summary of the reference's
    CPU path (ORACLE-F): 32^3 point samples, three decimated planes, whole-field statistics."""
    import torch
    gold = np.load(GOLDEN_512)
    steps = int(gold["steps"][0])
    zl = n // world
    z0, z1 = rank * zl, (rank + 1) * zl
    ic = smooth_ic(n, n, n, 0.5, 0.05)
    ev.setReal("phi", ic)
    del ic
    ev.prepareProblem()
    ev.advanceTime(steps)
    ev.copyAllDataToHost()
    phi = ev.fieldReal("phi")[z0:z1, :, :, 0]   # this rank's slab (view of the host mirror)
    s, d = n // 32, n // 128
    zi, yi, xi = (37 * n) // 512, (201 * n) // 512, (333 * n) // 512
    acc = {}

    def add(key, got, want):
        a = acc.setdefault(key, [0.0, 0.0])
        g, w = np.asarray(got, np.float64), np.asarray(want, np.float64)
        a[0] += float(((g - w) ** 2).sum()); a[1] += float((w ** 2).sum())

    zs = [z for z in range(z0, z1) if z % s == 0]
    if zs:
        add("sub", phi[[z - z0 for z in zs]][:, ::s, ::s], gold["sub"][[z // s for z in zs]])
    zd = [z for z in range(z0, z1) if z % d == 0]
    if zd:
        loc, glo = [z - z0 for z in zd], [z // d for z in zd]
        add("planes", phi[loc][:, yi, ::d], gold["plane_y"][glo])
        add("planes", phi[loc][:, ::d, xi], gold["plane_x"][glo])
    if z0 <= zi < z1:
        add("planes", phi[zi - z0, ::d, ::d], gold["plane_z"])
    p64 = phi.astype(np.float64)
    sums = [acc.get("sub", [0, 0])[0], acc.get("sub", [0, 0])[1], acc.get("planes", [0, 0])[0], acc.get("planes", [0, 0])[1],
            float(p64.sum()), float((p64 ** 2).sum()), float((np.abs(p64) ** 3).sum())]
    mn, mx = float(p64.min()), float(p64.max())
    finite = bool(np.isfinite(p64).all())
    del p64
    if dist is not None:
        t = torch.tensor(sums, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        sums = [float(v) for v in t.cpu()]
        t = torch.tensor([mn, -mx, 1.0 if finite else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)   # min, -max, AND of the finite flags
        mn, mx, finite = float(t[0].item()), -float(t[1].item()), float(t[2].item()) == 1.0
    want = gold["stats"]   # mean, L2 norm, min, max, sum |phi|^3
    rel_sub = float(np.sqrt(sums[0] / sums[1])) if sums[1] > 0 else float("nan")
    rel_pl = float(np.sqrt(sums[2] / sums[3])) if sums[3] > 0 else float("nan")
    mean, l2, a3 = sums[4] / n ** 3, float(np.sqrt(sums[5])), sums[6]
    stats_err = {"mean_abs": abs(mean - want[0]), "l2_rel": abs(l2 - want[1]) / abs(want[1]), "abs3_rel": abs(a3 - want[4]) / abs(want[4]),
                 "min_rel": abs(mn - want[2]) / abs(want[2]), "max_rel": abs(mx - want[3]) / abs(want[3])}
    ok = (finite and rel_sub < PARITY_TOL and rel_pl < PARITY_TOL and stats_err["mean_abs"] < 2e-7 + 1e-5 * abs(want[0]) and
          stats_err["l2_rel"] < 1e-5 and stats_err["abs3_rel"] < 1e-5 and stats_err["min_rel"] < 1e-4 and stats_err["max_rel"] < 1e-4)
    return {"ok": bool(ok), "tol": PARITY_TOL, "steps": steps, "rel_l2_sub": rel_sub, "rel_l2_planes": rel_pl, "stats_err": stats_err,
            "ranks_checked": world, "against": "tests/golden/ch3d_512_ref.npz: the reference CPU path (ORACLE-F) at 512^3, smooth seeded IC, "
            "32^3 samples + 3 planes + whole-field statistics; every rank checks its own z-slab"}


# ---------------------------------------------------------------------------------------------- secondary configurations
def profile_kernels(ev, reps):
    prof = {}
    for _ in range(reps):
        for name, t_ms, by in ev.profileStep():
            a = prof.setdefault(name, [0.0, 0.0, 0])
            a[0] += t_ms; a[1] += by; a[2] += 1
    return {k: {"ms": v[0] / reps, "bytes": v[1] / reps, "launches_per_step": v[2] / reps} for k, v in prof.items() if k != "bump"}


def quick_parity(name, n, steps, Evolver):
    """Product vs the compiled reference CPU path (ORACLE-F) on the NAMED grid of a secondary configuration for a few steps
    from the seeded smooth initial condition (100-step runs at reduced sizes live in tests/test_gpu_parity.py)."""
    from cupss_b200.capi import RUN_CPU, RUN_GPU
    lib = os.path.join(ROOT, "oracle", "_ref", "libcupss_ref_f.so")
    if not os.path.exists(lib):
        return {"ok": None, "why": "oracle/_ref/libcupss_ref_f.so not built"}
    _shim_threads(lib, os.cpu_count() or 1)
    sz = n if name in ("ch3d", "kpz3d") else 1
    fld = "h" if name == "kpz3d" else "phi"
    # structure at the scale of the 32 / 64-point parity cases (wavenumbers scaled with the grid): derived fields (gradients,
    # stresses, velocities) are then compared at their natural scale, not at the float32 round-off floor of a 2-period pattern
    amp = (1.0, 0.1) if name == "kpz3d" else ((0.5, 0.025) if name == "modelh" else (0.4, 0.04))
    ic = smooth_ic(n, n, sz, *amp, 1, max(1, n // 64))
    out = {}
    for tag, lib_, dev in (("got", None, RUN_GPU), ("want", lib, RUN_CPU)):
        ev = make_named_system(Evolver, dev, name, n, lib=lib_, noise=False)
        ev.setReal(fld, ic)
        ev.prepareProblem()
        ev.advanceTime(steps)
        if dev == RUN_GPU:
            ev.copyAllDataToHost()
        fields = [f for f, _ in (MODELH_FIELDS if name == "modelh" else (KPZ_FIELDS if name == "kpz3d" else [("phi", 1)]))]
        out[tag] = {f: ev.real(f) for f in fields}
        ev.close()
    errs = {}
    for f in out["got"]:
        w = out["want"][f].astype(np.float64)
        nw = float(np.linalg.norm(w))
        errs[f] = float(np.linalg.norm(out["got"][f].astype(np.float64) - w) / (nw if nw > 0 else 1.0))
    worst = max(errs.values())
    # Model H: the projected velocity and its curl are differences of nearly equal terms; no float32 pipeline agrees with another
    # to 1e-5 on them (tests/golden/f32_floor.py: numpy float32 vs float64 2.4e-5 / 4.0e-5 / 1.4e-4 on vx / vy / w, 1e-5 on the
    # stresses) -- same per-field tolerances as tests/cases.py::modelh_256; the dynamic field and its gradients keep 1e-5
    tols = dict(sigxx=1.5e-5, sigxy=1.5e-5, P=1.5e-5, vx=3.6e-5, vy=6e-5, w=2.1e-4) if name == "modelh" else {}
    ok = all(e < tols.get(f, PARITY_TOL) for f, e in errs.items())
    return {"ok": bool(ok), "steps": steps, "grid": n, "rel_l2_max": worst, "rel_l2": errs, "tol": PARITY_TOL, "tol_per_field": tols or None,
            "against": "oracle/_ref/libcupss_ref_f.so (reference CPU path with the GPU kernels' semantics) on the same grid and input"}


def run_extra(name, n, steps, peak, Evolver, parity_steps, rank=0, world=1, dist=None, uid_fn=None):
    """One secondary configuration: steps/s (HBM-resident, CUDA events), per-kernel GB/s, parity witness."""
    import torch
    from cupss_b200.capi import RUN_GPU
    what = {"ch2d": f"examples/02_cahn_hilliard {n}^2 deterministic", "modelh": f"examples/04_model_h {n}^2 (9 fields)",
            "kpz3d": f"examples/06_kpz as a 3-D system {n}^3, noise 2*D on h, (grad h)^2 dealiased"}[name]
    ev = make_named_system(Evolver, RUN_GPU, name, n)
    if world > 1:
        ev.setPartition(rank, world, uid_fn())
    if name != "kpz3d":
        rng = np.random.default_rng(1324)
        ev.setReal("phi", (0.1 * (2.0 * rng.random((1, n, n), dtype=np.float32) - 1.0)).astype(np.float32))
    ev.prepareProblem()
    ev.advanceTime(10)
    ev.sync()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = ev.timeSteps(steps)
    if dist is not None:
        tt = torch.tensor([ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = steps / (ms * 1e-3)
    kern = profile_kernels(ev, 3)
    sb = ev.bytesPerStep()
    npts = n ** 3 if name == "kpz3d" else n * n
    finite = True
    if name == "kpz3d" and world == 1:
        ev.copyAllDataToHost()
        finite = bool(np.isfinite(ev.fieldReal("h")[..., 0]).all())
    line = {"workload": what, "n_gpus": world, "steps_per_s": value, "ms_per_step": ms / steps, "grid_point_steps_per_s": value * npts,
            "algorithmic_bytes_per_step": sb, "bytes_per_point_step": sb * world / npts, "step_GBps_per_gpu": sb * value / 1e9,
            "step_frac_of_measured_hbm": sb * value / 1e9 / peak, "launches_per_step": ev.launchesPerStep(),
            "per_kernel": {k: {"ms": round(v["ms"], 4), "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                               "frac": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peak, 3)} for k, v in kern.items()}}
    ev.close()
    if parity_steps > 0 and world == 1 and name in ("ch2d", "modelh"):
        line["parity"] = quick_parity(name, n, parity_steps, Evolver)
    elif name == "kpz3d":
        line["parity"] = {"ok": finite if world == 1 else None, "what": "field finite after the run; the stochastic system is checked statistically "
                          "(noise variance, per-mode spectrum incl. kz, Hermitian planes) and its deterministic part against the reference "
                          "in tests/test_gpu_parity.py"}
    return line


def reference_cufft_context(n, steps):
    """The reference's OWN GPU path (unmodified sources, cuFFT + cuRAND, built for sm_100a: oracle/_ref/libcupss_ref_gpu.so) on the
    benchmark configuration, same box, timed by the wall clock around advanceTime with a device synchronise on both sides."""
    import torch
    from cupss_b200.capi import Evolver, RUN_GPU
    lib = os.path.join(ROOT, "oracle", "_ref", "libcupss_ref_gpu.so")
    if not os.path.exists(lib):
        return {"ref_cufft_steps_per_s": None, "why": "oracle/_ref/libcupss_ref_gpu.so not built"}
    ev = make_named_system(Evolver, RUN_GPU, "ch3d", n, lib=lib)
    ev.setReal("phi", synthetic_ic(n))
    ev.prepareProblem()
    ev.advanceTime(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev.advanceTime(steps)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    ev.close()
    return {"ref_cufft_steps_per_s": 1.0 / sec, "ref_cufft_ms_per_step": 1e3 * sec, "steps_timed": steps,
            "what": f"unmodified reference GPU path (cuFFT / cuRAND, sm_100a build) on CH-3D {n}^3, same box, 1 GPU"}


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--config", default="ch3d", choices=["ch3d", "kpz3d"], help="ch3d = the headline metric; kpz3d = BASELINE.json configs[4] (secondary)")
    ap.add_argument("--cpu-steps", type=int, default=2, help="timed steps of the reference CPU path (same grid) for cpu_baseline")
    ap.add_argument("--clock-period-ms", type=float, default=25.0, help="NVML clock / throttle-reason polling period inside the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-context", action="store_true", help="skip timing the reference's cuFFT build")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configurations (extra_configs)")
    ap.add_argument("--extra-steps", type=int, default=50)
    ap.add_argument("--extra-parity-steps", type=int, default=5)
    args = ap.parse_args()

    global CONFIG
    CONFIG = args.config
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cupss_b200 import capi
    from cupss_b200.capi import Evolver, RUN_GPU
    import ctypes as C
    eng = capi.load_engine()

    def unique_id():
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            capi.engine_check(eng.cupss_b200_nccl_unique_id(idbuf), "nccl_unique_id")
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    n = args.size
    ev = make_system(Evolver, RUN_GPU, n)
    if world > 1:
        ev.setPartition(rank, world, unique_id())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity witness on the very evolver (and partition) that is timed below
    parity = None
    if CONFIG == "ch3d" and not args.no_parity:
        if n == 512 and os.path.exists(GOLDEN_512):
            parity = parity_witness(ev, rank, world, dist, n)
        else:
            parity = {"ok": None, "why": "the committed full-size witness exists for the 512^3 grid only"}

    ic = synthetic_ic(n) if CONFIG == "ch3d" else None   # kpz3d starts from h = 0: the (zero) host mirrors are left untouched
    if ic is not None:
        ev.setReal(main_field(), ic)
    ev.prepareProblem()

    # ---- warm-up (graph capture happens here), then K timed steps; inputs (4 arrays x 0.55 GB) exceed L2
    ev.advanceTime(max(3, args.warmup))
    ev.sync()
    # the sampler thread (NVML start-up takes milliseconds) is started BEFORE the barrier: every rank must enter the timed
    # loop at the same moment, or the late rank's delay is charged to its peers at the first exchange (max over ranks)
    sampler = ClockSampler(local, args.clock_period_ms * 1e-3)
    if rank == 0:
        sampler.start()
    barrier()
    t_begin = time.perf_counter()
    ms = ev.timeSteps(args.steps)
    t_end = time.perf_counter()
    barrier()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    if dist is not None:
        tt = torch.tensor([ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = args.steps / (ms * 1e-3)

    # ---- per-launch breakdown (events between launches, graph bypassed) for the roofline of the dominant kernel
    kernels = profile_kernels(ev, 5)
    barrier()
    top = max((k for k in kernels if not k.startswith(("a2a", "xbar"))), key=lambda k: kernels[k]["ms"])
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", FALLBACK_HBM))
    per_launch_ms = kernels[top]["ms"] / kernels[top]["launches_per_step"]
    per_launch_bytes = kernels[top]["bytes"] / kernels[top]["launches_per_step"]
    achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")   # dram__bytes_read+write per launch from the committed ncu --set full capture
    if world == 1 and os.path.exists(tr_path):
        traffic = json.load(open(tr_path)).get(f"{n}", {}).get(top)
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": per_launch_bytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "per_kernel": {k: {"ms": round(v["ms"], 4), "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                                   "frac": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peak, 3)} for k, v in kernels.items()},
                "step_bytes": ev.bytesPerStep(), "step_GBps": ev.bytesPerStep() * value / 1e9,
                "step_frac_of_peak": ev.bytesPerStep() * value / 1e9 / peak}

    # ---- end to end through the public API from host buffers: upload IC, K steps, download the field
    e2e = None
    if not args.no_e2e:
        barrier()
        # The host buffer of the public API is the page-locked float2 mirror `fieldsReal[name]` (value in .x, the reference's
        # layout); the user's initial condition is written there before the clock starts (it is user code, a numpy strided
        # copy here).  Timed: prepareProblem (H2D upload + transform + plan), K x advanceTime, copyAllDataToHost (inverse
        # transform + D2H of the real AND the Fourier array, as the reference does), result read on the host.
        # One untimed pass of the same sequence first (3 steps): the first D2H of a process pays one-off driver set-up.
        if ic is not None:
            ev.setReal(main_field(), ic)
        ev.prepareProblem()
        ev.advanceTime(3)
        ev.copyAllDataToHost()
        t0 = time.perf_counter()
        if ic is not None:
            ev.setReal(main_field(), ic)
        barrier()
        t1 = time.perf_counter()
        ev.prepareProblem()
        t2 = time.perf_counter()
        ev.advanceTime(args.steps)
        ev.sync()
        t3 = time.perf_counter()
        ev.copyAllDataToHost()
        _ = float(ev.fieldReal(main_field())[0, 0, 0, 0])
        barrier()
        t4 = time.perf_counter()
        el = t4 - t1
        phases = {"untimed_fill_of_the_host_mirror_s": t1 - t0, "prepareProblem_h2d_s": t2 - t1, "steps_s": t3 - t2, "copyAllDataToHost_d2h_s": t4 - t3}
        if dist is not None:
            tt = torch.tensor([el], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            el = float(tt.item())
        slab_bytes = n * n * (n // world) * 8
        e2e = {"value": args.steps / el, "unit": "steps/s", "h2d_bytes_per_step": slab_bytes * world / args.steps,
               "d2h_bytes_per_step": 2 * slab_bytes * world / args.steps, "seconds": el, "phases": phases,
               "what": "host mirror filled -> prepareProblem (H2D) + K x advanceTime + copyAllDataToHost (D2H of real and Fourier arrays), wall clock, max over ranks"}

    launches = ev.launchesPerStep() * args.steps
    comm = ev.commBytesPerStep()
    ev.close()

    nvlink = None
    if world > 1:
        nvlink = {"bytes_per_step_per_gpu": comm, "GBps_per_direction_over_the_whole_step": comm * value / 1e9,
                  "peak_nominal": NVLINK_NOMINAL, "peak_measured_peer_stores": NVLINK_MEASURED,
                  "frac_of_nominal": comm * value / 1e9 / NVLINK_NOMINAL, "frac_of_measured": comm * value / 1e9 / NVLINK_MEASURED,
                  "what": "bytes this rank stores into peers' receive slots per step / step time: the share of the step during which the link would be busy at its peak"}

    # ---- context: the reference's own GPU build; the reference's CPU path on the same grid; the secondary configurations
    context, cpu, extras = None, None, None
    if rank == 0 and world == 1 and CONFIG == "ch3d" and not args.no_context:
        try:
            context = reference_cufft_context(n, 10)
            context["speedup_over_ref_cufft"] = value / context["ref_cufft_steps_per_s"] if context.get("ref_cufft_steps_per_s") else None
        except Exception as e:   # context must never cost the headline line
            context = {"ref_cufft_steps_per_s": None, "why": f"{type(e).__name__}: {e}"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            nref, why = reference_grid(n)
            cores = os.cpu_count() or 1
            per, _ = cpu_reference_steps(nref, 1, args.cpu_steps, cores)
            sec = float(np.mean(per)) * (n ** 3 / nref ** 3)
            cpu = {"value": 1.0 / sec, "unit": "steps/s", "cores": cores, "kind": "reference",
                   "sample": f"unmodified reference CPU path on the SAME CH-3D {nref}^3 grid: {args.cpu_steps} timed steps after 1 warm-up step, FFTW-API shim threaded over "
                             f"{cores} host threads (the reference's pointwise passes stay serial); the one-thread figure is in the `--impl reference` line"
                             + (f"; {why}, scaled by point count" if why else ""),
                   "host_cores_available": cores}
        except Exception as e:
            cpu = {"value": None, "unit": "steps/s", "cores": 0, "kind": "reference", "sample": f"failed: {type(e).__name__}: {e}"}
    if CONFIG == "ch3d" and not args.no_extra:
        extras = {}
        todo = [("cfg02_ch2d_4096", "ch2d", 4096), ("cfg04_modelh_2048", "modelh", 2048), ("cfg06_kpz3d_512", "kpz3d", 512)] if world == 1 else \
               ([("cfg06_kpz3d_1024_8gpu", "kpz3d", 1024)] if world == 8 else [])
        for key, nm, size in todo:
            try:
                res = run_extra(nm, size, args.extra_steps, peak, Evolver, args.extra_parity_steps, rank, world, dist, unique_id if world > 1 else None)
                if rank == 0:
                    extras[key] = res
            except Exception as e:
                extras[key] = {"error": f"{type(e).__name__}: {e}"}

    rc = 0
    if rank == 0:
        metric = "timesteps/s, Cahn-Hilliard 3D 512^3" if CONFIG == "ch3d" else f"timesteps/s, KPZ 3D {n}^3 with noise (secondary configuration)"
        line = {"metric": metric, "value": value, "unit": "steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": (f"examples/03_cahn_hilliard_3d {n}^3, dt={DT}, a=-1 b=1 k=4, deterministic, IC 0.01*(2u-1)" if CONFIG == "ch3d" else
                                        f"examples/06_kpz as a 3-D system {n}^3, dt={DT}, D=0.5 l=0.5, noise 2*D on h, IC h=0"),
                           "partition": f"z-slabs over {world} GPU(s)", "l2": "working set 4 x 0.55 GB per step >> 126 MB L2 (no flush needed)",
                           "grid_point_steps_per_s": value * n ** 3},
                "parity": parity, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "a2a_bytes_per_step_per_gpu": comm, "nvlink": nvlink, "context": context, "extra_configs": extras}
        print(json.dumps(line), flush=True)
        if parity is not None and parity.get("ok") is False:
            print(f"bench.py: PARITY FAILED against the reference (tolerance {PARITY_TOL}): {json.dumps(parity)}", file=sys.stderr, flush=True)
            rc = 1
    if dist is not None:
        flag = torch.tensor([rc], device="cuda")
        dist.broadcast(flag, 0)
        rc = int(flag.item())
        dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()
