#!/usr/bin/env python
"""bench.py -- timesteps/s of Cahn-Hilliard 3-D 512^3 (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            own arm (CUDA engine through the evolver API)
  python bench.py --impl reference ...                     the reference's own CPU implementation (oracle/_ref)

One JSON line on rank 0.  `value` = steps/s with the state resident in HBM (CUDA events on the engine's stream,
max over ranks); `e2e` = the same job through the public API from HOST buffers (upload of the initial condition,
K steps, download of the result, all inside the timed region); `roofline` = the dominant kernel's algorithmic
bytes / its event-timed duration against MEASURED_PEAKS.json; `cpu_baseline` = the reference CPU path on a bounded
sample.  Strong scaling: the 512^3 grid is fixed and slab-partitioned over the ranks.
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


CONFIG = "ch3d"   # --config: "ch3d" (the headline, BASELINE.json configs[2]) or "kpz3d" (configs[4]: 3-D KPZ with noise)


def make_system(Evolver, dev, n, lib=None):
    ev = Evolver(dev, n, n, n, 1.0, 1.0, 1.0, DT, lib=lib)
    if CONFIG == "kpz3d":   # examples/06_kpz lifted to 3-D: h + three gradient constraint fields, white noise on h
        for f, d in (("h", True), ("iqxh", False), ("iqyh", False), ("iqzh", False)):
            ev.createField(f, d)
        ev.addParameter("D", 0.5)
        ev.addParameter("l", 0.5)
        for e in ("dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2 + l*iqzh^2", "iqxh = iqx*h", "iqyh = iqy*h", "iqzh = iqz*h"):
            ev.addEquation(e)
        ev.addNoise("h", "2*D")
        if lib is None:
            ev.setNoiseSeed(1234)
        return ev
    ev.createField("phi", True)
    for k, v in PARAMS:
        ev.addParameter(k, v)
    ev.addEquation(EQ)
    return ev


def main_field():
    return "h" if CONFIG == "kpz3d" else "phi"


def synthetic_ic(n, seed=1324):
    rng = np.random.default_rng(seed)
    return (0.01 * (2.0 * rng.random((n, n, n), dtype=np.float32) - 1.0)).astype(np.float32)


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


def cpu_reference_rate(n_sample, steps, threads):
    """Reference CPU path (oracle/_ref/libcupss_ref_u.so = unmodified sources + FFTW-API shim) on an n_sample^3 grid.
    Returns (seconds per step, points)."""
    import ctypes as C
    from cupss_b200.capi import Evolver, RUN_CPU
    lib = os.path.join(ROOT, "oracle", "_ref", "libcupss_ref_u.so")
    shim = C.CDLL(lib)
    if hasattr(shim, "cupss_shim_set_threads"):
        shim.cupss_shim_set_threads(int(threads))
    ev = make_system(Evolver, RUN_CPU, n_sample, lib=lib)
    ev.setReal("phi", synthetic_ic(n_sample))
    ev.prepareProblem()
    ev.advanceTime(1)
    t0 = time.perf_counter()
    ev.advanceTime(steps)
    dt = (time.perf_counter() - t0) / steps
    ev.close()
    return dt, n_sample ** 3


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.size
    cores = os.cpu_count() or 1
    ns = args.cpu_sample
    per = []
    for _ in range(max(1, args.warmup)):
        cpu_reference_rate(ns, 1, cores)
    for _ in range(args.steps):
        dt, pts = cpu_reference_rate(ns, args.cpu_steps, cores)
        per.append(dt)
    sec_per_step_sample = float(np.mean(per))
    # bounded sample: an ns^3 grid; the metric is quoted for n^3, so scale by the point count (FFT log factor ignored, in the CPU's favour)
    sec_per_step = sec_per_step_sample * (n ** 3) / (ns ** 3)
    value = 1.0 / sec_per_step
    sample = f"CH-3D {ns}^3 grid, {args.cpu_steps} steps per timed step, scaled by point count to {n}^3"
    line = {"impl": "reference", "metric": "timesteps/s, Cahn-Hilliard 3D 512^3", "value": value, "unit": "steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"examples/03_cahn_hilliard_3d {n}^3, dt={DT}, a=-1 b=1 k=4, deterministic, IC 0.01*(2u-1)",
                       "partition": "host cores of the box (reference CPU path, FFT shim threaded)", "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--config", default="ch3d", choices=["ch3d", "kpz3d"], help="ch3d = the headline metric; kpz3d = BASELINE.json configs[4] (secondary)")
    ap.add_argument("--cpu-sample", type=int, default=128, help="grid edge of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--clock-period-ms", type=float, default=25.0, help="NVML clock / throttle-reason polling period inside the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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

    n = args.size
    ev = make_system(Evolver, RUN_GPU, n)
    if world > 1:
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            capi.engine_check(eng.cupss_b200_nccl_unique_id(idbuf), "nccl_unique_id")
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        ev.setPartition(rank, world, bytes(t.cpu().numpy().tobytes()))
    ic = synthetic_ic(n) if CONFIG == "ch3d" else None   # kpz3d starts from h = 0: the (zero) host mirrors are left untouched
    if ic is not None:
        ev.setReal(main_field(), ic)
    ev.prepareProblem()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

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
    prof = {}
    reps = 5
    for _ in range(reps):
        for name, t_ms, by in ev.profileStep():
            a = prof.setdefault(name, [0.0, 0.0, 0])
            a[0] += t_ms; a[1] = by; a[2] += 1
    barrier()
    kernels = {k: {"ms": v[0] / reps, "bytes": v[1] * v[2] / reps, "launches_per_step": v[2] / reps} for k, v in prof.items() if k != "bump"}
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
                "per_kernel": {k: {"ms": round(v["ms"], 4), "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in kernels.items()},
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
        ev.copyAllDataToHost() if world == 1 else ev._lib.cupss_capi_copy_all_data_to_host(ev._h)
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
        ev.copyAllDataToHost() if world == 1 else ev._lib.cupss_capi_copy_all_data_to_host(ev._h)
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
               "d2h_bytes_per_step": (2 if world == 1 else 1) * slab_bytes * world / args.steps, "seconds": el, "phases": phases,
               "what": "host mirror filled -> prepareProblem (H2D) + K x advanceTime + copyAllDataToHost (D2H of real and Fourier arrays), wall clock, max over ranks"}

    launches = ev.launchesPerStep() * args.steps
    comm = ev.commBytesPerStep()
    ev.close()

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        ns = args.cpu_sample
        sec, pts = cpu_reference_rate(ns, args.cpu_steps, 1)
        sec_full = sec * n ** 3 / pts
        cpu = {"value": 1.0 / sec_full, "unit": "steps/s", "cores": 1, "kind": "reference",
               "sample": f"unmodified reference CPU path (serial, FFTW-API shim) on a {ns}^3 CH-3D grid, {args.cpu_steps} steps, scaled by point count to {n}^3",
               "host_cores_available": os.cpu_count()}

    if rank == 0:
        metric = "timesteps/s, Cahn-Hilliard 3D 512^3" if CONFIG == "ch3d" else f"timesteps/s, KPZ 3D {n}^3 with noise (secondary configuration)"
        line = {"metric": metric, "value": value, "unit": "steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": (f"examples/03_cahn_hilliard_3d {n}^3, dt={DT}, a=-1 b=1 k=4, deterministic, IC 0.01*(2u-1)" if CONFIG == "ch3d" else
                                        f"examples/06_kpz as a 3-D system {n}^3, dt={DT}, D=0.5 l=0.5, noise 2*D on h, IC h=0"),
                           "partition": f"z-slabs over {world} GPU(s)", "l2": "working set 4 x 0.55 GB per step >> 126 MB L2 (no flush needed)",
                           "grid_point_steps_per_s": value * n ** 3},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "a2a_bytes_per_step_per_gpu": comm}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
