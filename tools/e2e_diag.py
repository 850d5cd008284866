"""GPU-box diagnostic: PCIe copy bandwidth (torch, pinned) next to the engine's upload / download calls at 512^3."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cupss_b200 import capi
from cupss_b200.capi import Evolver, RUN_GPU

n = 512
h = torch.empty(n * n * n * 2, dtype=torch.float32).pin_memory()
d = torch.empty_like(h, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"torch pinned {name}: 1 GiB in {dt*1e3:.1f} ms = {h.numel()*4/dt/1e9:.1f} GB/s")
del h, d

eng = capi.load_engine()
ev = bench.make_system(Evolver, RUN_GPU, n)
ev.setReal("phi", bench.synthetic_ic(n))
t0 = time.perf_counter(); ev.prepareProblem(); print(f"prepareProblem (first, incl. allocations): {(time.perf_counter()-t0)*1e3:.1f} ms")
ev.advanceTime(5); ev.sync()
plan = C.c_void_p(ev._lib.cupss_capi_engine_plan(ev._h))
real = ev.fieldReal("phi"); comp = ev.fieldFourier("phi")
for rep in range(2):
    t0 = time.perf_counter()
    capi.engine_check(eng.cupss_b200_download_real(plan, 0, real.ctypes.data_as(C.c_void_p)), "download_real")
    t1 = time.perf_counter()
    capi.engine_check(eng.cupss_b200_download_comp(plan, 0, comp.ctypes.data_as(C.c_void_p)), "download_comp")
    t2 = time.perf_counter()
    capi.engine_check(eng.cupss_b200_upload_real(plan, 0, real.ctypes.data_as(C.c_void_p)), "upload_real")
    t3 = time.perf_counter()
    print(f"rep {rep}: download_real {(t1-t0)*1e3:.1f} ms, download_comp {(t2-t1)*1e3:.1f} ms, upload_real {(t3-t2)*1e3:.1f} ms")
t0 = time.perf_counter(); ev.copyAllDataToHost(); print(f"copyAllDataToHost: {(time.perf_counter()-t0)*1e3:.1f} ms")
ev.close()
