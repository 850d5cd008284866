"""Run under torchrun (one rank per GPU): the slab-partitioned engine against the single-GPU run of the same system.

  * CH-3D, KPZ-3D (deterministic) and the 3-D operator case: every rank's z-slab of every field bit-identical to the 1-GPU run;
  * fieldsFourier (comp_array) of a partitioned run: each rank's kz planes bit-identical to the 1-GPU spectrum
    (evolver::copyAllDataToHost copies both arrays, src/evolver.cpp:364-368);
  * real-space user callbacks on a partitioned run (mirror boundary condition, slab-local);
  * the default noise seed is the same on every rank, and a noisy run with that seed equals the 1-GPU run with it.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C

import cases
from cupss_b200 import capi


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = capi.load_engine()
    ok = True

    def unique_id():
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            capi.engine_check(eng.cupss_b200_nccl_unique_id(idbuf), "uid")
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    def check(tag, mine, ref, bitwise=True):
        nonlocal ok
        same = np.array_equal(mine, ref)
        err = cases.rel_l2(mine, ref)
        print(f"rank {rank} {tag}: bitwise={same} rel={err:.2e}", flush=True)
        ok = ok and (same if bitwise else err < 1e-6)

    bc3d = dict(shape=(32, 32, 16), dt=0.002, fields=[("v", 1), ("iqxv", 0), ("x", 0)], params={},
                eqs=["dt v +0.5*q^2*v = iqx*x*iqxv + x*iqy^2*v", "iqxv = iqx*v"],
                ic=dict(v=("smooth", (0.6, 0.05)), x=("smooth", (0.3, 0.0))), steps=12, callbacks=[("v", False), ("x", False)], device=0)
    # long axes shared by thread-block clusters, with the fused push exchange: 1024-row y passes (pushed forward pass, pruned
    # inverse from the exchange layout) and a 1024-row z axis (k stage whose cross-level gather pushes the rows to their owners)
    ch_long_y = dict(cases.CASES["ch3d_64x32x16"]); ch_long_y["shape"] = (32, 1024, 16)
    ch_long_z = dict(cases.CASES["ch3d_64x32x16"]); ch_long_z["shape"] = (32, 16, 1024)
    kpz_long_z = dict(cases.CASES["kpz3d_32_det"]); kpz_long_z["shape"] = (32, 16, 1024)
    for name, case, steps in (("ch3d_64x32x16", cases.CASES["ch3d_64x32x16"], 20), ("kpz3d_32_det", cases.CASES["kpz3d_32_det"], 10),
                              ("ops3d_16", cases.CASES["ops3d_16"], 3), ("bc3d_mirror_callbacks", bc3d, 12),
                              ("ch3d_32x1024x16", ch_long_y, 8), ("ch3d_32x16x1024", ch_long_z, 8), ("kpz3d_32x16x1024", kpz_long_z, 6)):
        sx, sy, sz = case["shape"]
        dev = case.get("device", 1)
        ev = cases.build_system(case, device=dev)
        ev.setPartition(rank, world, unique_id())
        ev.prepareProblem()
        ev.advanceTime(steps)
        ev.copyAllDataToHost()
        zl = sz // world
        mine = {n: ev.real(n)[rank * zl:(rank + 1) * zl] for n, _ in case["fields"]}
        mine_c = {n: ev.comp(n)[rank * zl:(rank + 1) * zl] for n, _ in case["fields"]}
        ev.close()
        ev = cases.build_system(case, device=dev)
        ev.prepareProblem()
        ev.advanceTime(steps)
        ev.copyAllDataToHost()
        for n, _ in case["fields"]:
            check(f"{name} {n} real", mine[n], ev.real(n)[rank * zl:(rank + 1) * zl])
            check(f"{name} {n} fieldsFourier", mine_c[n], ev.comp(n)[rank * zl:(rank + 1) * zl])
        ev.close()

    # default seed: shared by the ranks (hash of the NCCL id), and the noisy partitioned run equals the 1-GPU run with that seed
    from cupss_b200.capi import Evolver

    def noisy(partition, seed):
        ev = Evolver(1, 32, 32, 32, 1.0, 1.0, 1.0, 0.01)
        ev.createField("h", True)
        ev.addEquation("dt h + 0.5*q^2*h = 0")
        ev.addNoise("h", "1.0")
        if partition:
            ev.setPartition(rank, world, unique_id())
        if seed is not None:
            ev.setNoiseSeed(seed)
        ev.prepareProblem()
        ev.advanceTime(4)
        ev.copyAllDataToHost()
        out, c, s = ev.real("h"), ev.comp("h"), ev.getNoiseSeed()
        ev.close()
        return out, c, s
    zl = 32 // world
    part, part_c, seed = noisy(True, None)
    seeds = [None] * world
    dist.all_gather_object(seeds, seed)
    print(f"rank {rank} default seeds {seeds}", flush=True)
    ok = ok and len(set(seeds)) == 1 and seeds[0] != 0
    single, single_c, _ = noisy(False, seeds[0])
    check("noise default-seed real", part[rank * zl:(rank + 1) * zl], single[rank * zl:(rank + 1) * zl])
    check("noise default-seed fieldsFourier", part_c[rank * zl:(rank + 1) * zl], single_c[rank * zl:(rank + 1) * zl])

    # KPZ with noise (BASELINE.json configs[4] in small): the lean noisy evaluator next to the fused exchange, incl. a z axis
    # that a cluster shares -- bit-identical to one GPU
    def kpz_noisy(partition, shape):
        ev = Evolver(1, *shape, 1.0, 1.0, 1.0, 0.01)
        for f, d in (("h", 1), ("iqxh", 0), ("iqyh", 0), ("iqzh", 0)):
            ev.createField(f, d)
        ev.addParameter("l", 0.5)
        ev.addParameter("D", 0.5)
        for e in ("dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2 + l*iqzh^2", "iqxh = iqx*h", "iqyh = iqy*h", "iqzh = iqz*h"):
            ev.addEquation(e)
        ev.addNoise("h", "2*D")
        ev.setReal("h", cases.smooth_ic(*shape, 1.0, 0.1))
        if partition:
            ev.setPartition(rank, world, unique_id())
        ev.setNoiseSeed(777)
        ev.prepareProblem()
        ev.advanceTime(6)
        ev.copyAllDataToHost()
        out = ev.real("h")
        ev.close()
        return out
    for shape in ((64, 32, 32), (32, 16, 1024)):
        zs = shape[2] // world
        got, want = kpz_noisy(True, shape), kpz_noisy(False, shape)
        check(f"kpz noisy {shape}", got[rank * zs:(rank + 1) * zs], want[rank * zs:(rank + 1) * zs])

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag.item()) == 1:
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
