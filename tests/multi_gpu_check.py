"""Run under torchrun (one rank per GPU): slab-partitioned CH-3D and KPZ-3D vs the single-GPU run of the same system."""
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
    for name, steps in (("ch3d_64x32x16", 20), ("kpz3d_32_det", 10), ("ops3d_16", 3)):
        case = cases.CASES[name]
        sx, sy, sz = case["shape"]
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            capi.engine_check(eng.cupss_b200_nccl_unique_id(idbuf), "uid")
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        ev = cases.build_system(case)
        ev.setPartition(rank, world, bytes(t.cpu().numpy().tobytes()))
        ev.prepareProblem()
        ev.advanceTime(steps)
        ev._lib.cupss_capi_copy_all_data_to_host(ev._h)
        zl = sz // world
        mine = {n: ev.real(n)[rank * zl:(rank + 1) * zl] for n, _ in case["fields"]}
        ev.close()
        single = cases.run_case(case, steps=steps)
        for n, _ in case["fields"]:
            ref = single[n][rank * zl:(rank + 1) * zl]
            same = np.array_equal(mine[n], ref)
            err = cases.rel_l2(mine[n], ref)
            print(f"rank {rank} {name} {n}: bitwise={same} rel={err:.2e}", flush=True)
            ok = ok and err < 1e-6
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag.item()) == 1:
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
