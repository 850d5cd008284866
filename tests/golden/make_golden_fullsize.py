"""Full-size witness for BASELINE.json configs[2]: the reference's own CPU path (oracle/_ref/libcupss_ref_f.so = its
sources with the two one-token fixes that give its GPU kernels' semantics, FFTW-API shim threaded over the host cores)
run HERE on Cahn-Hilliard 3-D at 512^3 for STEPS steps from the seeded smooth initial condition of tests/cases.py.

A 512^3 snapshot does not belong in git, so the fixture keeps what a test needs to tell a wrong run from a right one:
  sub      phi[::16, ::16, ::16]                      (32^3 point samples)
  planes   phi[z0], phi[:, y0], phi[:, :, x0] at 16x   (three full planes, decimated by 4 in-plane)
  stats    mean, L2 norm, min, max, sum |phi|^3 of the whole field
tests/test_gpu_parity.py::test_cahn_hilliard_3d_full_size_matches_the_reference compares the product with it.
Usage (build container, ~10 GB RAM, a few minutes): python tests/golden/make_golden_fullsize.py [steps [N]]
N = 512 (default, 12 steps) writes ch3d_512_ref.npz; N = 256 with 100 steps (the north star's "relative L2 <= 1e-5 per
field after 100 steps" at a non-toy 3-D size, SURVEY.md 8d "03 at <= 256^3 fully") writes ch3d_256_ref.npz.
"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N = 512
STEPS = 12


def summarise(phi):
    """Samples / planes / statistics of a cubic field; the strides and plane indices scale with the edge (512: as committed)."""
    n = phi.shape[0]
    s, d = max(1, n // 32), max(1, n // 128)
    z0, y0, x0 = (37 * n) // 512, (201 * n) // 512, (333 * n) // 512
    p64 = phi.astype(np.float64)
    return dict(sub=phi[::s, ::s, ::s].copy(),
                plane_z=phi[z0, ::d, ::d].copy(), plane_y=phi[::d, y0, ::d].copy(), plane_x=phi[::d, ::d, x0].copy(),
                stats=np.array([p64.mean(), np.sqrt((p64 ** 2).sum()), p64.min(), p64.max(), (np.abs(p64) ** 3).sum()]))


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else STEPS
    n = int(sys.argv[2]) if len(sys.argv) > 2 else N
    import cases
    from cases import CASES, ORACLE_F
    case = dict(CASES["ch3d_32"])
    case["shape"] = (n, n, n)
    case["steps"] = steps
    shim = C.CDLL(ORACLE_F)
    if hasattr(shim, "cupss_shim_set_threads"):
        shim.cupss_shim_set_threads(os.cpu_count())
    t0 = time.perf_counter()
    phi = cases.run_case(case, lib=ORACLE_F, device=0)["phi"]
    print(f"reference CPU path, {n}^3, {steps} steps: {time.perf_counter() - t0:.1f} s")
    s = summarise(phi)
    print({k: (v.shape if k != "stats" else v) for k, v in s.items()})
    np.savez_compressed(os.path.join(HERE, f"ch3d_{n}_ref.npz"), steps=np.array([steps]), **s)


if __name__ == "__main__":
    main()
