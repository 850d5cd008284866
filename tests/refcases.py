"""The reference's own unit tests (/root/reference/tests/tests.cpp) restated once, runnable against any library
that exports the facade: the compiled reference (oracle) or the product.  Same grids (16 points per axis), same
calls, same 1e-4 absolute tolerance, same golden files -- plus the iqy/iqz columns the reference forgets to assert
(tests.cpp:325-337 re-checks iqx)."""
import numpy as np

from cases import load_truth
from cupss_b200.capi import Evolver

TOL = 1e-4   # EXPECT_NEAR(..., 1e-4), tests/tests.cpp:77


def _ctor(dim, device, lib):
    if dim == 1:
        return Evolver(device, 16, 1, 1, 1.0, 1.0, 1.0, 1.0, write_every=1, lib=lib)     # evolver(dev, 16, 1.0, 1.0, 1)
    if dim == 2:
        return Evolver(device, 16, 16, 1, 1.0, 1.0, 1.0, 1.0, write_every=1, lib=lib)
    return Evolver(device, 16, 16, 16, 1.0, 1.0, 1.0, 1.0, write_every=1, lib=lib)


def init_case(dim, device, lib):
    """OneD/TwoD/ThreeD{CPU,GPU}Init (tests.cpp:64-189): droplet -> prepareProblem [-> copyAllDataToHost]."""
    ev = _ctor(dim, device, lib)
    ev.createField("phi", True)
    ev.initializeDroplet("phi", 0, 1, 16 / 8, 4, 16 // 2, 16 // 2 if dim > 1 else 0, 0)
    ev.prepareProblem()
    if device:
        ev.copyAllDataToHost()
    got = ev.real("phi").ravel()
    ev.close()
    want = load_truth(f"phi_{dim}d")
    return float(np.max(np.abs(got - want)))


def operators_case(dim, device, lib, flavour):
    """OneD/ThreeD{CPU,GPU}Operators (tests.cpp:191-422): one advanceTime, constraint fields vs golden columns."""
    ev = _ctor(dim, device, lib)
    names = ["phi", "lapphi", "iqxphi", "invqphi"] + (["iqyphi", "iqzphi"] if dim == 3 else [])
    for n in names:
        ev.createField(n, n == "phi")
    ev.addEquation("dt phi+q^2*phi = iqxphi^2")
    ev.addEquation("lapphi = -q^2*phi")
    ev.addEquation("iqxphi =  iqx*phi")
    if dim == 3:
        ev.addEquation("iqyphi =  iqy*phi")
        ev.addEquation("iqzphi =  iqz*phi")
    ev.addEquation("invqphi =  1/q*phi")
    ev.initializeDroplet("phi", 0, 1, 16 / 8, 4, 16 // 2, 16 // 2 if dim > 1 else 0, 0)
    ev.prepareProblem()
    ev.advanceTime(1)
    if device:
        ev.copyAllDataToHost()
    errs = {}
    for n in names[1:]:
        want = load_truth(f"{n}_{flavour}_{dim}d")
        errs[n] = float(np.max(np.abs(ev.real(n).ravel() - want)))
    ev.close()
    return errs
