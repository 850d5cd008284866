"""CPU suite, part 1: pin the oracle.  The compiled reference (oracle/_ref, unmodified sources + FFTW-API shim) must
reproduce the reference's own golden vectors; the numpy restatement and the committed ref_runs fixtures must agree
with it; the shim FFT must agree with numpy."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import refcases
from cases import CASES, ORACLE_F, ORACLE_U, SHIM, rel_l2

pytestmark = pytest.mark.skipif(not os.path.exists(ORACLE_U) and not os.path.isdir("/root/reference/src"),
                                reason="oracle/_ref not built and /root/reference absent")


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_reference_init_cases_match_base_truths(built, dim):
    assert refcases.init_case(dim, 0, ORACLE_U) < refcases.TOL


@pytest.mark.parametrize("dim", [1, 3])
@pytest.mark.parametrize("lib", [ORACLE_U, ORACLE_F], ids=["unmodified", "fixed"])
def test_reference_operator_cases_match_base_truths(built, dim, lib):
    errs = refcases.operators_case(dim, 0, lib, "cpu")
    assert max(errs.values()) < refcases.TOL, errs


def test_shim_fft_matches_numpy(built):
    lib = C.CDLL(SHIM)
    lib.fftwf_plan_dft_3d.restype = C.c_void_p
    lib.fftwf_plan_dft_3d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
    lib.fftwf_execute.argtypes = [C.c_void_p]
    lib.fftwf_destroy_plan.argtypes = [C.c_void_p]
    rng = np.random.default_rng(0)
    for shape in [(8, 16, 32), (12, 10, 9), (1, 64, 64)]:
        x = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
        for sign, ref in ((-1, np.fft.fftn), (1, lambda a: np.fft.ifftn(a) * a.size)):
            for dbl in (1, 0):
                lib.cupss_shim_set_double(dbl)
                y = np.empty_like(x)
                p = lib.fftwf_plan_dft_3d(shape[0], shape[1], shape[2], x.ctypes.data, y.ctypes.data, sign, 64)
                lib.fftwf_execute(p)
                lib.fftwf_destroy_plan(p)
                assert rel_l2(y, ref(x.astype(np.complex128))) < (2e-7 if dbl else 2e-6)
    lib.cupss_shim_set_double(1)


@pytest.mark.parametrize("name", ["ch2d_64", "modelh_32", "kpz3d_32_det", "ops3d_16", "mixed2d_64x32", "mixed3d_32x16x8"])
def test_numpy_restatement_matches_compiled_reference(built, name):
    """Independent float64 restatement (oracle/restatement.py) vs the compiled reference with the two one-token fixes
    (== the reference GPU kernels' semantics).  20 steps; the float32 reference sits ~1e-6 from float64."""
    from cupss_b200.capi import Evolver  # noqa: F401
    from oracle.restatement import from_plan_dump
    case = dict(CASES[name])
    ev = cases.build_system(case, lib=ORACLE_F, device=0)
    ics = {n: ev.real(n) for n, _ in case["fields"]}
    ev.prepareProblem()
    sys_np = from_plan_dump(ev.dumpPlan(), case["shape"], (1.0, 1.0, 1.0), case["dt"], np.float64, "gpu")
    for n, a in ics.items():
        sys_np.real[n] = a.astype(np.float64)
    sys_np.prepare()
    steps = min(20, case["steps"])
    ev.advanceTime(steps)
    sys_np.step(steps)
    for n, _ in case["fields"]:
        ref = ev.real(n)
        if np.linalg.norm(ref) == 0:
            assert np.linalg.norm(sys_np.real[n]) < 1e-12
            continue
        assert rel_l2(sys_np.real[n], ref) < 2e-5, (n, rel_l2(sys_np.real[n], ref))
    ev.close()


def test_cpu_dealias_rule_differs_from_gpu_rule_in_2d(built):
    """SURVEY.md 8c hazard 2: the unmodified CPU loop wipes ky != 0 modes of the dealiased copy in 2-D."""
    a = cases.run_case(CASES["ch2d_64"], lib=ORACLE_F, device=0, steps=30)["phi"]
    b = cases.run_case(CASES["ch2d_64"], lib=ORACLE_U, device=0, steps=30)["phi"]
    assert rel_l2(a, b) > 1e-4


def test_committed_reference_runs_are_reproduced(built):
    """tests/golden/ref_runs.npz was produced in the build container from the compiled reference; the oracle build
    that travels to the GPU box must reproduce it (guards against a stale or mis-built oracle)."""
    path = os.path.join(cases.GOLDEN, "ref_runs.npz")
    gold = np.load(path)
    for name in ["ch3d_64x32x16", "ops1d_16"]:
        case = CASES[name]
        out = cases.run_case(case, lib=ORACLE_U if case.get("oracle") == "U" else ORACLE_F, device=0)
        for f, arr in out.items():
            assert rel_l2(arr, gold[f"{name}/{f}"]) < 1e-6 or np.linalg.norm(gold[f"{name}/{f}"]) == 0


def test_fourier_callback_semantics_of_the_reference(built):
    """field::setRHS (src/field.cpp:48-66, 88-89) with a Fourier-space callback: the callback edits comp_array after the
    update; toReal -> normalize keeps the REAL PART of the inverse transform; toComp transforms that back.  So what
    survives of a callback that breaks the Hermitian symmetry is the Hermitian part of its output.  Restated in numpy
    (float64) for one diffusion step with the facade's built-in callback of kind 1 and pinned on the compiled reference:
    this is the rule the engine's comp_view_commit implements (cupss_b200/csrc/engine.cu)."""
    sx, sy, dt = 32, 16, 0.05
    case = dict(shape=(sx, sy, 1), dt=dt, fields=[("phi", 1)], params={}, eqs=["dt phi + q^2*phi = 0"],
                ic=dict(phi=("smooth", (0.5, 0.2))), steps=1, fourier_callbacks=[("phi", 1)], device=0, oracle="U")
    got = cases.run_case(case, lib=ORACLE_U, device=0)["phi"][0]
    phi0 = cases.smooth_ic(sx, sy, 1, 0.5, 0.2)[0].astype(np.float64)
    sm = lambda n: np.where(np.arange(n) <= n // 2, np.arange(n), np.arange(n) - n).astype(np.float64)   # the facade's signed_mode
    ny, nx = np.meshgrid(sm(sy), sm(sx), indexing="ij")
    q2 = (2 * np.pi * nx / sx) ** 2 + (2 * np.pi * ny / sy) ** 2
    F = np.fft.fft2(phi0) / (1.0 + dt * q2)
    F = F / (1.0 + 0.0005 * (nx ** 2 + ny ** 2))                       # kind 0 part: Hermitian low-pass
    band = (nx >= 1) & (nx <= 3) & (ny >= 0)
    F = np.where(band, F * (0.98 + 0.05j), F)                             # kind 1 part: one-sided, breaks the symmetry
    want = np.fft.ifft2(F).real
    assert rel_l2(got, want) < 5e-6, rel_l2(got, want)
    # and it is NOT what a symmetry-preserving reading (apply the factor to both k and -k) would give
    sym = np.fft.ifft2(np.where(band | ((nx <= -1) & (nx >= -3) & (ny <= 0)), F * 0 + np.fft.fft2(phi0) / (1.0 + dt * q2) / (1.0 + 0.0005 * (nx ** 2 + ny ** 2)) * 0.98, F)).real
    assert rel_l2(got, sym) > 1e-3


def test_smoke_checker_arm_is_a_real_run(built):
    """The oracle arm of __graft_entry__.smoke() must hold the reference's RESULT: calling copyAllDataToHost on a RUN_CPU
    reference evolver silently puts the initial condition back (its device arrays are never updated on that path)."""
    import __graft_entry__ as ge
    from cupss_b200.capi import RUN_CPU
    n, ic = ge._smoke_case()
    want = ge._smoke_run(ORACLE_F, RUN_CPU)
    assert np.isfinite(want).all()
    assert rel_l2(want, ic) > 0.05          # five steps moved the field
    case = dict(CASES["ch3d_32"])
    case["steps"] = 5
    case["ic"] = {}
    ev = cases.build_system(case, lib=ORACLE_F, device=0)
    ev.setReal("phi", ic)
    ev.prepareProblem()
    ev.advanceTime(5)
    assert np.array_equal(ev.real("phi"), want)
    ev.close()
