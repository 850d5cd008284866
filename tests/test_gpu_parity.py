"""GPU suite (-m gpu): the CUDA path, called through the C ABI behind the evolver API, against
  (1) the reference's golden vectors (its own unit tests, restated in refcases.py),
  (2) the compiled reference CPU path (oracle/_ref) on the same seeded inputs -- tolerance of BASELINE.json:
      relative L2 <= 1e-5 per field after 100 steps,
  (3) the committed outputs of the reference produced in the build container (tests/golden/ref_runs.npz),
  (4) size-independent properties at the benchmark's full sizes (mass conservation, closed-form diffusion decay,
      noise variance and spectrum).
Nothing here reads /root/reference."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import refcases
from cases import CASES, ORACLE_F, ORACLE_U, ROOT, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5   # BASELINE.json north_star: relative L2 <= 1e-5 per field after 100 steps


def _engine_loaded():
    with open("/proc/self/maps") as f:
        return "libcupss_b200.so" in f.read()


def test_native_engine_is_the_path_that_runs(built):
    import torch
    assert torch.cuda.is_available(), "GPU suite needs a CUDA device"
    out = cases.run_case(CASES["ops1d_16"], steps=1)
    assert np.isfinite(out["phi"]).all()
    assert _engine_loaded(), "libcupss_b200.so not mapped: the CUDA engine did not run"


@pytest.mark.parametrize("shape", [(16, 1, 1), (64, 1, 1), (2, 1, 1), (16, 16, 1), (64, 32, 1), (16, 16, 16), (32, 16, 64), (128, 128, 1), (64, 64, 64), (1024, 4, 1), (4, 2048, 1)])
def test_upload_download_round_trip_and_spectrum(built, shape):
    from cupss_b200.capi import Evolver
    sx, sy, sz = shape
    ev = Evolver(1, sx, sy, sz, 1.0, 1.0, 1.0, 0.1)
    ev.createField("phi", True)
    ev.addEquation("dt phi + q^2*phi = 0")
    ic = cases.smooth_ic(sx, sy, sz, 0.5, 0.5)
    ev.setReal("phi", ic)
    ev.prepareProblem()
    ev.copyAllDataToHost()
    assert rel_l2(ev.real("phi"), ic) < 1e-6
    assert rel_l2(ev.comp("phi"), np.fft.fftn(ic.astype(np.float64))) < 1e-6
    ev.close()


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_reference_unit_tests_init(built, dim):
    assert refcases.init_case(dim, 1, None) < refcases.TOL


@pytest.mark.parametrize("dim", [1, 3])
@pytest.mark.parametrize("flavour", ["gpu", "cpu"])
def test_reference_unit_tests_operators(built, dim, flavour):
    errs = refcases.operators_case(dim, 1, None, flavour)
    assert max(errs.values()) < refcases.TOL, errs


@pytest.mark.parametrize("name", ["diffusion2d_256", "ch2d_64", "ch2d_64_cpu_rule", "ch3d_32", "ch3d_64x32x16", "ch3d_128x16x16", "ch2d_512x16", "ch3d_512x8x8", "burgers_like_128", "sh2d_512x16_two_monomials", "sh3d_256x16x8_two_monomials", "quartic2d_1024x16", "quartic1d_128",
                                  "ch3d_256x16x8", "ch2d_1024x32", "ch2d_2048x32", "ch2d_4096x32", "ch3d_1024x32x8", "burgers1d_2048",
                                  "modelh_32", "kpz3d_32_det", "kpz3d_128x16x16_det", "kpz2d_512x16_mixed_powers", "kpz3d_1024x16x8_det", "kpz2d_256x32_mixed_powers", "ops1d_16", "ops3d_16", "bc_even_inhomogeneous_64", "bc_odd_diffusion_64",
                                  "fcb_lowpass_ch2d_64", "fcb_asym_ch3d_16", "fcb_constraint_kpz2d_32", "fcb_band_diffusion_1d_64", "fcb_band_allen_cahn_2d_64",
                                  "bc_clamp_product_64", "bc_clamp_product_1d_128", "ch2d_1024", "ch2d_64x4096", "modelh_256",
                                  "ch3d_32x1024x8", "ch3d_32x8x2048", "kpz2d_128x2048_det", "ch3d_32x64x64",
                                  "mixed2d_2048x16", "mixed2d_1024x16", "mixed2d_4096x8", "mixed3d_1024x8x8", "modelh_2048x64", "mixed1d_2048", "mixed2d_64x32", "mixed3d_32x16x8"])
def test_parity_with_compiled_reference(built, name):
    case = CASES[name]
    lib = ORACLE_U if case.get("oracle") == "U" else ORACLE_F
    got = cases.run_case(case)                       # product, CUDA
    want = cases.run_case(case, lib=lib, device=0)   # reference CPU path
    # Per-field tolerance: TOL (1e-5) unless the case states the float32 floor of a derived field (case["tol"], measured by
    # tests/golden/f32_floor.py: numpy float32 restatement vs float64 -- no float32 pipeline gets closer than that).
    errs = {}
    for f, _ in case["fields"]:
        assert np.linalg.norm(want[f]) > 0
        errs[f] = rel_l2(got[f], want[f])
    bad = {f: e for f, e in errs.items() if not e < case.get("tol", {}).get(f, TOL)}
    assert not bad, (bad, errs)


@pytest.mark.parametrize("name", ["fcb_band_diffusion_1d_64", "fcb_band_allen_cahn_2d_64"])
def test_fourier_callback_device_flavour_equals_host_flavour(built, name):
    """RUN_GPU flavour: callbackFourier receives the DEVICE pointer of the full spectrum (comp_array_d, src/field.cpp:53-54);
    the built-in device callback zeroes the same band with cudaMemset2D that the host callback zeroes in a loop."""
    case = dict(CASES[name])
    host = cases.run_case(case)                                   # RUN_CPU flavour of the product (host function)
    case["device"], case["fourier_device_flavour"] = 1, True
    dev = cases.run_case(case)                                    # RUN_GPU flavour (device pointer)
    for f, _ in case["fields"]:
        assert np.isfinite(dev[f]).all()
        if case["shape"][1] == 1:
            assert np.array_equal(dev[f], host[f]), (f, rel_l2(dev[f], host[f]))
    if case["shape"][1] > 1:   # 2-D: the two flavours differ by the dealias rule, so the witness is ORACLE-F + host callback
        want = cases.run_case(dict(CASES[name]), lib=ORACLE_F, device=0)
        for f, _ in case["fields"]:
            assert rel_l2(dev[f], want[f]) < TOL, (f, rel_l2(dev[f], want[f]))


@pytest.mark.parametrize("name", ["ch3d_32", "ch2d_64", "modelh_32"])
def test_reference_gpu_path_agrees_with_oracle_f_and_product(built, name):
    """ORACLE-F (CPU sources + two one-token fixes) is meant to reproduce the semantics of the reference's GPU
    kernels.  Here the reference's OWN cuFFT/cuRAND build (unmodified sources, oracle/_ref/libcupss_ref_gpu.so) runs on the
    same input as a third witness.  The reference's two paths do not agree with EACH OTHER to the 1e-5 of the north star
    (measured: 5.4e-5 on ch3d_32 after 100 steps -- FMA contraction and cuFFT vs FFTW-order rounding, amplified by the
    linearly unstable dynamics), so this is a loose check (3e-4) that the one-token fixes of ORACLE-F reproduce the GPU
    kernels' semantics -- a wrong dealias mask or a wrong constraint-field denominator shows up at 1e-2.
    The binding parity gate is the CPU path (test_parity_with_compiled_reference)."""
    if not os.path.exists(cases.REF_GPU):
        pytest.skip("oracle/_ref/libcupss_ref_gpu.so not built")
    case = CASES[name]
    ref_gpu = cases.run_case(case, lib=cases.REF_GPU, device=1)
    ref_cpu = cases.run_case(case, lib=ORACLE_F, device=0)
    got = cases.run_case(case)
    for f, _ in case["fields"]:
        tol = 3e-4
        assert rel_l2(ref_cpu[f], ref_gpu[f]) < tol, ("oracle-F vs reference GPU", f, rel_l2(ref_cpu[f], ref_gpu[f]))
        assert rel_l2(got[f], ref_gpu[f]) < tol, ("product vs reference GPU", f, rel_l2(got[f], ref_gpu[f]))


def test_parity_with_committed_reference_outputs(built):
    gold = np.load(os.path.join(cases.GOLDEN, "ref_runs.npz"))
    for name in ["ch3d_32", "modelh_32", "kpz3d_32_det"]:
        got = cases.run_case(CASES[name])
        for f, arr in got.items():
            assert rel_l2(arr, gold[f"{name}/{f}"]) < TOL, (name, f)


def test_cahn_hilliard_3d_full_size_matches_the_reference(built):
    """BASELINE.json configs[2] at its FULL size (512^3) against the reference's own CPU path run in the build container
    (tests/golden/make_golden_fullsize.py, ORACLE-F, 12 steps from the seeded smooth IC): 32^3 point samples, three full
    planes and whole-field statistics, relative L2 <= 1e-5."""
    path = os.path.join(cases.GOLDEN, "ch3d_512_ref.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ch3d_512_ref.npz not generated")
    sys.path.insert(0, cases.GOLDEN)
    from make_golden_fullsize import summarise
    gold = np.load(path)
    case = dict(CASES["ch3d_32"])
    case["shape"], case["steps"] = (512, 512, 512), int(gold["steps"][0])
    got = summarise(cases.run_case(case)["phi"])
    for k in ("sub", "plane_z", "plane_y", "plane_x"):
        assert np.linalg.norm(gold[k]) > 0
        assert rel_l2(got[k], gold[k]) < TOL, (k, rel_l2(got[k], gold[k]))
    want = gold["stats"]
    assert abs(got["stats"][0] - want[0]) < 2e-7 + 1e-5 * abs(want[0])            # mean (conserved)
    for i in (1, 4):                                                                 # L2 norm, sum |phi|^3
        assert abs(got["stats"][i] - want[i]) < 1e-5 * abs(want[i]), (i, got["stats"][i], want[i])
    for i in (2, 3):                                                                 # extrema
        assert abs(got["stats"][i] - want[i]) < 1e-4 * abs(want[i]), (i, got["stats"][i], want[i])


def test_cahn_hilliard_3d_256_hundred_steps_matches_the_reference(built):
    """North star, literally: relative L2 <= 1e-5 per field after 100 steps, at a non-toy 3-D size.  CH-3D 256^3 x 100 steps
    against the reference CPU path (ORACLE-F) run in the build container (tests/golden/make_golden_fullsize.py 100 256:
    336 s on 8 cores): 32^3 point samples, three planes, whole-field statistics."""
    path = os.path.join(cases.GOLDEN, "ch3d_256_ref.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ch3d_256_ref.npz not generated")
    sys.path.insert(0, cases.GOLDEN)
    from make_golden_fullsize import summarise
    gold = np.load(path)
    assert int(gold["steps"][0]) == 100
    case = dict(CASES["ch3d_32"])
    case["shape"], case["steps"] = (256, 256, 256), 100
    got = summarise(cases.run_case(case)["phi"])
    for k in ("sub", "plane_z", "plane_y", "plane_x"):
        assert gold[k].shape == got[k].shape and np.linalg.norm(gold[k]) > 0
        assert rel_l2(got[k], gold[k]) < TOL, (k, rel_l2(got[k], gold[k]))
    want = gold["stats"]
    assert abs(got["stats"][0] - want[0]) < 2e-7 + 1e-5 * abs(want[0])
    for i in (1, 4):
        assert abs(got["stats"][i] - want[i]) < 1e-5 * abs(want[i]), (i, got["stats"][i], want[i])
    for i in (2, 3):
        assert abs(got["stats"][i] - want[i]) < 1e-4 * abs(want[i]), (i, got["stats"][i], want[i])


def test_generic_and_lean_kstage_agree_bitwise(built):
    """KS_SCALAR_Q2 and the generic interpreter implement the same IEEE operation sequence."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, cases\n"
            "out = cases.run_case(cases.CASES['ch3d_64x32x16'])\n"
            "np.save(sys.argv[1], out['phi'])\n") % (ROOT, os.path.join(ROOT, "tests"))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        a, b = os.path.join(d, "a.npy"), os.path.join(d, "b.npy")
        subprocess.run([sys.executable, "-c", code, a], check=True, cwd=ROOT)
        subprocess.run([sys.executable, "-c", code, b], check=True, cwd=ROOT, env=dict(os.environ, CUPSS_B200_GENERIC_KSTAGE="1"))
        assert np.array_equal(np.load(a), np.load(b))


@pytest.mark.parametrize("name", ["modelh_32", "kpz3d_32_det", "bc_even_inhomogeneous_64"])
def test_plan_specialised_kstage_matches_interpreter(built, name):
    """Generic sweeps: the k stage compiled at run time for the plan's structure (NVRTC, CUPSS_B200_JIT=1) and the
    interpreter over the same descriptors (CUPSS_B200_JIT=0) execute the same IEEE operation sequence."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, cases\n"
            "out = cases.run_case(cases.CASES[%r])\n"
            "np.savez(sys.argv[1], **out)\n") % (ROOT, os.path.join(ROOT, "tests"), name)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        a, b = os.path.join(d, "a.npz"), os.path.join(d, "b.npz")
        ra = subprocess.run([sys.executable, "-c", code, a], cwd=ROOT, env=dict(os.environ, CUPSS_B200_JIT="1"), capture_output=True, text=True)
        assert ra.returncode == 0, ra.stderr[-2000:]
        assert "not specialised" not in ra.stderr, ra.stderr[-2000:]
        subprocess.run([sys.executable, "-c", code, b], check=True, cwd=ROOT, env=dict(os.environ, CUPSS_B200_JIT="0"))
        A, B = np.load(a), np.load(b)
        for f in A.files:
            assert rel_l2(A[f], B[f]) < 1e-6, (f, rel_l2(A[f], B[f]))


def test_step0_quirk_nonlinear_term_is_skipped_on_first_step(built):
    """real_dealiased is zero until a field's first setRHS: b = 1 and b = 0 agree after one step."""
    c1 = dict(CASES["ch2d_64"])
    c0 = dict(c1, params=dict(a=-1.0, b=0.0, k=4.0))
    a = cases.run_case(c1, steps=1)["phi"]
    b = cases.run_case(c0, steps=1)["phi"]
    assert np.array_equal(a, b)
    assert not np.array_equal(cases.run_case(c1, steps=2)["phi"], cases.run_case(c0, steps=2)["phi"])


def test_diffusion_closed_form_full_size(built):
    """Config-02-sized grid (4096^2): every mode decays by 1/(1 + dt D q^2) per step."""
    from cupss_b200.capi import Evolver
    n, dt, steps = 4096, 0.1, 20
    ev = Evolver(1, n, n, 1, 1.0, 1.0, 1.0, dt)
    ev.createField("phi", True)
    ev.addParameter("D", 1.0)
    ev.addEquation("dt phi + D * q^2 * phi = 0")
    y, x = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    modes = [(3, 5, 1.0), (40, 0, 0.5), (0, 700, 0.25)]
    ic = sum(a * np.cos(2 * np.pi * (mx * x + my * y) / n) for mx, my, a in modes).astype(np.float32)
    ev.setReal("phi", ic[None])
    ev.prepareProblem()
    ev.advanceTime(steps)
    ev.copyAllDataToHost()
    want = sum(a * np.cos(2 * np.pi * (mx * x + my * y) / n) / (1 + dt * ((2 * np.pi * mx / n) ** 2 + (2 * np.pi * my / n) ** 2)) ** steps
               for mx, my, a in modes)
    assert rel_l2(ev.real("phi")[0], want) < 2e-6
    ev.close()


def test_cahn_hilliard_3d_full_size_mass_conservation(built):
    """Benchmark configuration (512^3): the q = 0 mode is invariant, the field stays finite and bounded."""
    from cupss_b200.capi import Evolver
    n = 512
    ev = Evolver(1, n, n, n, 1.0, 1.0, 1.0, 0.01)
    ev.createField("phi", True)
    for k, v in cases.CH_PARAMS.items():
        ev.addParameter(k, v)
    ev.addEquation("dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 ")
    rng = np.random.default_rng(3)
    ic = (0.2 + 0.3 * (2 * rng.random((n, n, n), dtype=np.float32) - 1)).astype(np.float32)
    ev.setReal("phi", ic)
    ev.prepareProblem()
    ev.advanceTime(20)
    ev.copyAllDataToHost()
    out = ev.real("phi")
    assert np.isfinite(out).all()
    assert abs(float(out.mean(dtype=np.float64)) - float(ic.mean(dtype=np.float64))) < 1e-6
    assert float(np.abs(out).max()) < 1.0 and out.std() < ic.std()
    ev.close()


def test_noise_variance_and_spectrum(built):
    """Stochastic runs match statistically: dt h = 0 + noise 2D gives site variance A*dt*n/dV; conserved noise
    2*D*q^2 gives a per-step increment spectrum <|dh_q|^2>/N = 2 D q^2 dt / dV."""
    from cupss_b200.capi import Evolver
    n, dt, D, dx = 64, 0.01, 0.5, 0.5
    ev = Evolver(1, n, n, 1, dx, dx, 1.0, dt)
    ev.createField("h", True)
    ev.addParameter("D", D)
    ev.addEquation("dt h = 0")
    ev.addNoise("h", "2*D")
    ev.setNoiseSeed(1234)
    ev.prepareProblem()
    steps = 50
    ev.advanceTime(steps)
    ev.copyAllDataToHost()
    var = float(ev.real("h").var())
    expect = 2 * D * dt * steps / (dx * dx)
    assert abs(var / expect - 1) < 0.06, (var, expect)          # 4096 sites: sigma ~ 2.2 %
    assert abs(float(ev.real("h").mean())) < 5 * np.sqrt(expect / n / n)
    ev.close()

    ev = Evolver(1, n, n, 1, dx, dx, 1.0, dt)
    ev.createField("p", True)
    ev.addParameter("D", D)
    ev.addEquation("dt p = 0")
    ev.addNoise("p", "2*D*q^2")
    ev.setNoiseSeed(99)
    ev.prepareProblem()
    prev = np.zeros((n, n), np.complex128)
    acc = np.zeros((n, n))
    reps = 200
    for _ in range(reps):
        ev.advanceTime(1)
        ev.copyAllDataToHost()
        cur = np.fft.fft2(ev.real("p")[0].astype(np.float64))
        acc += np.abs(cur - prev) ** 2
        prev = cur
    q = 2 * np.pi * np.fft.fftfreq(n, d=dx)
    q2 = q[None, :] ** 2 + q[:, None] ** 2
    expect = 2 * D * q2 * dt / (dx * dx) * n * n
    ratio = (acc / reps)[q2 > 0] / expect[q2 > 0]
    assert abs(float(ratio.mean()) - 1) < 0.02, float(ratio.mean())
    assert float(ratio.std()) < 0.25                              # per-mode chi^2 scatter ~ sqrt(1/200..2/200)
    ev.close()


def _noise_evolver(shape, dx, dt, amp_expr, seed, field="h"):
    from cupss_b200.capi import Evolver
    ev = Evolver(1, *shape, dx, dx, dx, dt)
    ev.createField(field, True)
    ev.addParameter("D", 0.5)
    ev.addEquation(f"dt {field} = 0")
    ev.addNoise(field, amp_expr)
    ev.setNoiseSeed(seed)
    ev.prepareProblem()
    return ev


@pytest.mark.parametrize("shape", [(32, 32, 32), (64, 16, 32)])
def test_noise_3d_variance_spectrum_and_hermitian_planes(built, shape):
    """The 3-D noise path of BASELINE.json configs[4] (z-axis k stage with one Philox call per column pair), checked where
    a C2R download cannot hide anything: on the SPECTRUM the engine holds (download_comp, the full comp_array layout).
      * per-step increments from rest, `dt h = 0` + addNoise("h", "2*D"): <|dh_q|^2> = A dt N / dV for every mode, with the
        mean taken separately over the kx, ky and kz axes of the spectrum and over the bulk;
      * the planes kx = 0 and kx = sx/2 are Hermitian-consistent: comp[kz, ky, kx] == conj(comp[-kz, -ky, kx]) exactly, and
        self-conjugate bins have zero imaginary part and carry the whole variance N in their real part;
      * site variance of the real field A dt n / dV.
    (src/field.cpp:300-330, src/field_init.cpp:269-278; the reference FFTs real white noise, which has exactly this structure.)"""
    sx, sy, sz = shape
    D, dt, dx = 0.5, 0.01, 0.5
    dV = dx ** 3
    ev = _noise_evolver(shape, dx, dt, "2*D", 4321)
    N = sx * sy * sz
    expect = 2 * D * dt / dV * N            # E|dh_q|^2 per step
    reps = 150
    acc = np.zeros((sz, sy, sx))
    acc_self_re2 = 0.0
    prev = np.zeros((sz, sy, sx), np.complex128)
    selfbins = [(kz, ky, kx) for kz in (0, sz // 2) for ky in (0, sy // 2) for kx in (0, sx // 2)]
    for rep in range(reps):
        ev.advanceTime(1)
        ev.copyAllDataToHost()
        cur = ev.comp("h").astype(np.complex128)
        inc = cur - prev
        prev = cur
        acc += np.abs(inc) ** 2
        for kx in (0, sx // 2):
            plane = inc[:, :, kx]
            mirror = np.conj(np.roll(np.roll(plane[::-1, ::-1], 1, axis=0), 1, axis=1))   # (kz, ky) -> (-kz, -ky)
            assert np.array_equal(plane, mirror), f"plane kx={kx} is not Hermitian-consistent at step {rep}"
        for b in selfbins:
            assert inc[b].imag == 0.0, (b, inc[b])
            acc_self_re2 += inc[b].real ** 2
        # the full spectrum the engine hands out is the spectrum of a REAL field: X(-k) = conj X(k) everywhere
        if rep == 0:
            full_mirror = np.conj(np.roll(np.roll(np.roll(cur[::-1, ::-1, ::-1], 1, 0), 1, 1), 1, 2))
            assert np.array_equal(cur, full_mirror)
            real = ev.real("h").astype(np.float64)
            assert rel_l2(np.fft.fftn(real), cur) < 2e-6   # comp_array IS the transform of real_array
    ratio = acc / reps / expect
    # each mode: chi^2 with 2*reps (ordinary) or reps (self-conjugate) degrees of freedom -> relative sd 1/sqrt(reps) resp. sqrt(2/reps)
    assert abs(ratio.mean() - 1) < 5 / np.sqrt(reps * N), ratio.mean()
    for axis_name, line in (("kx", ratio[0, 0, 1:sx // 2]), ("ky", ratio[0, 1:sy // 2, 0]), ("kz", ratio[1:sz // 2, 0, 0])):
        assert abs(line.mean() - 1) < 5 / np.sqrt(reps * line.size), (axis_name, line.mean())
        assert np.all(np.abs(line - 1) < 6 / np.sqrt(reps)), (axis_name, line.min(), line.max())
    assert abs(acc_self_re2 / (reps * len(selfbins)) / expect - 1) < 5 * np.sqrt(2.0 / (reps * len(selfbins)))
    assert float(np.abs(ratio - 1).max()) < 8 * np.sqrt(2.0 / reps)   # no mode is off (a dead or doubled bin would read 0 or 2)
    var = float(ev.real("h").var())
    assert abs(var / (2 * D * dt * reps / dV) - 1) < 5 * np.sqrt(2.0 / N) + 0.01, var
    ev.close()


def test_noise_3d_conserved_amplitude_follows_q2(built):
    """Conserved noise 2*D*q^2 on a 3-D grid with different extents per axis: <|dp_q|^2>/N = 2 D q^2 dt / dV along every
    axis (the |q| factor of precomp_noise, src/field_init.cpp:269-278, with qz included)."""
    sx, sy, sz = 32, 16, 64
    D, dt, dx = 0.5, 0.01, 0.5
    ev = _noise_evolver((sx, sy, sz), dx, dt, "2*D*q^2", 77, field="p")
    N = sx * sy * sz
    reps = 120
    acc = np.zeros((sz, sy, sx))
    prev = np.zeros((sz, sy, sx), np.complex128)
    for _ in range(reps):
        ev.advanceTime(1)
        ev.copyAllDataToHost()
        cur = ev.comp("p").astype(np.complex128)
        acc += np.abs(cur - prev) ** 2
        prev = cur
    qx, qy, qz = (2 * np.pi * np.fft.fftfreq(n, d=dx) for n in (sx, sy, sz))
    q2 = qz[:, None, None] ** 2 + qy[None, :, None] ** 2 + qx[None, None, :] ** 2
    expect = 2 * D * q2 * dt / dx ** 3 * N
    m = q2 > 0
    ratio = (acc / reps)[m] / expect[m]
    assert abs(ratio.mean() - 1) < 5 / np.sqrt(reps * m.sum()), ratio.mean()
    assert float(np.abs(ratio - 1).max()) < 8 * np.sqrt(2.0 / reps)
    assert acc[0, 0, 0] == 0.0   # q = 0 receives nothing: the noise conserves the mean
    ev.close()


_NOISY_CODE = ("import sys, os; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
               "import numpy as np\n"
               "from cupss_b200.capi import Evolver\n"
               "shape = tuple(int(v) for v in sys.argv[1].split('x')); kind = sys.argv[2]\n"
               "ev = Evolver(1, *shape, 1.0, 1.0, 1.0, 0.01)\n"
               "if kind == 'kpz':\n"
               "    fields = [('h', 1), ('iqxh', 0), ('iqyh', 0)] + ([('iqzh', 0)] if shape[2] > 1 else [])\n"
               "    for f, d in fields: ev.createField(f, d)\n"
               "    ev.addParameter('l', 0.5); ev.addParameter('D', 0.5)\n"
               "    ev.addEquation('dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2' + (' + l*iqzh^2' if shape[2] > 1 else ''))\n"
               "    for f, _ in fields[1:]: ev.addEquation('%%s = %%s*h' %% (f, f[:3]))\n"
               "    ev.addNoise('h', '2*D')\n"
               "else:\n"   # conserved noise on a Cahn-Hilliard field (Model B): amplitude ~ sqrt(q^2), fused inverse in the same kernel
               "    ev.createField('phi', 1)\n"
               "    ev.addParameter('D', 0.1)\n"
               "    ev.addEquation('dt phi + q^2*(-1 + 4*q^2)*phi = - q^2*phi^3')\n"
               "    ev.addNoise('phi', '2*D*q^2')\n"
               "rng = np.random.default_rng(3)\n"
               "name = 'h' if kind == 'kpz' else 'phi'\n"
               "ev.setReal(name, (0.2 * (2 * rng.random(shape[::-1]) - 1)).astype(np.float32))\n"
               "ev.setNoiseSeed(4242)\n"
               "ev.prepareProblem(); ev.advanceTime(12); ev.copyAllDataToHost()\n"
               "np.save(sys.argv[3], ev.real(name)); ev.close()\n") % (ROOT, os.path.join(ROOT, "tests"))


@pytest.mark.parametrize("shape,kind", [("64x32x32", "kpz"), ("32x16x1024", "kpz"), ("64x64x1", "kpz"), ("64x32x32", "modelb"), ("32x2048x1", "modelb")])
def test_lean_noisy_kstage_equals_the_generic_evaluator_bitwise(built, tmp_path, shape, kind):
    """A noisy field whose prefactors depend on q^2 only runs the lean evaluator compiled at run time for its signature
    (kernels_axis.cuh: NOISE); the generic evaluator (interpreter / plan-specialised) draws the same numbers and rounds the
    same way: bit-identical fields, incl. the cluster-shared axes and the conserved (q^2) amplitude."""
    out = {}
    for tag, env in (("lean", {}), ("generic", {"CUPSS_B200_NO_LEAN_NOISE": "1"}), ("interp", {"CUPSS_B200_NO_LEAN_NOISE": "1", "CUPSS_B200_JIT": "0"})):
        e = dict(os.environ); e.update(env)
        path = str(tmp_path / (tag + ".npy"))
        r = subprocess.run([sys.executable, "-c", _NOISY_CODE, shape, kind, path], env=e, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        out[tag] = np.load(path)
    assert np.isfinite(out["lean"]).all() and out["lean"].std() > 0
    assert np.array_equal(out["generic"], out["interp"])
    assert np.array_equal(out["lean"], out["generic"]), rel_l2(out["lean"], out["generic"])


def test_noise_stream_is_reproducible_and_seed_dependent(built):
    from cupss_b200.capi import Evolver

    def run(seed):
        ev = Evolver(1, 32, 32, 32, 1.0, 1.0, 1.0, 0.01)
        ev.createField("h", True)
        ev.addEquation("dt h + 0.5*q^2*h = 0")
        ev.addNoise("h", "1.0")
        ev.setNoiseSeed(seed)
        ev.prepareProblem()
        ev.advanceTime(5)
        ev.copyAllDataToHost()
        out = ev.real("h")
        ev.close()
        return out
    a, b, c = run(7), run(7), run(8)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(float(a.mean())) < 0.05


def test_update_parameter_rebakes_the_plan(built):
    """evolver::updateParameter (src/evolver.cpp:386-394) on product and reference."""
    def run(lib, device):
        ev = cases.build_system(CASES["ch2d_64"], lib=lib, device=device)
        ev.prepareProblem()
        ev.advanceTime(10)
        ev.updateParameter("b", 2.0)
        ev.updateParameter("k", 3.0)
        ev.advanceTime(10)
        if device:
            ev.copyAllDataToHost()
        out = ev.real("phi")
        ev.close()
        return out
    cwd = os.getcwd()
    assert rel_l2(run(None, 1), run(ORACLE_F, 0)) < TOL
    assert os.getcwd() == cwd


def _write_out_files(lib, device, shape, tmp, steps):
    """Runs writeOut() at step 0 and after `steps` steps in directory tmp (a subprocess: the writer uses the cwd); returns {file: bytes}."""
    code = ("import sys, os; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, cases\n"
            "from cupss_b200.capi import Evolver\n"
            "lib = None if sys.argv[1] == 'product' else sys.argv[1]\n"
            "dev, sx, sy, sz, steps = (int(v) for v in sys.argv[2:7])\n"
            "ev = Evolver(dev, sx, sy, sz, 1.0, 1.0, 1.0, 0.05, lib=lib)\n"
            "ev.createField('phi', True); ev.createField('gphi', False)\n"
            "ev.addParameter('D', 0.7)\n"
            "ev.addEquation('dt phi + D*q^2*phi = - 0.5*q^2*phi^3'); ev.addEquation('gphi = iqx*phi')\n"
            "ev.setOutputField('phi', True); ev.setOutputField('gphi', True)\n"
            "rng = np.random.default_rng(11)\n"
            "ev.setReal('phi', rng.integers(-48, 49, size=(sz, sy, sx)) / 64.0)   # multiples of 1/64: exact in %%.6f\n"
            "ev.prepareProblem()\n"
            "ev.writeOut()\n"
            "ev.advanceTime(steps)\n"
            "ev.writeOut()\n"
            "ev.setWritePrecision(3)\n"
            "ev.advanceTime(1)\n"
            "ev.writeOut()\n") % (ROOT, os.path.join(ROOT, "tests"))
    os.makedirs(tmp, exist_ok=True)
    r = subprocess.run([sys.executable, "-c", code, lib or "product", str(device), *[str(v) for v in shape], str(steps)], cwd=tmp, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = {}
    for f in sorted(os.listdir(os.path.join(tmp, "data"))):
        if ".csv." in f:
            out[f] = open(os.path.join(tmp, "data", f), "rb").read()
    return out


@pytest.mark.parametrize("shape", [(16, 1, 1), (16, 8, 1), (8, 4, 4)])
def test_output_files_match_the_reference_byte_for_byte(built, tmp_path, shape):
    """field::writeToFile (src/field.cpp:350-402) through evolver::writeOut on product and on the reference CPU path
    (ORACLE-F): same file names, the same header, index columns and line count, and
      * at step 0 (initial condition in multiples of 1/64) the files are identical byte for byte at writePrecision 6;
      * after 6 steps every line carries the same indices and a value within 1.5e-6 of the reference's (the last printed
        digit can round either way for 1e-7 differences between two float32 implementations: a value of O(0.1-1) printed
        with 6 decimals flips its last digit with probability ~|difference| / 1e-6, i.e. a few per cent of the lines);
        at least 90 % of the lines are identical;
      * after one more step at writePrecision 3 the files are again identical byte for byte."""
    steps = 6
    got = _write_out_files(None, 1, shape, str(tmp_path / "product"), steps)
    want = _write_out_files(ORACLE_F, 0, shape, str(tmp_path / "reference"), steps)
    assert sorted(got) == sorted(want) and len(got) == 6, (sorted(got), sorted(want))
    for name in sorted(want):
        g, w = got[name].decode().splitlines(), want[name].decode().splitlines()
        assert g[0] == w[0] and len(g) == len(w), name
        step = int(name.rsplit(".", 1)[1])
        if step == 0 and name.startswith("phi"):
            assert got[name] == want[name], name
            continue
        same = 0
        tol = 1.5e-6 if step <= steps else 1.1e-3
        for a, b in zip(g[1:], w[1:]):
            pa, pb = a.split(", "), b.split(", ")
            assert pa[:-1] == pb[:-1], (name, a, b)
            assert len(pa[-1].split(".")[1]) == len(pb[-1].split(".")[1])   # same number of printed digits
            assert abs(float(pa[-1]) - float(pb[-1])) <= tol, (name, a, b)
            same += a == b
        assert same >= 0.90 * (len(w) - 1), (name, same, len(w) - 1)
        if step > steps:
            assert same >= 0.995 * (len(w) - 1), (name, same, len(w) - 1)


_VARIANT_CODE = ("import sys, os; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
                 "import numpy as np, cases\n"
                 "out = cases.run_case(cases.CASES[sys.argv[1]])\n"
                 "np.savez(sys.argv[2], **out)\n") % (ROOT, os.path.join(ROOT, "tests"))


def _run_variant(name, env, path):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", _VARIANT_CODE, name, path], env=e, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return dict(np.load(path))


@pytest.mark.parametrize("name,env,bitwise", [
    ("ch3d_32x64x64", {"CUPSS_B200_TMA": "0"}, True),             # TMA box loads vs per-thread cp.async: the same tile, the same arithmetic
    ("ch3d_64x32x16", {"CUPSS_B200_TMA": "0"}, True),
    ("modelh_32", {"CUPSS_B200_TMA": "0"}, True),
    ("ch3d_64x32x16", {"CUPSS_B200_ZCHUNK": "4"}, True),           # z-chunked x -> forward-y pipeline on two lanes
    ("ch3d_64x32x16", {"CUPSS_B200_ZCHUNK": "8", "CUPSS_B200_ZCHUNK_LANES": "1"}, True),
    ("kpz3d_32_det", {"CUPSS_B200_ZCHUNK": "8"}, True),
    ("ch3d_512x8x8", {"CUPSS_B200_X3_NOPRUNE": "1"}, False),      # pruned strided level of the x pass: other rounding, same transform
    ("ch2d_4096x32", {"CUPSS_B200_X4_NOPRUNE": "1"}, False),     # ... and of the three-level kernel (sx = 256, 1024, 2048, 4096)
    ("ch3d_1024x32x8", {"CUPSS_B200_X4_NOPRUNE": "1"}, False),
    ("ch3d_256x16x8", {"CUPSS_B200_X4_NOPRUNE": "1"}, False),
    ("ch3d_512x8x8", {"CUPSS_B200_NO_GRAPH": "1"}, True),
    ("mixed2d_2048x16", {"CUPSS_B200_NO_XS1": "1"}, True),         # one-job stash x pass vs the two-job kernel: the same arithmetic per line
    ("mixed3d_1024x8x8", {"CUPSS_B200_NO_XS1": "1"}, True),
    ("modelh_2048x64", {"CUPSS_B200_NO_XS1": "1"}, True),          # ... and the other grouping of the product terms
    ("modelh_2048x64", {"CUPSS_B200_XS_OB": "2", "CUPSS_B200_XS_TWG": "0", "CUPSS_B200_XS_NT": "128"}, True),   # its launch variants
    ("mixed2d_4096x8", {"CUPSS_B200_XS_OB": "2", "CUPSS_B200_XS_NT": "512"}, True),
    ("modelh_2048x64", {"CUPSS_B200_NO_FANOUT": "1"}, True),       # independent launches side by side on the lanes vs one after the other
    ("modelh_32", {"CUPSS_B200_NO_FANOUT": "1"}, True),
])
def test_environment_variants_agree_with_the_default_path(built, tmp_path, name, env, bitwise):
    """The switchable code paths (README.md: environment switches) produce the default path's result: bit for bit where the
    arithmetic is the same (other prologue, other launch structure), to 1e-6 where only the rounding differs."""
    base = _run_variant(name, {}, str(tmp_path / "base.npz"))
    var = _run_variant(name, env, str(tmp_path / "var.npz"))
    for f in base:
        if bitwise:
            assert np.array_equal(base[f], var[f]), (f, rel_l2(var[f], base[f]))
        else:
            assert rel_l2(var[f], base[f]) < 1e-6, (f, rel_l2(var[f], base[f]))


def test_copy_host_to_device_uploads_the_edited_array(built):
    """field::copyHostToDevice mid-run (src/field.cpp:337-341): the edited host real array becomes the device state."""
    case = CASES["ch3d_64x32x16"]
    ev = cases.build_system(case)
    ev.prepareProblem()
    ev.advanceTime(5)
    ev.copyAllDataToHost()
    edited = (0.5 * ev.real("phi") + 0.01).astype(np.float32)
    ev.setReal("phi", edited)
    ev.copyHostToDevice("phi")
    ev.setReal("phi", np.zeros_like(edited))      # wipe the host copy: what comes back is the device state
    ev.copyAllDataToHost()
    assert rel_l2(ev.real("phi"), edited) < 1e-6
    assert rel_l2(ev.comp("phi"), np.fft.fftn(edited.astype(np.float64))) < 1e-6
    ev.advanceTime(3)
    ev.copyAllDataToHost()
    after = ev.real("phi")
    ev.close()
    # the run continues from the edited state: same as a fresh run started from it, except for the step-0 quirk of a fresh
    # run (its products start from zero), hence one linear step of difference at most
    assert np.isfinite(after).all() and rel_l2(after, edited) < 0.2


def test_user_code_calling_the_transform_entry_points(built, tmp_path):
    """A C++ user program against inc/cupss.h that calls field::toReal / normalize / toComp / dealias between steps
    (public in the reference, inc/cupss/field.h:121-124; semantics in INTEGRATION.md): what it adds to the host real array
    reaches the device spectrum through toComp() and is still there after further steps."""
    from cupss_b200 import capi
    src = os.path.join(ROOT, "tools", "ubench", "user_entry_check.cpp")
    exe = str(tmp_path / "user_entry_check")
    libdir, engdir = os.path.dirname(capi.PRODUCT_LIB), os.path.dirname(capi.ENGINE_LIB)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-I", os.path.join(ROOT, "inc"), "-I", "/usr/local/cuda/include", src,
                        "-L", libdir, "-lcupss", "-L", engdir, "-lcupss_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
                        "-Wl,-rpath," + libdir, "-Wl,-rpath," + engdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0 and "USER_ENTRY_OK" in r.stdout, r.stdout + r.stderr


def test_multi_gpu_slab_partition_matches_single_gpu(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(ROOT, "tests", "multi_gpu_check.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", script], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_GPU_OK" in r.stdout


def _random_system(seed):
    """A random but well-posed system in the reference's grammar: one or two dynamic fields with diffusive implicit parts,
    a few constraint fields (gradients, a filtered copy, a product), random linear and product terms on the right."""
    rng = np.random.default_rng(seed)
    dim3 = bool(rng.integers(0, 2))
    shape = (32, 16, 8) if dim3 else (64, 32, 1)
    axes = ["iqx", "iqy"] + (["iqz"] if dim3 else [])
    fields = [("u", 1), ("gu", 0), ("w", 0)]
    eqs = []
    a1, a2 = axes[rng.integers(0, len(axes))], axes[rng.integers(0, len(axes))]
    two = bool(rng.integers(0, 2))
    if two:
        fields.append(("v", 1))
    rhs_u = [f"- c1*{a1}*gu*u", f"+ c2*q^2*w", f"- c3*u^{int(rng.integers(2, 4))}"]
    if two:
        rhs_u.append(f"+ c4*{a2}*v*gu")
    rng.shuffle(rhs_u)
    eqs.append("dt u + (d1*q^2 + d2*q^4)*u = " + " ".join(rhs_u[: int(rng.integers(1, len(rhs_u) + 1))]).lstrip("+ "))
    eqs.append(f"gu = {a1}*u")
    eqs.append("w*(1 + q^2) = u^2" if rng.integers(0, 2) else f"w = {a2}*gu - 0.5*u")
    if two:
        eqs.append(f"dt v + d1*q^2*v = - c1*{a2}*u*v + c2*{a1}^2*w")
    params = dict(c1=float(rng.uniform(0.2, 1.0)), c2=float(rng.uniform(0.2, 1.0)), c3=float(rng.uniform(0.2, 1.0)), c4=float(rng.uniform(0.2, 1.0)),
                  d1=float(rng.uniform(0.5, 1.5)), d2=float(rng.uniform(0.1, 0.5)))
    ic = dict(u=("smooth", (0.4, 0.04)))
    if two:
        ic["v"] = ("smooth", (0.3, 0.03))
    return dict(shape=shape, dt=0.01, fields=fields, params=params, eqs=eqs, ic=ic, steps=25)


@pytest.mark.parametrize("seed", list(range(12)))
def test_random_systems_match_the_compiled_reference(built, seed):
    """Fuzz over the plan machinery (several product groups, constraint fields with implicit parts, extra inverse
    transforms, merged prefactors): product vs the reference CPU path (ORACLE-F) on the same random system."""
    case = _random_system(seed)
    got = cases.run_case(case)
    want = cases.run_case(case, lib=ORACLE_F, device=0)
    for f, _ in case["fields"]:
        assert np.isfinite(want[f]).all(), (case["eqs"], f)
        scale = np.linalg.norm(want[f])
        if scale == 0:
            assert np.linalg.norm(got[f]) == 0
            continue
        assert rel_l2(got[f], want[f]) < 2e-5, (case["eqs"], f, rel_l2(got[f], want[f]))


def test_stochastic_run_matches_the_reference_statistically(built):
    """North star: stochastic runs must match statistically.  Edwards-Wilkinson (examples/05) on 64^2: the time-averaged
    structure factor S(q) = <|h_q|^2> of the product (Philox, generated in k-space) and of the reference CPU path
    (mt19937 white noise in real space + FFT, seeded from the clock) agree bin by bin within the sampling error,
    and both follow the discrete stationary spectrum nu^2 N / (c^2 - 1), c = 1 + dt D q^2."""
    from cupss_b200.capi import Evolver
    n, dt, D, dx = 64, 0.05, 1.0, 1.0
    warm, samples, stride = 300, 120, 5

    def spectrum(lib, device, seed=None):
        ev = Evolver(device, n, n, 1, dx, dx, 1.0, dt, lib=lib)
        ev.createField("h", True)
        ev.addParameter("D", D)
        ev.addEquation("dt h + D*q^2*h = 0")
        ev.addNoise("h", "2*D")
        if lib is None:
            ev.setNoiseSeed(seed)
        ev.prepareProblem()
        ev.advanceTime(warm)
        acc = np.zeros((n, n))
        for _ in range(samples):
            ev.advanceTime(stride)
            if device:
                ev.copyAllDataToHost()
            acc += np.abs(np.fft.fft2(ev.real("h")[0].astype(np.float64))) ** 2
        ev.close()
        return acc / samples

    sp = spectrum(None, 1, seed=2024)
    sr = spectrum(ORACLE_F, 0)
    q = 2 * np.pi * np.fft.fftfreq(n, d=dx)
    q2 = q[None, :] ** 2 + q[:, None] ** 2
    c = 1 + dt * D * q2
    nu2 = 2 * D * dt / (dx * dx)
    theory = np.where(q2 > 0, nu2 * n * n / np.maximum(c * c - 1, 1e-30), 0.0)
    # radial bins with at least 40 modes each; only modes that have relaxed within the warm-up (c^(-2*warm) << 1)
    relaxed = (q2 > 0) & (c ** (-2.0 * warm) < 1e-3)
    edges = np.linspace(np.sqrt(q2[relaxed].min()), np.sqrt(q2.max()) * 1.0001, 9)
    for lo, hi in zip(edges[:-1], edges[1:]):
        m = relaxed & (np.sqrt(q2) >= lo) & (np.sqrt(q2) < hi)
        if m.sum() < 40:
            continue
        a, b, t = sp[m].mean(), sr[m].mean(), theory[m].mean()
        # each mode's average over `samples` correlated snapshots has a relative error of a few x 1/sqrt(samples); a bin of >= 40 modes: ~3 %
        assert abs(a / b - 1) < 0.12, ("product vs reference", lo, hi, a, b)
        assert abs(a / t - 1) < 0.12 and abs(b / t - 1) < 0.12, ("vs stationary spectrum", lo, hi, a, b, t)
