"""Shared parity cases: the same seeded system is driven through the product library, the compiled reference
(oracle/_ref) and, where useful, the numpy restatement.  Inputs are deterministic functions of the shape."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_F = os.path.join(ROOT, "oracle", "_ref", "libcupss_ref_f.so")
ORACLE_U = os.path.join(ROOT, "oracle", "_ref", "libcupss_ref_u.so")
SHIM = os.path.join(ROOT, "oracle", "_ref", "libfftw_shim.so")
REF_GPU = os.path.join(ROOT, "oracle", "_ref", "libcupss_ref_gpu.so")   # the reference's own cuFFT path (needs a GPU)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def smooth_ic(sx, sy, sz, amp=0.5, noise=0.05, seed=1, kmult=1):
    """O(amp) smooth structure + white noise: exercises the nonlinear terms without sitting on the float32 floor
    (SURVEY.md Appendix C).  kmult scales the wavenumbers of the structure: on a large grid the gradients of a 2-period
    pattern are tiny and DERIVED fields (iqx*phi, stresses, velocities of Model H) would be compared at their own
    float32 round-off floor instead of at their natural scale.  A tuple gives one factor per axis (anisotropic grids)."""
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    kmx, kmy, kmz = kmult if isinstance(kmult, tuple) else (kmult, kmult, kmult)
    f = amp * np.sin(2 * np.pi * (2 * kmx) * x / sx)
    if sy > 1:
        f = f * np.cos(2 * np.pi * (3 * kmy) * y / sy)
    if sz > 1:
        f = f * np.cos(2 * np.pi * kmz * z / sz)
    return (f + noise * (2 * rng.random((sz, sy, sx)) - 1)).astype(np.float32)


def rel_l2(a, b):
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a, b = np.asarray(a, dtype=dt).ravel(), np.asarray(b, dtype=dt).ravel()
    n = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (n if n > 0 else 1.0))


CH_PARAMS = dict(a=-1.0, b=1.0, k=4.0)
MODELH_FIELDS = [("phi", 1), ("iqxphi", 0), ("iqyphi", 0), ("sigxx", 0), ("sigxy", 0), ("vx", 0), ("vy", 0), ("w", 0), ("P", 0)]
MODELH_PARAMS = dict(a=-1, b=1, k=4, eta=1, friction=0, ka=4)
MODELH_EQS = [   # examples/04_model_h/modelh_base.cpp:33-43
    "dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 -vx*iqxphi - vy*iqyphi", "iqxphi = iqx*phi", "iqyphi = iqy*phi",
    "sigxx = - 0.5*ka *iqxphi * iqxphi + 0.5*ka*iqyphi*iqyphi", "sigxy = - ka *iqxphi * iqyphi",
    "-q^2*P = (iqx*iqx-iqy*iqy)*sigxx + 2.0 * iqx*iqy*sigxy", "vx * (friction + eta*q^2) = -iqx*P + iqx*sigxx + iqy*sigxy",
    "vy * (friction + eta*q^2) = -iqy*P + iqx*sigxy - iqy*sigxx", "w = 0.5*iqx * vy - 0.5*iqy*vx "]

CASES = {
    # config 01: examples/01_diffusion/01_diffusion.cpp:7-24 at its full size (the reference's CPU-runnable case)
    "diffusion2d_256": dict(shape=(256, 256, 1), dt=0.1, fields=[("phi", 1)], params=dict(D=1.0), eqs=["dt phi + D * q^2 * phi = 0"],
                            ic=dict(phi=("droplet", (0.0, 1.0, 30.0, 5.0, 128, 128, 0))), steps=100, oracle="U"),
    # config 02 (reduced size): examples/02_cahn_hilliard/02_cahn-hilliard.cpp:17-35
    "ch2d_64": dict(shape=(64, 64, 1), dt=0.1, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3"],
                    ic=dict(phi=("smooth", (0.1, 0.01))), steps=100),
    "ch2d_64_cpu_rule": dict(shape=(64, 64, 1), dt=0.1, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3"],
                             ic=dict(phi=("smooth", (0.1, 0.01))), steps=100, oracle="U", device=0),
    # config 03 (reduced size): examples/03_cahn_hilliard_3d/03_cahn_hilliard_3d.cpp:13-23
    "ch3d_32": dict(shape=(32, 32, 32), dt=0.01, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "],
                    ic=dict(phi=("smooth", (0.5, 0.05))), steps=100),
    "ch3d_64x32x16": dict(shape=(64, 32, 16), dt=0.01, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "],
                          ic=dict(phi=("smooth", (0.5, 0.05))), steps=50),
    # the two-level x kernel (kernels_x3.cu) covers sx = 128 and 512: same systems on grids with those x extents
    "ch3d_128x16x16": dict(shape=(128, 16, 16), dt=0.01, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "],
                           ic=dict(phi=("smooth", (0.5, 0.05))), steps=100),
    "ch2d_512x16": dict(shape=(512, 16, 1), dt=0.05, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3"],
                        ic=dict(phi=("smooth", (0.4, 0.04))), steps=100),
    "ch3d_512x8x8": dict(shape=(512, 8, 8), dt=0.01, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "],
                         ic=dict(phi=("smooth", (0.5, 0.05))), steps=50),
    # the three-level x kernel (kernels_x4.cu) covers sx = 1024, 2048, 4096 (configs[1] is 4096^2)
    "ch3d_256x16x8": dict(shape=(256, 16, 8), dt=0.01, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "],
                          ic=dict(phi=("smooth", (0.5, 0.05))), steps=50),
    "ch2d_1024x32": dict(shape=(1024, 32, 1), dt=0.05, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3"],
                        ic=dict(phi=("smooth", (0.4, 0.04))), steps=100),
    "ch2d_2048x32": dict(shape=(2048, 32, 1), dt=0.05, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3"],
                        ic=dict(phi=("smooth", (0.4, 0.04))), steps=100),
    "ch2d_4096x32": dict(shape=(4096, 32, 1), dt=0.05, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3"],
                        ic=dict(phi=("smooth", (0.4, 0.04))), steps=100),
    "ch3d_1024x32x8": dict(shape=(1024, 32, 8), dt=0.01, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "],
                          ic=dict(phi=("smooth", (0.5, 0.05))), steps=50),
    "burgers1d_2048": dict(shape=(2048, 1, 1), dt=0.01, fields=[("u", 1)], params=dict(nu=0.5), eqs=["dt u + nu*q^2*u = -0.5*iqx*u^2"],
                           ic=dict(u=("smooth", (0.5, 0.05))), steps=100),
    # two monomials / other powers of ONE input through the specialised x kernels (Swift-Hohenberg-like quadratic + cubic, quartic)
    "sh2d_512x16_two_monomials": dict(shape=(512, 16, 1), dt=0.02, fields=[("u", 1)], params=dict(r=0.3, g=0.8),
                                      eqs=["dt u + (1 - 2*q^2 + q^4 - r)*u = g*u^2 - u^3"], ic=dict(u=("smooth", (0.5, 0.05))), steps=60),
    "sh3d_256x16x8_two_monomials": dict(shape=(256, 16, 8), dt=0.02, fields=[("u", 1)], params=dict(r=0.3, g=0.8),
                                        eqs=["dt u + (1 - 2*q^2 + q^4 - r)*u = g*u^2 - u^3"], ic=dict(u=("smooth", (0.5, 0.05))), steps=40),
    "quartic2d_1024x16": dict(shape=(1024, 16, 1), dt=0.02, fields=[("u", 1)], params=dict(c=0.5), eqs=["dt u + q^2*u = - c*q^2*u^4"],
                              ic=dict(u=("smooth", (0.5, 0.05))), steps=60),
    "quartic1d_128": dict(shape=(128, 1, 1), dt=0.02, fields=[("u", 1)], params=dict(c=0.5), eqs=["dt u + q^2*u = - c*q^2*u^4 + 0.25*q^2*u^2"],
                          ic=dict(u=("smooth", (0.5, 0.05))), steps=60),
    # quadratic nonlinearity through the same kernel (single monomial c*r^2)
    "burgers_like_128": dict(shape=(128, 32, 1), dt=0.01, fields=[("u", 1)], params=dict(nu=0.5), eqs=["dt u + nu*q^2*u = -0.5*iqx*u^2"],
                             ic=dict(u=("smooth", (0.5, 0.05))), steps=100),
    # user callbacks: mirror boundary conditions on a doubled domain, RUN_CPU flavour (host callbacks)
    #   examples/07_inhomogeneous_diffusion: even BC on v and on the coefficient field x, products x*iqxv and x*v
    "bc_even_inhomogeneous_64": dict(shape=(64, 64, 1), dt=0.002, fields=[("v", 1), ("iqxv", 0), ("x", 0)], params={},
                                     eqs=["dt v +0.5*q^2*v = iqx*x*iqxv + x*iqy^2*v", "iqxv = iqx*v"],
                                     ic=dict(v=("droplet", (1.0, 0.0, 8.0, 3.0, 16, 16, 0)), x=("smooth", (0.3, 0.0))), steps=30,
                                     callbacks=[("v", False), ("x", False)], device=0, oracle="U"),
    #   examples/08_neumann_dirichlet_bc: odd BC (Dirichlet) on a diffusing field
    "bc_odd_diffusion_64": dict(shape=(64, 64, 1), dt=0.05, fields=[("phi", 1)], params={}, eqs=["dt phi + q^2*phi = 0"],
                                ic=dict(phi=("droplet", (0.0, 1.0, 6.0, 2.0, 16, 16, 0))), steps=60, callbacks=[("phi", True)], device=0, oracle="U"),
    # a NON-symmetric, nonlinear real-space callback (two boundary strips clamped / scaled) on a field that a product reads: what
    # the callback leaves in the dealiased copy is not band-limited, and the reference feeds it to computeProduct unchanged
    "bc_clamp_product_64": dict(shape=(64, 32, 1), dt=0.01, fields=[("u", 1)], params=dict(nu=0.5), eqs=["dt u + nu*q^2*u = -0.5*iqx*u^2"],
                                ic=dict(u=("smooth", (0.5, 0.05))), steps=40, callbacks=[("u", 2)], device=0, oracle="U"),
    "bc_clamp_product_1d_128": dict(shape=(128, 1, 1), dt=0.01, fields=[("u", 1), ("w", 0)], params=dict(nu=0.5),
                                    eqs=["dt u + nu*q^2*u = -0.5*iqx*u^2 + 0.1*w*u", "w = iqx*u"],
                                    ic=dict(u=("smooth", (0.5, 0.05))), steps=40, callbacks=[("u", 2), ("w", 2)], device=0, oracle="U"),
    # Fourier-space callbacks (field::callbackFourier, src/field.cpp:48-57), RUN_CPU flavour (host functions):
    #   a Hermitian low-pass on the dynamic field of a Cahn-Hilliard run (the dealiased copy must follow the callback)
    "fcb_lowpass_ch2d_64": dict(shape=(64, 64, 1), dt=0.1, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3"],
                                ic=dict(phi=("smooth", (0.1, 0.01))), steps=50, fourier_callbacks=[("phi", 1 - 1)], device=0, oracle="U"),
    #   a callback that breaks the Hermitian symmetry: what survives is the real-part projection of field::normalize
    "fcb_asym_ch3d_16": dict(shape=(16, 16, 16), dt=0.01, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "],
                             ic=dict(phi=("smooth", (0.5, 0.05))), steps=50, fourier_callbacks=[("phi", 1)], device=0, oracle="U"),
    #   on a constraint field that a product reads, together with a real-space callback on the dynamic field
    "fcb_constraint_kpz2d_32": dict(shape=(32, 32, 1), dt=0.01, fields=[("h", 1), ("iqxh", 0), ("iqyh", 0)], params=dict(l=0.5),
                                    eqs=["dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2", "iqxh = iqx*h", "iqyh = iqy*h"],
                                    ic=dict(h=("smooth", (1.0, 0.1))), steps=40, fourier_callbacks=[("iqxh", 1), ("h", 0)],
                                    callbacks=[("h", False)], device=0, oracle="U"),
    #   a band of columns zeroed (1-D and 2-D; the same callback exists in a device flavour, see the test)
    "fcb_band_diffusion_1d_64": dict(shape=(64, 1, 1), dt=0.05, fields=[("phi", 1)], params={}, eqs=["dt phi + q^2*phi = 0"],
                                     ic=dict(phi=("smooth", (0.5, 0.2))), steps=20, fourier_callbacks=[("phi", 2)], device=0, oracle="U"),
    "fcb_band_allen_cahn_2d_64": dict(shape=(64, 32, 1), dt=0.05, fields=[("phi", 1)], params={}, eqs=["dt phi + (q^2 - 1)*phi = -phi^3"],
                                      ic=dict(phi=("smooth", (0.5, 0.2))), steps=40, fourier_callbacks=[("phi", 2)], device=0, oracle="U"),
    # configs 02 and 04 at sizes where the strided axis is long and the oracle still finishes in a minute (VERDICT r1 #1d, #12):
    # the cubic term through the L = 1024 k stage, Model H at 256^2
    "ch2d_1024": dict(shape=(1024, 1024, 1), dt=0.05, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3"],
                      ic=dict(phi=("smooth", (0.4, 0.04))), steps=100, threads=0),
    "ch2d_64x4096": dict(shape=(64, 4096, 1), dt=0.05, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3"],
                         ic=dict(phi=("smooth", (0.4, 0.04))), steps=100, threads=0),
    # long strided axes shared by a thread-block cluster (kernels_axis.cuh: AxisCfg<L>::CL = 2 / 4 / 8): y passes and k stage of a
    # 3-D transform (forward, pruned inverse, fused k stage with its two cross levels), and a generic (run-time compiled) k stage
    # y and z axes long enough for the TMA prologue incl. the pruned inverse (cut-off 16 rows = one box per end)
    "ch3d_32x64x64": dict(shape=(32, 64, 64), dt=0.01, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "],
                          ic=dict(phi=("smooth", (0.5, 0.05))), steps=40, threads=0),
    "ch3d_32x1024x8": dict(shape=(32, 1024, 8), dt=0.01, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "],
                           ic=dict(phi=("smooth", (0.5, 0.05))), steps=40, threads=0),
    "ch3d_32x8x2048": dict(shape=(32, 8, 2048), dt=0.01, fields=[("phi", 1)], params=CH_PARAMS, eqs=["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "],
                           ic=dict(phi=("smooth", (0.5, 0.05))), steps=40, threads=0),
    "kpz2d_128x2048_det": dict(shape=(128, 2048, 1), dt=0.01, fields=[("h", 1), ("iqxh", 0), ("iqyh", 0)], params=dict(l=0.5),
                              eqs=["dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2", "iqxh = iqx*h", "iqyh = iqy*h"],
                              ic=dict(h=("smooth", (1.0, 0.1))), steps=40, threads=0),
    "modelh_256": dict(shape=(256, 256, 1), dt=0.1, fields=MODELH_FIELDS, params=MODELH_PARAMS, eqs=MODELH_EQS,
                       ic=dict(phi=("smooth", (0.5, 0.025, 1, 4))), steps=100, threads=0,   # 8 x 12 periods: linearly unstable band, gradients O(0.1)
                       # Derived fields built from differences of nearly equal terms (projected velocity, its curl): a plain float32
                       # pipeline (numpy restatement, complex64) sits 2.4e-5 / 4.0e-5 / 1.4e-4 from float64 on vx / vy / w and 1e-5 on
                       # the stresses after 100 steps, and the reference's own float32 run 0.7e-5 / 0.8e-5 / 2.5e-5
                       # (tests/golden/f32_floor.py).  Tolerances = 1.5 x that floor; the dynamic field and its gradients keep 1e-5.
                       tol=dict(sigxx=1.5e-5, sigxy=1.5e-5, P=1.5e-5, vx=3.6e-5, vy=6e-5, w=2.1e-4)),
    # generic ("stash") x pass on long lines -- the one-job kernel (kernels_xs.cu: sx = 1024, 2048, 4096): products of DIFFERENT
    # fields, three inputs shared by two outputs (different prefactors), in 2-D and 3-D; Model H with an sx = 2048 line, where
    # phi^3 keeps the single-input kernel and the advection term gets a launch of its own (disjoint inputs)
    "mixed2d_2048x16": dict(shape=(2048, 16, 1), dt=0.01, fields=[("u", 1), ("iqxu", 0), ("iqyu", 0)], params=dict(l=0.5, m=0.25),
                            eqs=["dt u + 0.5*q^2*u = l*iqxu*iqyu + m*q^2*u*iqxu - 0.125*q^2*iqyu^2", "iqxu = iqx*u", "iqyu = iqy*u"],
                            ic=dict(u=("smooth", (1.0, 0.1, 1, 8))), steps=40, threads=0),
    "mixed2d_1024x16": dict(shape=(1024, 16, 1), dt=0.01, fields=[("u", 1), ("iqxu", 0), ("iqyu", 0)], params=dict(l=0.5, m=0.25),
                            eqs=["dt u + 0.5*q^2*u = l*iqxu*iqyu + m*q^2*u*iqxu - 0.125*q^2*iqyu^2", "iqxu = iqx*u", "iqyu = iqy*u"],
                            ic=dict(u=("smooth", (1.0, 0.1, 1, 4))), steps=40, threads=0),
    "mixed2d_4096x8": dict(shape=(4096, 8, 1), dt=0.01, fields=[("u", 1), ("iqxu", 0), ("iqyu", 0)], params=dict(l=0.5, m=0.25),
                           eqs=["dt u + 0.5*q^2*u = l*iqxu*iqyu + m*q^2*u*iqxu - 0.125*q^2*iqyu^2", "iqxu = iqx*u", "iqyu = iqy*u"],
                           ic=dict(u=("smooth", (1.0, 0.1, 1, 16))), steps=40, threads=0),
    "mixed3d_1024x8x8": dict(shape=(1024, 8, 8), dt=0.01, fields=[("u", 1), ("iqxu", 0), ("iqzu", 0)], params=dict(l=0.5, m=0.25),
                             eqs=["dt u + 0.5*q^2*u = l*iqxu*iqzu*u + m*q^2*u*iqxu", "iqxu = iqx*u", "iqzu = iqz*u"],
                             ic=dict(u=("smooth", (1.0, 0.1, 1, 4))), steps=30, threads=0),
    # the same class of system small enough for the float64 numpy restatement (tests/test_oracle.py pins the oracle's semantics of
    # products of different fields shared by two term groups with an independent implementation)
    "mixed2d_64x32": dict(shape=(64, 32, 1), dt=0.01, fields=[("u", 1), ("iqxu", 0), ("iqyu", 0)], params=dict(l=0.5, m=0.25),
                          eqs=["dt u + 0.5*q^2*u = l*iqxu*iqyu + m*q^2*u*iqxu - 0.125*q^2*iqyu^2", "iqxu = iqx*u", "iqyu = iqy*u"],
                          ic=dict(u=("smooth", (1.0, 0.1))), steps=40),
    "mixed3d_32x16x8": dict(shape=(32, 16, 8), dt=0.01, fields=[("u", 1), ("iqxu", 0), ("iqzu", 0)], params=dict(l=0.5, m=0.25),
                            eqs=["dt u + 0.5*q^2*u = l*iqxu*iqzu*u + m*q^2*u*iqxu", "iqxu = iqx*u", "iqzu = iqz*u"],
                            ic=dict(u=("smooth", (1.0, 0.1))), steps=30),
    "mixed1d_2048": dict(shape=(2048, 1, 1), dt=0.01, fields=[("u", 1), ("w", 0)], params=dict(nu=0.5),   # a single line: the pair's second line does not exist
                         eqs=["dt u + nu*q^2*u = -u*w + 0.1*q^2*w*w*u", "w = iqx*u"], ic=dict(u=("smooth", (0.5, 0.05, 1, 8))), steps=60, threads=0),
    "modelh_2048x64": dict(shape=(2048, 64, 1), dt=0.1, fields=MODELH_FIELDS, params=MODELH_PARAMS, eqs=MODELH_EQS,
                           ic=dict(phi=("smooth", (0.5, 0.025, 1, (32, 1, 1)))), steps=60, threads=0,   # 64 x 3 periods: the unstable band, as modelh_256
                           tol=dict(sigxx=1.5e-5, sigxy=1.5e-5, P=1.5e-5, vx=3.6e-5, vy=6e-5, w=2.1e-4)),
    # config 04 (reduced size): Model H, 9 fields, constraint fields with implicit LHS (needs ORACLE-F)
    "modelh_32": dict(shape=(32, 32, 1), dt=0.1, fields=MODELH_FIELDS, params=MODELH_PARAMS, eqs=MODELH_EQS,
                      ic=dict(phi=("smooth", (0.5, 0.025))), steps=100),
    # config 06 (3-D variant, noise off): h + three gradient constraint fields, merged product groups
    "kpz3d_32_det": dict(shape=(32, 32, 32), dt=0.01, fields=[("h", 1), ("iqxh", 0), ("iqyh", 0), ("iqzh", 0)], params=dict(l=0.5),
                         eqs=["dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2 + l*iqzh^2", "iqxh = iqx*h", "iqyh = iqy*h", "iqzh = iqz*h"],
                         ic=dict(h=("smooth", (1.0, 0.1))), steps=50),
    # sum-of-powers x pass of the two-level kernel (sx = 128, 512): one power of each of several inputs, one output
    "kpz3d_128x16x16_det": dict(shape=(128, 16, 16), dt=0.01, fields=[("h", 1), ("iqxh", 0), ("iqyh", 0), ("iqzh", 0)], params=dict(l=0.5),
                                eqs=["dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2 + l*iqzh^2", "iqxh = iqx*h", "iqyh = iqy*h", "iqzh = iqz*h"],
                                ic=dict(h=("smooth", (1.0, 0.1))), steps=50),
    "kpz3d_1024x16x8_det": dict(shape=(1024, 16, 8), dt=0.01, fields=[("h", 1), ("iqxh", 0), ("iqyh", 0), ("iqzh", 0)], params=dict(l=0.5),
                                eqs=["dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2 + l*iqzh^2", "iqxh = iqx*h", "iqyh = iqy*h", "iqzh = iqz*h"],
                                ic=dict(h=("smooth", (1.0, 0.1))), steps=30),
    "kpz2d_256x32_mixed_powers": dict(shape=(256, 32, 1), dt=0.01, fields=[("h", 1), ("iqxh", 0), ("iqyh", 0)], params=dict(l=0.5, m=0.25),
                                      eqs=["dt h + 0.5*q^2*h = l*iqxh^2 + m*iqyh^3", "iqxh = iqx*h", "iqyh = iqy*h"],
                                      ic=dict(h=("smooth", (1.0, 0.1))), steps=50),
    "kpz2d_512x16_mixed_powers": dict(shape=(512, 16, 1), dt=0.01, fields=[("h", 1), ("iqxh", 0), ("iqyh", 0)], params=dict(l=0.5, m=0.25),
                                      eqs=["dt h + 0.5*q^2*h = l*iqxh^2 + m*iqyh^3", "iqxh = iqx*h", "iqyh = iqy*h"],
                                      ic=dict(h=("smooth", (1.0, 0.1))), steps=50),
    # the reference's operator tests (tests/tests.cpp:191-422) extended to 3 steps
    "ops1d_16": dict(shape=(16, 1, 1), dt=0.1, fields=[("phi", 1), ("lapphi", 0), ("iqxphi", 0), ("invqphi", 0)], params={},
                     eqs=["dt phi + q^2*phi = iqxphi^2", "lapphi = -q^2*phi", "iqxphi = iqx*phi", "invqphi = 1/q*phi"],
                     ic=dict(phi=("smooth", (0.5, 0.05))), steps=3),
    "ops3d_16": dict(shape=(16, 16, 16), dt=0.1, fields=[("phi", 1), ("lapphi", 0), ("iqxphi", 0), ("invqphi", 0), ("iqyphi", 0), ("iqzphi", 0)],
                     params={}, eqs=["dt phi + q^2*phi = iqxphi^2", "lapphi = -q^2*phi", "iqxphi = iqx*phi", "invqphi = 1/q*phi",
                                     "iqyphi = iqy*phi", "iqzphi = iqz*phi"], ic=dict(phi=("smooth", (0.5, 0.05))), steps=3),
}


def build_system(case, lib=None, device=1):
    from cupss_b200.capi import Evolver
    sx, sy, sz = case["shape"]
    ev = Evolver(device, sx, sy, sz, 1.0, 1.0, 1.0, case["dt"], lib=lib)
    for n, d in case["fields"]:
        ev.createField(n, d)
    for k, v in case["params"].items():
        ev.addParameter(k, v)
    for e in case["eqs"]:
        ev.addEquation(e)
    for f, a in case.get("noise", []):
        ev.addNoise(f, a)
    for f, odd in case.get("callbacks", []):
        ev.setMirrorCallback(f, odd)
    for f, kind in case.get("fourier_callbacks", []):
        ev.setFourierCallback(f, kind, device_flavour=bool(device) and lib is None and case.get("fourier_device_flavour", False))
    for name, (kind, args) in case["ic"].items():
        if kind == "smooth":
            ev.setReal(name, smooth_ic(sx, sy, sz, *args))
        elif kind == "droplet":
            ev.initializeDroplet(name, *args)
    return ev


def set_shim_threads(lib, threads):
    """Threads of the oracle's FFT shim (0: all host cores).  Lines are independent: the result does not depend on it."""
    import ctypes as C
    h = C.CDLL(lib)
    if hasattr(h, "cupss_shim_set_threads"):
        h.cupss_shim_set_threads(int(threads) if threads else (os.cpu_count() or 1))


def run_case(case, lib=None, device=None, steps=None):
    """Returns {field: real array} after `steps` advanceTime calls."""
    if device is None:
        device = case.get("device", 1)
    if lib is not None and "threads" in case:
        set_shim_threads(lib, case["threads"])
    ev = build_system(case, lib=lib, device=device)
    ev.prepareProblem()
    ev.advanceTime(case["steps"] if steps is None else steps)
    if device:   # RUN_GPU: host mirrors are refreshed on request; RUN_CPU: they are live (and the reference's
        ev.copyAllDataToHost()   # copyDeviceToHost would overwrite them with its unused device arrays)
    out = {n: ev.real(n) for n, _ in case["fields"]}
    ev.close()
    return out


def load_truth(name):
    """A reference base_truth column as a flat float64 array (last CSV column)."""
    path = os.path.join(GOLDEN, "base_truths", name)
    vals = []
    with open(path) as f:
        for line in f:
            parts = [p.strip() for p in line.replace(",", " ").split()]
            try:
                vals.append(float(parts[-1]))
            except (ValueError, IndexError):
                continue
    return np.array(vals)
