"""CPU suite, part 2: host logic of the product (parser, initialisers, C ABI surface, failure behaviour, the
register-level FFT core run on the CPU, and the slab-exchange layout under a 2-rank gloo group)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import cases
from cases import CASES, ORACLE_F, ROOT, load_truth

EQUATION_SETS = {
    "modelh": (cases.MODELH_FIELDS, cases.MODELH_PARAMS, cases.MODELH_EQS, []),
    "ch3d": ([("phi", 1)], cases.CH_PARAMS, ["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 "], []),
    "ch2d_noise": ([("phi", 1)], dict(a=-1, b=1, k=4, D=0.01), ["dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3"], [("phi", "2*D*q^2")]),
    "kpz": ([("h", 1), ("iqh", 0)], dict(D=0.5, l=0.5), ["dt h + 0.5 * q^2 * h = l * iqh^2", "iqh = iqx*h"], [("h", "2*D")]),
    "nested": ([("u", 1), ("v", 0)], dict(a=2.0, b=3.0, c=0.5),
               ["dt u + (a - b*(q^2 - c*q^4))*u = -(u^2 - a*(v - u)*u)*iqx - 1/a*u*v/b", "v*(1/q^2 + c) = -iqy^2*u + 2.5*iqx*iqz*u"], [("v", "a*1/q^2")]),
    "zero_rhs": ([("phi", 1)], dict(D=1.0), ["dt phi + D * q^2 * phi = 0"], []),
}


def _dump(lib, fields, params, eqs, noise):
    from cupss_b200.capi import Evolver
    ev = Evolver(0 if lib else 1, 16, 16, 16, 1.0, 1.0, 1.0, 0.1, lib=lib)
    for n, d in fields:
        ev.createField(n, d)
    for k, v in params.items():
        ev.addParameter(k, v)
    for e in eqs:
        ev.addEquation(e)
    for f, a in noise:
        ev.addNoise(f, a)
    d = ev.dumpPlan()
    ev.close()
    return d


@pytest.mark.parametrize("name", sorted(EQUATION_SETS))
@pytest.mark.skipif(not os.path.exists(ORACLE_F) and not os.path.isdir("/root/reference/src"), reason="oracle not available")
def test_parser_emits_the_reference_plan(built, name):
    """Same strings through the product's parser and the reference's: identical fields, implicit monomials, grouped
    terms (order included) and noise amplitudes."""
    spec = EQUATION_SETS[name]
    assert _dump(None, *spec) == _dump(ORACLE_F, *spec)


def test_parser_known_answer_modelh(built):
    """Golden from the reference's printInformation() for Model H."""
    d = _dump(None, *EQUATION_SETS["modelh"])
    assert "field phi dynamic=1" in d
    assert "implicit {1,1,0,0,0,0} {-4,2,0,0,0,0}" in d
    assert "term {-1,1,0,0,0,0} ( phi phi phi )" in d
    assert "term {2,0,1,1,0,0} ( sigxy )" in d and "term {1,0,2,0,0,0} {-1,0,0,2,0,0} ( sigxx )" in d
    assert "implicit {0,0,0,0,0,0} {1,1,0,0,0,0}" in d   # vx: friction + eta*q^2


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_product_droplet_initialiser_matches_base_truths(built, dim):
    from cupss_b200.capi import Evolver
    ev = Evolver(1, 16, 16 if dim > 1 else 1, 16 if dim > 2 else 1, 1.0, 1.0, 1.0, 1.0)
    ev.createField("phi", True)
    ev.initializeDroplet("phi", 0, 1, 16 / 8, 4, 8, 8 if dim > 1 else 0, 0)
    assert np.max(np.abs(ev.real("phi").ravel() - load_truth(f"phi_{dim}d"))) < 1e-4
    ev.close()


def test_c_abi_library_exports_every_declared_symbol(built):
    from cupss_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "cupss_b200.h")).read()
    declared = set(re.findall(r"\b(cupss_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    import ctypes
    lib = ctypes.CDLL(capi.ENGINE_LIB)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    capi.load_engine()
    capi.load_facade(capi.PRODUCT_LIB)


def test_plan_specialised_kstage_compiles_at_run_time(built):
    """The generic k stage is compiled per plan with NVRTC from cupss_b200/csrc/*.cuh; the compile step needs no GPU.
    Guards the kernel headers against constructs NVRTC cannot take (host headers, non-constexpr plan reads)."""
    import ctypes
    from cupss_b200 import capi
    eng = capi.load_engine()
    log = ctypes.create_string_buffer(4096)
    rc = eng.cupss_b200_jit_selftest(log, len(log))
    if rc == 3:
        pytest.skip("NVRTC not installed: " + log.value.decode())
    assert rc == 0, log.value.decode()


def test_engine_kernels_are_sm100a_and_not_library_fft(built):
    from cupss_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.ENGINE_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", capi.ENGINE_LIB], capture_output=True, text=True).stdout
    assert "cufft" not in ldd and "curand" not in ldd


def test_product_fails_loudly_without_a_gpu(built):
    """No CPU fallback: on a box without a CUDA device prepareProblem must abort with a message (exit code 1)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from cupss_b200.capi import Evolver\n"
            "ev = Evolver(1, 16, 16, 1, 1.0, 1.0, 1.0, 0.1)\n"
            "ev.createField('phi', True); ev.addEquation('dt phi + q^2*phi = 0'); ev.prepareProblem()\n"
            "print('NOT REACHED')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode != 0
    assert "NOT REACHED" not in r.stdout
    assert "no CPU fallback" in (r.stdout + r.stderr)


def test_fft_core_on_the_cpu(built, tmp_path):
    """cupss_b200/csrc/fft_core.cuh is __host__ __device__: the butterflies and Stockham index arithmetic the kernels
    use are executed thread by thread on the CPU for every supported length and checked against a naive DFT."""
    exe = tmp_path / "fft_core_check"
    src = os.path.join(ROOT, "tests", "host", "fft_core_check.cu")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-x", "cu", src, "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout


def test_noise_generator_on_the_cpu(built, tmp_path):
    """cupss_b200/csrc/kstage.cuh is __host__ __device__: philox4x32_10 against the Random123 known-answer vectors, the
    Hermitian structure of white_noise_mode in the self-conjugate planes kx = 0 / sx/2 (mode == conj(mirror), real
    self-conjugate bins, E|xi|^2 = N in every class of bins) and the Box-Muller moments, all on the CPU."""
    exe = tmp_path / "noise_check"
    src = os.path.join(ROOT, "tests", "host", "noise_check.cu")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-x", "cu", src, "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout
    assert "6627e8d5 e169c58d bc57ac4c 9b00dbd8" in r.stdout   # Random123 kat_vectors: philox4x32 10, zero counter and key


def test_philox_known_answer_vectors_are_the_published_algorithm():
    """The three known-answer vectors used by tests/host/noise_check.cu, reproduced by an independent pure-Python
    Philox4x32-10 written from the published round function (Salmon et al., SC'11): multipliers 0xD2511F53 / 0xCD9E8D57,
    Weyl key increments 0x9E3779B9 / 0xBB67AE85."""
    def philox(c, k):
        c, k = list(c), list(k)
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
            k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
        return c
    assert philox((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_bench_parity_initial_condition_is_the_test_suites(built):
    """bench.py builds the 512^3 parity initial condition by broadcasting; it must be bit-identical to cases.smooth_ic."""
    sys.path.insert(0, ROOT)
    import bench
    for shape in [(32, 16, 8), (64, 32, 1), (16, 1, 1), (40, 24, 12)]:
        assert np.array_equal(cases.smooth_ic(*shape, 0.5, 0.05), bench.smooth_ic(*shape, 0.5, 0.05))


SYSTEM_FILE = """# a system file in the reference's format (src/parser.cpp:11-96)
Fields
phi 1 1
iqxphi 0 0

mu 0 1
Parameters
a -1.0
b 1
k 4.0
# comment between entries
Equations
dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 - 0.5*iqx*iqxphi*phi
iqxphi = iqx*phi
mu = a*phi + b*phi^3 + k*q^2*phi
"""


@pytest.mark.skipif(not os.path.exists(ORACLE_F), reason="oracle not built")
def test_create_from_file_matches_the_reference(built, tmp_path):
    """evolver::createFromFile (src/parser.cpp:11-96): the same Fields / Parameters / Equations file through the product's
    text front end and the reference's gives the same plan (fields, output flags via the dump of terms, parameters) and the
    same printInformation() text; the echo of the file on stdout is the same, too."""
    path = tmp_path / "system.in"
    path.write_text(SYSTEM_FILE)
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from cupss_b200.capi import Evolver\n"
            "lib = sys.argv[1] if sys.argv[1] != 'product' else None\n"
            "ev = Evolver(0, 16, 16, 1, 1.0, 1.0, 1.0, 0.1, lib=lib)\n"
            "rc = ev.createFromFile(sys.argv[2])\n"
            "sys.stdout.flush()\n"
            "print('RC', rc); print('PLAN'); print(ev.dumpPlan()); print('PARAMS', ev.getParameter('a'), ev.getParameter('b'), ev.getParameter('k'))\n"
            "sys.stdout.flush()\n"
            "print('INFO'); sys.stdout.flush(); ev.printInformation()\n") % ROOT
    outs = {}
    for tag, lib in (("product", "product"), ("reference", ORACLE_F)):
        r = subprocess.run([sys.executable, "-c", code, lib, str(path)], capture_output=True, text=True, cwd=str(tmp_path))
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = r.stdout
    assert "RC 0" in outs["product"] and "term" in outs["product"] and "( phi phi phi )" in outs["product"]
    assert outs["product"] == outs["reference"]


def test_print_information_known_answer(built, tmp_path):
    """the reference's printInformation() text for the Model H system (parser known answer)."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import cases\n"
            "from cupss_b200.capi import Evolver\n"
            "ev = Evolver(0, 16, 16, 1, 1.0, 1.0, 1.0, 0.1)\n"
            "for n, d in cases.MODELH_FIELDS: ev.createField(n, d)\n"
            "for k, v in cases.MODELH_PARAMS.items(): ev.addParameter(k, v)\n"
            "for e in cases.MODELH_EQS: ev.addEquation(e)\n"
            "ev.printInformation()\n") % (ROOT, os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout
    assert "Field 0: phi is dynamic. and has 3 explicit terms and 2 implicit terms." in out
    assert "(d/dt)phi = [+1.000000(q)^(2)-4.000000(q)^(4)]phi [ + (-1.000000)(q)^(2)] ( phi phi phi ) + [ + (-1.000000)] ( iqxphi vx ) + [ + (-1.000000)] ( iqyphi vy )" in out
    assert "sigxx =  [ + (-2.000000)] ( iqxphi iqxphi ) + [ + (2.000000)] ( iqyphi iqyphi )" in out
    assert "vx[0.000000+1.000000(q)^(2)] =  [ + (-1.000000)(iqx)^(1)] ( P ) + [ + (1.000000)(iqx)^(1)] ( sigxx ) + [ + (1.000000)(iqy)^(1)] ( sigxy )" in out
    assert "P[-1.000000(q)^(2)] =  [ + (2.000000)(iqx)^(1)(iqy)^(1)] ( sigxy ) + [ + (1.000000)(iqx)^(2) +  + (-1.000000)(iqy)^(2)] ( sigxx )" in out


def test_non_power_of_two_grid_aborts_with_a_message(built, tmp_path):
    """the engine plans powers of two up to 8192 per axis; anything else is refused by cupss_b200_create
    with a message (and the C++ layer exits with it) instead of computing something else."""
    import ctypes as C
    from cupss_b200 import capi
    eng = capi.load_engine()
    h = C.c_void_p()
    for shape in [(48, 16, 1), (16, 100, 1), (16, 16, 12), (16384, 1, 1)]:
        rc = eng.cupss_b200_create(C.byref(h), *shape, 1.0, 1.0, 1.0, 0.1)
        assert rc != 0, shape
        assert b"power of two" in eng.cupss_b200_last_error(), eng.cupss_b200_last_error()


@pytest.mark.skipif(not os.path.exists(cases.ORACLE_U), reason="oracle not built")
def test_bench_reference_arm_runs_the_named_grid(built):
    """`bench.py --impl reference` times the reference CPU path on the grid it is asked for (no sampling, no scaling)."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "32", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    assert "32^3" in line["config"]["workload"] and "scaled" not in line["cpu_baseline"]["sample"]
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["one_core"]["cores"] == 1
    assert abs(line["value"] * line["ms_per_step"] / 1e3 - 1) < 1e-9


def _slab_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sx, sy, sz = 16, 8, 8
    zl, kyl, ncol = sz // world, sy // world, sx // 2 + 1
    rng = np.random.default_rng(5)
    full = rng.standard_normal((sz, sy, sx))
    mine = full[rank * zl:(rank + 1) * zl]
    xy = np.fft.fft(np.fft.rfft(mine, axis=2), axis=1)                     # x pass then y pass on the z-slab
    # forward exchange layout written by the y pass: [peer][z_local][ky_local][kx]  (engine.cu make_y_axis); the ky rows are dealt
    # out cyclically, ky = ky_local * P + peer, so that the low-|ky| rows a dealiased inverse keeps are spread over the ranks
    send = np.ascontiguousarray(xy.reshape(zl, kyl, world, ncol).transpose(2, 0, 1, 3))
    recv = np.empty_like(send)
    t_send, t_recv = torch.from_numpy(send.view(np.float64).copy()), torch.from_numpy(recv.view(np.float64).copy())
    dist.all_to_all_single(t_recv, t_send)
    got = t_recv.numpy().view(np.complex128).reshape(world * zl, kyl, ncol)   # == natural [sz][ky_local][kx]
    spec = np.fft.fft(got, axis=0)                                           # z pass
    want = np.fft.rfftn(full)[:, rank::world, :]
    err = float(np.max(np.abs(spec - want)))
    # inverse direction: z-major chunks are contiguous; receiver reads [peer][z_local][ky_local][kx] as ky = ky_local*P + peer
    back = np.fft.ifft(spec, axis=0)
    t_send = torch.from_numpy(np.ascontiguousarray(back).view(np.float64).copy())
    t_recv = torch.empty_like(t_send)
    dist.all_to_all_single(t_recv, t_send)
    r = t_recv.numpy().view(np.complex128).reshape(world, zl, kyl, ncol).transpose(1, 2, 0, 3).reshape(zl, sy, ncol)
    real = np.fft.irfft(np.fft.ifft(r, axis=1), n=sx, axis=2)
    err2 = float(np.max(np.abs(real - mine)))
    q.put((rank, err, err2))
    dist.destroy_process_group()


def test_slab_exchange_layout_two_ranks_gloo():
    """world_size-2 CPU rendition of the multi-GPU path: z-slabs in real space, ky-slabs in Fourier space, one
    all-to-all per 3-D transform in the [peer][z_local][ky_local][kx] layout the y-pass kernels address (cyclic ky ownership)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for _, e1, e2 in res:
        assert e1 < 1e-10 and e2 < 1e-12


def test_bench_reference_arm_rank1_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


REF_EXAMPLES = "/root/reference/examples"


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference sources are only present in the build container")
def test_reference_examples_compile_unchanged(built, tmp_path):
    """Drop-in check at the source level: every example of the reference compiles and links, unedited,
    against inc/cupss.h + lib/libcupss.so.  The sources are symlinked from where they lie (never copied into the repo);
    the tree mirrors the reference's so that the examples' relative includes ("../../inc/cupss.h") resolve to OUR headers."""
    from cupss_b200 import capi
    (tmp_path / "examples").mkdir()
    os.symlink(os.path.join(ROOT, "inc"), tmp_path / "inc")
    srcs = []
    for d in sorted(os.listdir(REF_EXAMPLES)):
        full = os.path.join(REF_EXAMPLES, d)
        if not os.path.isdir(full):
            continue
        (tmp_path / "examples" / d).mkdir()
        for f in sorted(os.listdir(full)):
            if f.endswith((".cpp", ".cu")):
                os.symlink(os.path.join(full, f), tmp_path / "examples" / d / f)
                srcs.append(str(tmp_path / "examples" / d / f))
    assert len(srcs) >= 10
    libdir, engdir = os.path.dirname(capi.PRODUCT_LIB), os.path.dirname(capi.ENGINE_LIB)
    for s in srcs:
        out = s + ".bin"
        if s.endswith(".cu"):   # example 07: user kernels + callbacks, built with nvcc as its README says
            cmd = ["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-I", str(tmp_path / "inc"), s,
                   "-L", libdir, "-lcupss", "-L", engdir, "-lcupss_b200", "-o", out]
        else:
            cmd = ["g++", "-std=c++17", "-O1", "-w", "-I", str(tmp_path / "inc"), "-I", "/usr/local/cuda/include", s,
                   "-L", libdir, "-lcupss", "-L", engdir, "-lcupss_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, (s, r.stderr[-2000:])


def test_user_code_calling_the_transform_entry_points_links(built, tmp_path):
    """field::toComp / toReal / normalize / dealias are public in the reference (inc/cupss/field.h:121-124) and user code may call
    them between steps: they exist here with the semantics INTEGRATION.md states (the forwarding targets -- upload of the real
    mirror, download of the mirrors -- are covered by the GPU tests)."""
    from cupss_b200 import capi
    src = tmp_path / "user.cpp"
    src.write_text('#include <cupss.h>\n'
                   'int main() {\n'
                   '    evolver system(RUN_GPU, 32, 32, 1.0f, 1.0f, 0.1f, 10);\n'
                   '    system.createField("phi", true);\n'
                   '    system.addEquation("dt phi + q^2*phi = 0");\n'
                   '    system.prepareProblem();\n'
                   '    system.advanceTime();\n'
                   '    field *f = system.fields[0];\n'
                   '    f->toReal(); f->normalize();\n'
                   '    f->real_array[0].x += 1.0f;\n'
                   '    f->toComp(); f->dealias();\n'
                   '    return 0;\n'
                   '}\n')
    libdir, engdir = os.path.dirname(capi.PRODUCT_LIB), os.path.dirname(capi.ENGINE_LIB)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-I", os.path.join(ROOT, "inc"), "-I", "/usr/local/cuda/include", str(src),
                        "-L", libdir, "-lcupss", "-L", engdir, "-lcupss_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-o", str(tmp_path / "user.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_one_job_stash_xpass_padded_offsets_fold():
    """kernels_xs.cu addresses its padded lines as xspad(row0) + a compile-time offset per butterfly leg.  That is only the
    same element as xspad(row0 + M*q) if the low four bits never carry; checked here for every virtual thread of every level of
    the three line lengths the kernel is built for (radices as in fft_core.cuh: FftLevels), together with the claim that a half
    warp's 64-bit accesses fall into 16 different banks on the consecutive and on the innermost (stride R) patterns."""
    def xspad(i):
        return i + (i >> 4)
    for sx, rad in ((1024, (16, 8, 8)), (2048, (16, 16, 8)), (4096, (16, 16, 16))):
        n = sx
        for lv, r in enumerate(rad[:-1]):          # the shared -> shared levels and the level-0 store (the last level is contiguous)
            m = n // r
            assert m % 8 == 0 and n % 16 == 0
            for v in range(sx // r):
                blk, j = divmod(v, m)
                row0 = blk * n + j
                for q in range(r):
                    assert xspad(row0 + m * q) == xspad(row0) + m * q + ((m * q) >> 4), (sx, lv, v, q)
            n //= r
        rl = rad[-1]
        for v0 in range(0, sx // rl, 16):          # innermost level: lane v reads elements v*RL + q
            for q in range(rl):
                banks = {xspad(v * rl + q) % 16 for v in range(v0, v0 + 16)}
                assert len(banks) == 16, (sx, v0, q)
        for i0 in range(0, sx, 16):                # consecutive elements (outer levels, untangle)
            assert len({xspad(i) % 16 for i in range(i0, i0 + 16)}) == 16
