"""Developer diagnostic (run on a GPU box): product path vs the oracle builds on small grids."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cupss_b200.capi import Evolver, RUN_GPU, RUN_CPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORF = os.path.join(ROOT, "oracle/_ref/libcupss_ref_f.so")
ORU = os.path.join(ROOT, "oracle/_ref/libcupss_ref_u.so")

def rel(a, b):
    n = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (n if n > 0 else 1.0))

def smooth_ic(sx, sy, sz, amp=0.5, noise=0.05, seed=1):
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    f = amp * np.sin(2 * np.pi * 2 * x / sx)
    if sy > 1: f = f * np.cos(2 * np.pi * 3 * y / sy)
    if sz > 1: f = f * np.cos(2 * np.pi * z / sz)
    return (f + noise * (2 * rng.random((sz, sy, sx)) - 1)).astype(np.float32)

def roundtrip(sx, sy, sz):
    ev = Evolver(RUN_GPU, sx, sy, sz, 1.0, 1.0, 1.0, 0.1)
    ev.createField("phi", True)
    ev.addEquation("dt phi + q^2*phi = 0")
    ic = smooth_ic(sx, sy, sz, noise=0.5)
    ev.setReal("phi", ic)
    ev.prepareProblem()
    ev.copyAllDataToHost()
    r = ev.real("phi"); c = ev.comp("phi")
    cref = np.fft.fftn(ic.astype(np.float64))
    print(f"roundtrip {sx}x{sy}x{sz}: real {rel(r, ic):.2e}  comp {rel(c, cref):.2e}", flush=True)
    ev.close()

def build(lib, dev, shape, dt, fields, params, eqs, noise=(), dxyz=(1.0, 1.0, 1.0)):
    sx, sy, sz = shape
    ev = Evolver(dev, sx, sy, sz, dxyz[0], dxyz[1], dxyz[2], dt, lib=lib)
    for n, d in fields: ev.createField(n, d)
    for k, v in params.items(): ev.addParameter(k, v)
    for e in eqs: ev.addEquation(e)
    for f, a in noise: ev.addNoise(f, a)
    return ev

def compare(tag, shape, dt, fields, params, eqs, ic, steps, oracle=ORF, dev=RUN_GPU, report_every=None):
    a = build(None, dev, shape, dt, fields, params, eqs)
    b = build(oracle, RUN_CPU, shape, dt, fields, params, eqs)
    for name, arr in ic.items():
        a.setReal(name, arr); b.setReal(name, arr)
    a.prepareProblem(); b.prepareProblem()
    done = 0
    marks = sorted(set([1, 2, steps] + ([] if not report_every else list(range(report_every, steps, report_every)))))
    for m in marks:
        if m > steps: continue
        a.advanceTime(m - done); b.advanceTime(m - done); done = m
        a.copyAllDataToHost()
        errs = {n: rel(a.real(n), b.real(n)) for n, _ in fields}
        cerr = {n: rel(a.comp(n), b.comp(n)) for n, _ in fields}
        print(f"{tag} step {m}: real " + " ".join(f"{n}={e:.2e}" for n, e in errs.items()), flush=True)
        print(f"{tag} step {m}: comp " + " ".join(f"{n}={e:.2e}" for n, e in cerr.items()), flush=True)
    a.close(); b.close()

CH2 = ("dt phi + q^2*(a + k*q^2)*phi= - b*q^2*phi^3",)
CH3 = ("dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 ",)
CHP = dict(a=-1.0, b=1.0, k=4.0)
MH_F = [("phi", 1), ("iqxphi", 0), ("iqyphi", 0), ("sigxx", 0), ("sigxy", 0), ("vx", 0), ("vy", 0), ("w", 0), ("P", 0)]
MH_P = dict(a=-1, b=1, k=4, eta=1, friction=0, ka=4)
MH_E = ["dt phi + ( a *q^2 + k*q^4)*phi= - b* q^2* phi^3 -vx*iqxphi - vy*iqyphi", "iqxphi = iqx*phi", "iqyphi = iqy*phi",
        "sigxx = - 0.5*ka *iqxphi * iqxphi + 0.5*ka*iqyphi*iqyphi", "sigxy = - ka *iqxphi * iqyphi",
        "-q^2*P = (iqx*iqx-iqy*iqy)*sigxx + 2.0 * iqx*iqy*sigxy", "vx * (friction + eta*q^2) = -iqx*P + iqx*sigxx + iqy*sigxy",
        "vy * (friction + eta*q^2) = -iqy*P + iqx*sigxy - iqy*sigxx", "w = 0.5*iqx * vy - 0.5*iqy*vx "]

if __name__ == "__main__":
    which = sys.argv[1:] or ["rt", "ops", "ch2", "ch3", "mh", "kpz"]
    t0 = time.time()
    if "rt" in which:
        for s in [(16, 1, 1), (64, 1, 1), (16, 16, 1), (64, 32, 1), (16, 16, 16), (32, 16, 64), (128, 128, 1), (64, 64, 64)]:
            roundtrip(*s)
    if "ops" in which:
        for shape in [(16, 1, 1), (16, 16, 16)]:
            f = [("phi", 1), ("lapphi", 0), ("iqxphi", 0), ("invqphi", 0)]
            e = ["dt phi + q^2*phi = iqxphi^2", "lapphi = -q^2*phi", "iqxphi = iqx*phi", "invqphi = 1/q*phi"]
            if shape[2] > 1:
                f += [("iqyphi", 0), ("iqzphi", 0)]; e += ["iqyphi = iqy*phi", "iqzphi = iqz*phi"]
            ic = {"phi": smooth_ic(*shape)}
            compare(f"ops{shape}", shape, 0.1, f, {}, e, ic, 3)
    if "ch2" in which:
        compare("ch2d-64 vs F", (64, 64, 1), 0.1, [("phi", 1)], CHP, CH2, {"phi": smooth_ic(64, 64, 1, 0.1, 0.01)}, 100, report_every=25)
        compare("ch2d-64 vs U(cpu-rule)", (64, 64, 1), 0.1, [("phi", 1)], CHP, CH2, {"phi": smooth_ic(64, 64, 1, 0.1, 0.01)}, 100, oracle=ORU, dev=RUN_CPU)
    if "ch3" in which:
        compare("ch3d-32 vs F", (32, 32, 32), 0.01, [("phi", 1)], CHP, CH3, {"phi": smooth_ic(32, 32, 32, 0.5, 0.05)}, 100, report_every=25)
        compare("ch3d-32 vs U(cpu-rule)", (32, 32, 32), 0.01, [("phi", 1)], CHP, CH3, {"phi": smooth_ic(32, 32, 32, 0.5, 0.05)}, 100, oracle=ORU, dev=RUN_CPU)
    if "mh" in which:
        compare("modelh-32", (32, 32, 1), 0.1, MH_F, MH_P, MH_E, {"phi": smooth_ic(32, 32, 1, 0.5, 0.025)}, 100, report_every=25)
    if "kpz" in which:
        f = [("h", 1), ("iqxh", 0), ("iqyh", 0), ("iqzh", 0)]
        e = ["dt h + 0.5*q^2*h = l*iqxh^2 + l*iqyh^2 + l*iqzh^2", "iqxh = iqx*h", "iqyh = iqy*h", "iqzh = iqz*h"]
        compare("kpz3d-det-32", (32, 32, 32), 0.01, f, dict(l=0.5), e, {"h": smooth_ic(32, 32, 32, 1.0, 0.1)}, 50)
    print(f"total {time.time() - t0:.1f}s")
