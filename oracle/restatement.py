"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

numpy restatement of the reference's advanceTime loop (full complex spectra, like the reference), in float64 or
float32.  It is an independent cross-check of the compiled reference under oracle/_ref and the yard-stick that tells
round-off growth apart from semantic differences.  Pinned against the reference's own golden vectors
(tests/base_truths/*) in tests/test_oracle.py.

Followed reference code (paths under /root/reference):
  evolver::advanceTime         src/evolver.cpp:199-226      four sweeps: constraint terms, constraint RHS, dynamic terms, dynamic RHS
  field::updateTerms           src/field.cpp:15-23
  term::update                 src/term.cpp:12-25           single field: copy spectrum; else real-space product + FFT; then prefactor
  term::computeProduct         src/term.cpp:66-102          product of real_dealiased[].x
  term::precomputePrefactors   src/term_init.cpp:135-200    sum_p sgn*pre*q2^n*qx^a*qy^b*qz^c*invq^m ; 1/|q| := 0 at index 0
  term::applyPres_vector       src/term.cpp:225-255         multiply; rotate by i for an odd power of i
  field::setRHS                src/field.cpp:25-92          update, dealias, inverse FFT, normalise (imag := 0), forward FFT
  field::stepEuler             src/field.cpp:171-193        comp += dt*sum(terms); comp /= precomp_implicit
  field::setNotDynamic         src/field.cpp:94-148 with the `imp++` fix == setNotDynamic_k, src/field_kernels.cu:130-197
  field::precalculateImplicit  src/field_init.cpp:237-284   1 - dt*sum(pre*q2^n*invq^m), invq rule (i>0||j>0)
  field::dealias               src/field.cpp:203-232 (CPU rule, `nj` typo) / dealias_k src/field_kernels.cu:229-256 (GPU rule)
  term::prepareDevice          src/term_init.cpp:118-126    needsaliasing / aliasing_order
"""
from __future__ import annotations

import numpy as np

PI32 = np.float32(3.1415926535)   # inc/cupss/defines.h:54


class Pres:
    def __init__(self, pre=0.0, q2n=0, iqx=0, iqy=0, iqz=0, invq=0):
        self.pre, self.q2n, self.iqx, self.iqy, self.iqz, self.invq = float(pre), q2n, iqx, iqy, iqz, invq


class System:
    """fields: list of (name, dynamic); implicit[name] = [Pres]; terms[name] = [([Pres], [field names])]."""

    def __init__(self, shape, d, dt, dtype=np.float64, dealias_rule="gpu"):
        self.sx, self.sy, self.sz = shape
        self.dx, self.dy, self.dz = d
        self.dt = dt
        self.dtype = np.dtype(dtype)
        self.ctype = np.complex128 if self.dtype == np.float64 else np.complex64
        self.rule = dealias_rule
        self.names, self.dynamic = [], {}
        self.implicit, self.terms = {}, {}
        self.real, self.comp, self.real_dealiased = {}, {}, {}
        self.alias, self.order = {}, {}
        f32 = np.float32
        # wavenumbers built in float32 exactly like the reference, then promoted
        def axis(n, dd):
            step = f32(2.0) * PI32 / (f32(dd) * f32(n))
            i = np.arange(n)
            return (np.where(i < (n + 1) // 2, i, i - n).astype(np.float32) * step).astype(self.dtype)
        self.qx = axis(self.sx, self.dx)[None, None, :]
        self.qy = axis(self.sy, self.dy)[None, :, None] if self.sy > 1 else np.zeros((1, 1, 1), self.dtype)
        self.qz = axis(self.sz, self.dz)[:, None, None] if self.sz > 1 else np.zeros((1, 1, 1), self.dtype)
        self.q2 = self.qx ** 2 + self.qy ** 2 + self.qz ** 2 + np.zeros((self.sz, self.sy, self.sx), self.dtype)
        with np.errstate(divide="ignore"):
            self.invq = np.where(self.q2 > 0, 1.0 / np.sqrt(self.q2), 0.0).astype(self.dtype)
        self.invq.flat[0] = 0.0
        ii = np.arange(self.sx)[None, None, :] + np.zeros((self.sz, self.sy, 1), int)
        jj = np.arange(self.sy)[None, :, None] + np.zeros((self.sz, 1, self.sx), int)
        self.invq_legacy = np.where((ii > 0) | (jj > 0), self.invq, 0.0)   # src/field_init.cpp:258

    def add_field(self, name, dynamic):
        self.names.append(name)
        self.dynamic[name] = bool(dynamic)
        self.implicit[name], self.terms[name] = [], []
        shape = (self.sz, self.sy, self.sx)
        self.real[name] = np.zeros(shape, self.dtype)
        self.comp[name] = np.zeros(shape, self.ctype)
        self.real_dealiased[name] = np.zeros(shape, self.dtype)
        self.alias[name], self.order[name] = False, 1

    def prepare(self):
        for n in self.names:
            self.comp[n] = np.fft.fftn(self.real[n]).astype(self.ctype)
        for n in self.names:
            for _, prod in self.terms[n]:
                if len(prod) != 1:
                    for g in prod:
                        self.alias[g] = True
                        self.order[g] = max(self.order[g], len(prod))
        self._pref = {n: [self._prefactor(p) for p, _ in self.terms[n]] for n in self.names}
        self._imp = {n: self._implicit(n) for n in self.names}

    def _prefactor(self, pres):
        mul_i = (pres[0].iqx + pres[0].iqy + pres[0].iqz) % 2
        tot = np.zeros((self.sz, self.sy, self.sx), self.dtype)
        for p in pres:
            units = p.iqx + p.iqy + p.iqz
            negate = -2 * (((units - mul_i) // 2) % 2) + 1
            v = np.full_like(tot, p.pre * negate)
            if p.q2n > 0: v = v * self.q2 ** p.q2n
            if p.iqx > 0: v = v * self.qx ** p.iqx
            if p.iqy > 0: v = v * self.qy ** p.iqy
            if p.iqz > 0: v = v * self.qz ** p.iqz
            if p.invq > 0: v = v * self.invq ** p.invq
            tot = tot + v
        return tot, mul_i

    def _implicit(self, n):
        dyn = self.dynamic[n]
        f = np.full((self.sz, self.sy, self.sx), 1.0 if dyn else 0.0, self.dtype)
        for p in self.implicit[n]:
            v = np.full_like(f, p.pre)
            if p.q2n != 0: v = v * self.q2 ** p.q2n
            if p.invq != 0: v = v * (self.invq_legacy if dyn else self.invq) ** p.invq
            f = f - self.dt * v if dyn else f + v
        return f

    def _mask(self, order):
        def n_abs(n):
            i = np.arange(n)
            return np.abs(np.where(i > n // 2, i - n, i))
        nx, ny, nz = n_abs(self.sx)[None, None, :], n_abs(self.sy)[None, :, None], n_abs(self.sz)[:, None, None]
        cx, cy, cz = self.sx // (order + 1), self.sy // (order + 1), self.sz // (order + 1)
        if self.rule == "gpu":
            return ~((nx > cx) | (ny > cy) | (nz > cz))
        return ~((nx > cx) | (ny > cy) | (ny > cz)) & np.ones((self.sz, 1, 1), bool)   # src/field.cpp:220

    def _update_terms(self, n):
        out = []
        for (pres, prod), (pf, mul_i) in zip(self.terms[n], self._pref[n]):
            if len(prod) == 1:
                t = self.comp[prod[0]].copy()
            else:
                r = np.ones((self.sz, self.sy, self.sx), self.dtype)
                for g in prod:
                    r = r * self.real_dealiased[g]
                t = np.fft.fftn(r).astype(self.ctype)
            t = t * pf
            if mul_i:
                t = 1j * t
            out.append(t.astype(self.ctype))
        return out

    def _set_rhs(self, n, terms):
        c = self.comp[n]
        if self.dynamic[n]:
            for t in terms:
                c = c + self.dt * t
            if self.implicit[n]:
                c = c / self._imp[n]
        else:
            if terms:
                c = terms[0]
                for t in terms[1:]:
                    c = c + t
            if self.implicit[n]:
                f = self._imp[n].copy()
                c0 = c.flat[0]
                f.flat[0] = 1.0
                c = c / f
                c.flat[0] = c0
        c = c.astype(self.ctype)
        if self.alias[n]:
            cd = np.where(self._mask(self.order[n]), c, 0)
            self.real_dealiased[n] = np.fft.ifftn(cd).real.astype(self.dtype)
        self.real[n] = np.fft.ifftn(c).real.astype(self.dtype)
        self.comp[n] = np.fft.fftn(self.real[n]).astype(self.ctype)

    def step(self, nsteps=1):
        for _ in range(nsteps):
            for dyn in (False, True):
                sel = [n for n in self.names if self.dynamic[n] == dyn]
                terms = {n: self._update_terms(n) for n in sel}
                for n in sel:
                    self._set_rhs(n, terms[n])


def from_plan_dump(dump: str, shape, d, dt, dtype=np.float64, dealias_rule="gpu") -> System:
    """Build a System from the text of tools/cupss_capi.cpp::cupss_capi_dump_plan (so the SAME parsed system drives
    the compiled reference, the product and this restatement)."""
    s = System(shape, d, dt, dtype, dealias_rule)
    cur = None
    def parse_pres(tok):
        v = tok.strip("{}").split(",")
        return Pres(float(v[0]), int(v[1]), int(v[2]), int(v[3]), int(v[4]), int(v[5]))
    for line in dump.splitlines():
        w = line.split()
        if not w:
            continue
        if w[0] == "field":
            cur = w[1]
            s.add_field(cur, w[2] == "dynamic=1")
        elif w[0] == "implicit":
            s.implicit[cur] = [parse_pres(t) for t in w[1:]]
        elif w[0] == "term":
            i = w.index("(")
            s.terms[cur].append(([parse_pres(t) for t in w[1:i]], w[i + 1:-1]))
    return s
