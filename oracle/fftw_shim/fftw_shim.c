/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product path.
 *
 * Single-precision complex DFT behind the FFTW-3 API subset declared in
 * fftw3.h (see that header for the reference call sites this replaces).
 * Arithmetic is float (as libfftw3f's is); twiddles are computed in double and
 * rounded once.  Any length works: powers of two use an iterative Stockham
 * radix-2 kernel, other lengths a recursive mixed-radix split with an O(p^2)
 * double-accumulated DFT for prime factors.
 *
 * Multi-dimensional transforms are done axis by axis on the output array
 * (copy in -> out first), gathering CUPSS_SHIM_BLOCK lines at a time into a
 * contiguous scratch so strided axes stay cache friendly.  Lines are
 * independent, so an OpenMP loop over line blocks is bit-identical to the
 * serial order.
 */
#include "fftw3.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CUPSS_SHIM_BLOCK 16

typedef struct { float re, im; } cpx;

struct cupss_shim_plan {
    int rank;
    int n[3];
    cpx *in, *out;
    int sign;
    cpx *tw[3]; /* tw[a][t] = exp(sign*2*pi*i*t/n[a]) */
};

static int g_threads = 1;
void cupss_shim_set_threads(int n) { g_threads = n > 0 ? n : 1; }
int cupss_shim_get_threads(void) { return g_threads; }

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

static cpx *make_twiddles(int n, int sign)
{
    cpx *t = (cpx *)malloc(sizeof(cpx) * (size_t)n);
    for (int i = 0; i < n; i++) {
        double a = (double)sign * 2.0 * M_PI * (double)i / (double)n;
        t[i].re = (float)cos(a);
        t[i].im = (float)sin(a);
    }
    return t;
}

/* Stockham autosort radix-2, n a power of two; x is input and output, y scratch. */
static void fft_pow2(cpx *x, cpx *y, int n, const cpx *tw)
{
    cpx *a = x, *b = y;
    int half = n / 2;
    for (int ns = 1; ns < n; ns <<= 1) {
        int tstep = n / (2 * ns);
        for (int j = 0; j < half; j++) {
            int k = j & (ns - 1);
            cpx w = tw[k * tstep];
            cpx u = a[j];
            cpx v = a[j + half];
            cpx vw;
            vw.re = v.re * w.re - v.im * w.im;
            vw.im = v.re * w.im + v.im * w.re;
            int j0 = ((j - k) << 1) + k;
            b[j0].re = u.re + vw.re;
            b[j0].im = u.im + vw.im;
            b[j0 + ns].re = u.re - vw.re;
            b[j0 + ns].im = u.im - vw.im;
        }
        cpx *t = a; a = b; b = t;
    }
    if (a != x) memcpy(x, a, sizeof(cpx) * (size_t)n);
}

static int smallest_factor(int n)
{
    for (int p = 2; p * p <= n; p++)
        if (n % p == 0) return p;
    return n;
}

/* Recursive decimation-in-time for arbitrary n.  in has stride `is`; out is
 * contiguous; tw is the table for the ROOT length nroot, tstride = nroot/n. */
static void fft_generic(const cpx *in, int is, cpx *out, int n, const cpx *tw, int nroot, int tstride, cpx *scratch)
{
    if (n == 1) { out[0] = in[0]; return; }
    int p = smallest_factor(n);
    int m = n / p;
    if (p == n) {
        for (int k = 0; k < n; k++) {
            double sr = 0.0, si = 0.0;
            for (int j = 0; j < n; j++) {
                cpx w = tw[((long)j * k % n) * tstride];
                cpx v = in[(long)j * is];
                sr += (double)v.re * w.re - (double)v.im * w.im;
                si += (double)v.re * w.im + (double)v.im * w.re;
            }
            out[k].re = (float)sr;
            out[k].im = (float)si;
        }
        return;
    }
    /* p interleaved sub-sequences of length m */
    for (int r = 0; r < p; r++)
        fft_generic(in + (long)r * is, is * p, out + (long)r * m, m, tw, nroot, tstride * p, scratch);
    /* combine: X[k + q*m] = sum_r w_n^{r(k+q*m)} Y_r[k] */
    for (int k = 0; k < m; k++) {
        for (int q = 0; q < p; q++) {
            int kk = k + q * m;
            float sr = 0.0f, si = 0.0f;
            for (int r = 0; r < p; r++) {
                cpx w = tw[((long)r * kk % n) * tstride];
                cpx v = out[(long)r * m + k];
                sr += v.re * w.re - v.im * w.im;
                si += v.re * w.im + v.im * w.re;
            }
            scratch[q].re = sr;
            scratch[q].im = si;
        }
        /* cannot overwrite out[r*m+k] before all q are formed */
        for (int q = 0; q < p; q++) { /* stash into a second scratch region */
            scratch[p + q] = scratch[q];
        }
        for (int q = 0; q < p; q++) out[(long)q * m + k] = scratch[p + q];
    }
    (void)nroot;
}

static void fft_line(cpx *x, cpx *work, int n, const cpx *tw)
{
    if (n == 1) return;
    if (is_pow2(n)) {
        fft_pow2(x, work, n, tw);
    } else {
        /* work: n outputs followed by 2*n scratch */
        fft_generic(x, 1, work, n, tw, n, 1, work + n);
        memcpy(x, work, sizeof(cpx) * (size_t)n);
    }
}

static fftwf_plan make_plan(int rank, const int *n, fftwf_complex *in, fftwf_complex *out, int sign)
{
    fftwf_plan p = (fftwf_plan)calloc(1, sizeof(*p));
    p->rank = rank;
    for (int a = 0; a < 3; a++) p->n[a] = 1;
    for (int a = 0; a < rank; a++) p->n[a] = n[a];
    p->in = (cpx *)in;
    p->out = (cpx *)out;
    p->sign = sign < 0 ? -1 : 1;
    for (int a = 0; a < rank; a++) p->tw[a] = make_twiddles(p->n[a], p->sign);
    return p;
}

fftwf_plan fftwf_plan_dft_1d(int n0, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{
    (void)flags;
    int n[1] = { n0 };
    return make_plan(1, n, in, out, sign);
}

fftwf_plan fftwf_plan_dft_2d(int n0, int n1, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{
    (void)flags;
    int n[2] = { n0, n1 };
    return make_plan(2, n, in, out, sign);
}

fftwf_plan fftwf_plan_dft_3d(int n0, int n1, int n2, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{
    (void)flags;
    int n[3] = { n0, n1, n2 };
    return make_plan(3, n, in, out, sign);
}

void fftwf_destroy_plan(fftwf_plan p)
{
    if (!p) return;
    for (int a = 0; a < 3; a++) free(p->tw[a]);
    free(p);
}

/* transform every line along axis `a` of the row-major array `data` in place */
static void transform_axis(cpx *data, const int *n, int rank, int a, const cpx *tw)
{
    long total = 1;
    for (int d = 0; d < rank; d++) total *= n[d];
    int len = n[a];
    if (len == 1) return;
    long stride = 1;
    for (int d = a + 1; d < rank; d++) stride *= n[d];
    long nlines = total / len;
    long nblocks = (nlines + CUPSS_SHIM_BLOCK - 1) / CUPSS_SHIM_BLOCK;
    int nth = g_threads;
    if (nblocks < nth) nth = (int)nblocks;
    if (nth < 1) nth = 1;

#ifdef _OPENMP
#pragma omp parallel num_threads(nth)
#endif
    {
        cpx *buf = (cpx *)malloc(sizeof(cpx) * (size_t)len * CUPSS_SHIM_BLOCK);
        cpx *work = (cpx *)malloc(sizeof(cpx) * (size_t)len * 4 + 64);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long b = 0; b < nblocks; b++) {
            long l0 = b * CUPSS_SHIM_BLOCK;
            int nl = (int)((nlines - l0) < CUPSS_SHIM_BLOCK ? (nlines - l0) : CUPSS_SHIM_BLOCK);
            if (stride == 1) {
                for (int l = 0; l < nl; l++)
                    fft_line(data + (l0 + l) * len, work, len, tw);
                continue;
            }
            /* line index -> (outer, inner): base = outer*len*stride + inner */
            for (int l = 0; l < nl; l++) {
                long li = l0 + l;
                long base = (li / stride) * (long)len * stride + (li % stride);
                cpx *dst = buf + (long)l * len;
                for (int j = 0; j < len; j++) dst[j] = data[base + (long)j * stride];
            }
            for (int l = 0; l < nl; l++) fft_line(buf + (long)l * len, work, len, tw);
            for (int l = 0; l < nl; l++) {
                long li = l0 + l;
                long base = (li / stride) * (long)len * stride + (li % stride);
                const cpx *src = buf + (long)l * len;
                for (int j = 0; j < len; j++) data[base + (long)j * stride] = src[j];
            }
        }
        free(buf);
        free(work);
    }
}

void fftwf_execute(const fftwf_plan p)
{
    long total = 1;
    for (int d = 0; d < p->rank; d++) total *= p->n[d];
    if (p->in != p->out) memcpy(p->out, p->in, sizeof(cpx) * (size_t)total);
    for (int a = p->rank - 1; a >= 0; a--)
        transform_axis(p->out, p->n, p->rank, a, p->tw[a]);
}
