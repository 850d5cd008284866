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
typedef struct { double re, im; } cpxd;

struct cupss_shim_plan {
    int rank;
    int n[3];
    cpx *in, *out;
    int sign;
    cpx *tw[3]; /* tw[a][t] = exp(sign*2*pi*i*t/n[a]) */
    cpxd *twd[3];
};

static int g_threads = 1;
void cupss_shim_set_threads(int n) { g_threads = n > 0 ? n : 1; }
int cupss_shim_get_threads(void) { return g_threads; }

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

#define REAL float
#define CPX cpx
#define FN(name) name##_f
#include "fftw_shim_core.h"
#undef REAL
#undef CPX
#undef FN
#define REAL double
#define CPX cpxd
#define FN(name) name##_d
#include "fftw_shim_core.h"
#undef REAL
#undef CPX
#undef FN

/* 0: float arithmetic (libfftw3f-like); 1 (default): double arithmetic inside, float in/out */
static int g_double = 1;
void cupss_shim_set_double(int on) { g_double = on ? 1 : 0; }
int cupss_shim_get_double(void) { return g_double; }

/* work array of the double-inside mode (the reference's CPU path is single-threaded: one transform at a time) */
static cpxd *g_work = NULL;
static size_t g_work_len = 0;

static fftwf_plan make_plan(int rank, const int *n, fftwf_complex *in, fftwf_complex *out, int sign)
{
    fftwf_plan p = (fftwf_plan)calloc(1, sizeof(*p));
    p->rank = rank;
    for (int a = 0; a < 3; a++) p->n[a] = 1;
    for (int a = 0; a < rank; a++) p->n[a] = n[a];
    p->in = (cpx *)in;
    p->out = (cpx *)out;
    p->sign = sign < 0 ? -1 : 1;
    for (int a = 0; a < rank; a++) { p->tw[a] = make_twiddles_f(p->n[a], p->sign); p->twd[a] = make_twiddles_d(p->n[a], p->sign); }
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
    for (int a = 0; a < 3; a++) { free(p->tw[a]); free(p->twd[a]); }
    free(p);
}

void fftwf_execute(const fftwf_plan p)
{
    long total = 1;
    for (int d = 0; d < p->rank; d++) total *= p->n[d];
    if (!g_double) {
        if (p->in != p->out) memcpy(p->out, p->in, sizeof(cpx) * (size_t)total);
        for (int a = p->rank - 1; a >= 0; a--)
            transform_axis_f(p->out, p->n, p->rank, a, p->tw[a], p->sign);
        return;
    }
    /* double work array: kept between calls (a 512^3 transform would otherwise map and fault 2 GiB every time) */
    if (g_work_len < (size_t)total) {
        free(g_work);
        g_work = (cpxd *)malloc(sizeof(cpxd) * (size_t)total);
        g_work_len = g_work ? (size_t)total : 0;
    }
    cpxd *w = g_work;
    const cpx *in = p->in;
    cpx *out = p->out;
    int nth = g_threads > 0 ? g_threads : 1;
    (void)nth;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nth)
#endif
    for (long i = 0; i < total; i++) { w[i].re = in[i].re; w[i].im = in[i].im; }
    for (int a = p->rank - 1; a >= 0; a--)
        transform_axis_d(w, p->n, p->rank, a, p->twd[a], p->sign);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nth)
#endif
    for (long i = 0; i < total; i++) { out[i].re = (float)w[i].re; out[i].im = (float)w[i].im; }
}
