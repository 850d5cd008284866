/*
 * ORACLE / TEST INFRASTRUCTURE ONLY.  Core of the FFTW-API shim, included twice by fftw_shim.c:
 * REAL = float  (arithmetic like libfftw3f) and REAL = double (float in/out, double inside: one rounding
 * per transform -- removes the shim's own coherent round-off so that parity runs measure the product,
 * not the stand-in FFT; see DESIGN.md "oracle").  Needs REAL, CPX and FN(name) defined.
 */
static CPX *FN(make_twiddles)(int n, int sign)
{
    CPX *t = (CPX *)malloc(sizeof(CPX) * (size_t)n);
    for (int i = 0; i < n; i++) {
        double a = (double)sign * 2.0 * M_PI * (double)i / (double)n;
        t[i].re = (REAL)cos(a);
        t[i].im = (REAL)sin(a);
    }
    return t;
}

/* Stockham autosort radix-2, n a power of two; x is input and output, y scratch. */
static void FN(fft_pow2)(CPX *x, CPX *y, int n, const CPX *tw)
{
    CPX *a = x, *b = y;
    int half = n / 2;
    for (int ns = 1; ns < n; ns <<= 1) {
        int tstep = n / (2 * ns);
        for (int j = 0; j < half; j++) {
            int k = j & (ns - 1);
            CPX w = tw[k * tstep];
            CPX u = a[j];
            CPX v = a[j + half];
            CPX vw;
            vw.re = v.re * w.re - v.im * w.im;
            vw.im = v.re * w.im + v.im * w.re;
            int j0 = ((j - k) << 1) + k;
            b[j0].re = u.re + vw.re;
            b[j0].im = u.im + vw.im;
            b[j0 + ns].re = u.re - vw.re;
            b[j0 + ns].im = u.im - vw.im;
        }
        CPX *t = a; a = b; b = t;
    }
    if (a != x) memcpy(x, a, sizeof(CPX) * (size_t)n);
}

#ifndef CUPSS_SHIM_SMALLEST_FACTOR
#define CUPSS_SHIM_SMALLEST_FACTOR
static int smallest_factor(int n)
{
    for (int p = 2; p * p <= n; p++)
        if (n % p == 0) return p;
    return n;
}
#endif

/* Recursive decimation-in-time for arbitrary n.  in has stride `is`; out is
 * contiguous; tw is the table for the ROOT length nroot, tstride = nroot/n. */
static void FN(fft_generic)(const CPX *in, int is, CPX *out, int n, const CPX *tw, int nroot, int tstride, CPX *scratch)
{
    if (n == 1) { out[0] = in[0]; return; }
    int p = smallest_factor(n);
    int m = n / p;
    if (p == n) {
        for (int k = 0; k < n; k++) {
            double sr = 0.0, si = 0.0;
            for (int j = 0; j < n; j++) {
                CPX w = tw[((long)j * k % n) * tstride];
                CPX v = in[(long)j * is];
                sr += (double)v.re * w.re - (double)v.im * w.im;
                si += (double)v.re * w.im + (double)v.im * w.re;
            }
            out[k].re = (REAL)sr;
            out[k].im = (REAL)si;
        }
        return;
    }
    /* p interleaved sub-sequences of length m */
    for (int r = 0; r < p; r++)
        FN(fft_generic)(in + (long)r * is, is * p, out + (long)r * m, m, tw, nroot, tstride * p, scratch);
    /* combine: X[k + q*m] = sum_r w_n^{r(k+q*m)} Y_r[k] */
    for (int k = 0; k < m; k++) {
        for (int q = 0; q < p; q++) {
            int kk = k + q * m;
            REAL sr = 0, si = 0;
            for (int r = 0; r < p; r++) {
                CPX w = tw[((long)r * kk % n) * tstride];
                CPX v = out[(long)r * m + k];
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

static void FN(fft_line)(CPX *x, CPX *work, int n, const CPX *tw)
{
    if (n == 1) return;
    if (is_pow2(n)) {
        FN(fft_pow2)(x, work, n, tw);
    } else {
        /* work: n outputs followed by 2*n scratch */
        FN(fft_generic)(x, 1, work, n, tw, n, 1, work + n);
        memcpy(x, work, sizeof(CPX) * (size_t)n);
    }
}

/* transform every line along axis `a` of the row-major array `data` in place */
static void FN(transform_axis)(CPX *data, const int *n, int rank, int a, const CPX *tw)
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
        CPX *buf = (CPX *)malloc(sizeof(CPX) * (size_t)len * CUPSS_SHIM_BLOCK);
        CPX *work = (CPX *)malloc(sizeof(CPX) * (size_t)len * 4 + 64);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long b = 0; b < nblocks; b++) {
            long l0 = b * CUPSS_SHIM_BLOCK;
            int nl = (int)((nlines - l0) < CUPSS_SHIM_BLOCK ? (nlines - l0) : CUPSS_SHIM_BLOCK);
            if (stride == 1) {
                for (int l = 0; l < nl; l++)
                    FN(fft_line)(data + (l0 + l) * len, work, len, tw);
                continue;
            }
            /* line index -> (outer, inner): base = outer*len*stride + inner */
            for (int l = 0; l < nl; l++) {
                long li = l0 + l;
                long base = (li / stride) * (long)len * stride + (li % stride);
                CPX *dst = buf + (long)l * len;
                for (int j = 0; j < len; j++) dst[j] = data[base + (long)j * stride];
            }
            for (int l = 0; l < nl; l++) FN(fft_line)(buf + (long)l * len, work, len, tw);
            for (int l = 0; l < nl; l++) {
                long li = l0 + l;
                long base = (li / stride) * (long)len * stride + (li % stride);
                const CPX *src = buf + (long)l * len;
                for (int j = 0; j < len; j++) data[base + (long)j * stride] = src[j];
            }
        }
        free(buf);
        free(work);
    }
}

