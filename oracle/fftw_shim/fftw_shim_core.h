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

/* Stockham autosort, n a power of two: radix-4 passes (one radix-2 pass first when log2 n is odd); x is input and
 * output, y scratch.  Pass with sub-transform length ns: out[(j-k)*R + k + q*ns] = sum_r w_{R ns}^{k r} ... as in the
 * radix-2 version, with two radix-2 levels merged into one sweep over the data. */
static void FN(fft_pow2)(CPX *x, CPX *y, int n, const CPX *tw, int sign)
{
    CPX *a = x, *b = y;
    int ns = 1;
    int lg = 0;
    while ((1 << lg) < n) lg++;
    if (lg & 1) {   /* radix-2 pass, ns = 1: twiddles are all 1 */
        int half = n / 2;
        for (int j = 0; j < half; j++) {
            CPX u = a[j], v = a[j + half];
            b[2 * j].re = u.re + v.re;     b[2 * j].im = u.im + v.im;
            b[2 * j + 1].re = u.re - v.re; b[2 * j + 1].im = u.im - v.im;
        }
        CPX *t = a; a = b; b = t;
        ns = 2;
    }
    int quarter = n / 4;
    for (; ns < n; ns <<= 2) {
        int tstep = n / (4 * ns);
        for (int j = 0; j < quarter; j++) {
            int k = j & (ns - 1);
            CPX w1 = tw[k * tstep], w2 = tw[2 * k * tstep], w3 = tw[3 * k * tstep];
            CPX v0 = a[j], v1 = a[j + quarter], v2 = a[j + 2 * quarter], v3 = a[j + 3 * quarter];
            CPX t1, t2, t3;
            t1.re = v1.re * w1.re - v1.im * w1.im; t1.im = v1.re * w1.im + v1.im * w1.re;
            t2.re = v2.re * w2.re - v2.im * w2.im; t2.im = v2.re * w2.im + v2.im * w2.re;
            t3.re = v3.re * w3.re - v3.im * w3.im; t3.im = v3.re * w3.im + v3.im * w3.re;
            CPX s02, d02, s13, d13;
            s02.re = v0.re + t2.re; s02.im = v0.im + t2.im;
            d02.re = v0.re - t2.re; d02.im = v0.im - t2.im;
            s13.re = t1.re + t3.re; s13.im = t1.im + t3.im;
            d13.re = t1.re - t3.re; d13.im = t1.im - t3.im;
            /* multiply d13 by sign*i: exp(sign*2*pi*i/4) */
            CPX r13;
            if (sign > 0) { r13.re = -d13.im; r13.im = d13.re; } else { r13.re = d13.im; r13.im = -d13.re; }
            int j0 = ((j - k) << 2) + k;
            b[j0].re = s02.re + s13.re;          b[j0].im = s02.im + s13.im;
            b[j0 + ns].re = d02.re + r13.re;     b[j0 + ns].im = d02.im + r13.im;
            b[j0 + 2 * ns].re = s02.re - s13.re; b[j0 + 2 * ns].im = s02.im - s13.im;
            b[j0 + 3 * ns].re = d02.re - r13.re; b[j0 + 3 * ns].im = d02.im - r13.im;
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

static void FN(fft_line)(CPX *x, CPX *work, int n, const CPX *tw, int sign)
{
    if (n == 1) return;
    if (is_pow2(n)) {
        FN(fft_pow2)(x, work, n, tw, sign);
    } else {
        /* work: n outputs followed by 2*n scratch */
        FN(fft_generic)(x, 1, work, n, tw, n, 1, work + n);
        memcpy(x, work, sizeof(CPX) * (size_t)n);
    }
}

/* transform every line along axis `a` of the row-major array `data` in place */
static void FN(transform_axis)(CPX *data, const int *n, int rank, int a, const CPX *tw, int sign)
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
                    FN(fft_line)(data + (l0 + l) * len, work, len, tw, sign);
                continue;
            }
            /* line index -> (outer, inner): base = outer*len*stride + inner */
            /* the nl lines of a block are neighbours in memory when the block does not wrap around `stride`: walk the
             * strided index outermost so that every cache line fetched is used completely (same values, other order) */
            const int together = (l0 % stride) + nl <= stride;
            const long base0 = (l0 / stride) * (long)len * stride + (l0 % stride);
            if (together) {
                for (int j = 0; j < len; j++) {
                    const CPX *src = data + base0 + (long)j * stride;
                    for (int l = 0; l < nl; l++) buf[(long)l * len + j] = src[l];
                }
            } else
            for (int l = 0; l < nl; l++) {
                long li = l0 + l;
                long base = (li / stride) * (long)len * stride + (li % stride);
                CPX *dst = buf + (long)l * len;
                for (int j = 0; j < len; j++) dst[j] = data[base + (long)j * stride];
            }
            for (int l = 0; l < nl; l++) FN(fft_line)(buf + (long)l * len, work, len, tw, sign);
            if (together) {
                for (int j = 0; j < len; j++) {
                    CPX *dst = data + base0 + (long)j * stride;
                    for (int l = 0; l < nl; l++) dst[l] = buf[(long)l * len + j];
                }
            } else
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

