// fft_core.cuh -- register/shared-memory FFT building blocks for the sm_100a step kernels.
//
// Replaces the cuFFT C2C executions on the reference's hot path
// (/root/reference/src/field.cpp:247-273 toReal/toComp, src/term.cpp:48-64 term::toComp).
//
// Every transform in this engine is a Stockham autosort FFT over a power-of-two line of L
// points, executed by T = L/E threads that each own E points in registers ("canonical"
// ownership: thread t holds points t + T*e, e = 0..E-1).  A pass of radix R does E/R
// register butterflies per thread; between passes the line goes once through shared memory.
// With E = 32 a 512-point line needs two passes (radix 32 then 16) and therefore ONE
// shared-memory exchange -- the shared-memory crossbar (128 B/clk/SM), not HBM, is the
// scarce resource for these kernels on B200, see DESIGN.md.
//
// Everything here is __host__ __device__ so tests/host_fft_check.cu can run the very same
// index arithmetic and butterflies on the CPU (there is no GPU in the build container).
#pragma once
#include <cuda_runtime.h>
#include <utility>

#define CUPSS_HD __host__ __device__ __forceinline__

namespace cupss {

// ---------------------------------------------------------------- complex helpers
CUPSS_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
CUPSS_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
CUPSS_HD float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}
CUPSS_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
CUPSS_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// ---------------------------------------------------------------- constexpr trig (compile-time twiddles)
constexpr double kPi = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double cx_sin(double x) {   // Taylor, |x| <= pi, evaluated by the compiler only
    double term = x, sum = x, x2 = x * x;
    for (int n = 1; n < 20; ++n) { term *= -x2 / ((2.0 * n) * (2.0 * n + 1.0)); sum += term; }
    return sum;
}
__host__ __device__ constexpr double cx_cos(double x) {
    double term = 1.0, sum = 1.0, x2 = x * x;
    for (int n = 1; n < 20; ++n) { term *= -x2 / ((2.0 * n - 1.0) * (2.0 * n)); sum += term; }
    return sum;
}

// v * exp(DIR * 2*pi*i * K / N), K and N compile-time.  DIR = -1 forward, +1 inverse.
template <int K, int N, int DIR>
CUPSS_HD float2 twiddle_mul(float2 v) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        return v;
    } else if constexpr (2 * k == N) {
        return make_float2(-v.x, -v.y);
    } else if constexpr (4 * k == N) {          // * (DIR * i)
        return DIR > 0 ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
    } else if constexpr (4 * k == 3 * N) {      // * (-DIR * i)
        return DIR > 0 ? make_float2(v.y, -v.x) : make_float2(-v.y, v.x);
    } else if constexpr (8 * k == N) {          // (1 + DIR*i)/sqrt2
        constexpr float h = 0.70710678118654752440f;
        return DIR > 0 ? make_float2((v.x - v.y) * h, (v.x + v.y) * h)
                       : make_float2((v.x + v.y) * h, (v.y - v.x) * h);
    } else if constexpr (8 * k == 3 * N) {      // (-1 + DIR*i)/sqrt2
        constexpr float h = 0.70710678118654752440f;
        return DIR > 0 ? make_float2((-v.x - v.y) * h, (v.x - v.y) * h)
                       : make_float2((v.y - v.x) * h, (-v.x - v.y) * h);
    } else {
        constexpr float c = (float)cx_cos(2.0 * kPi * k / N);
        constexpr float s = (float)(DIR * cx_sin(2.0 * kPi * k / N));
        return make_float2(fmaf(-v.y, s, v.x * c), fmaf(v.x, s, v.y * c));
    }
}

// ---------------------------------------------------------------- in-register DFT of N points, natural order in/out
template <int N, int DIR>
struct Dft {
    template <int K>
    static CUPSS_HD void comb1(float2 (&x)[N], const float2 (&e)[N / 2], const float2 (&o)[N / 2]) {
        float2 t = twiddle_mul<K, N, DIR>(o[K]);
        x[K] = cadd(e[K], t);
        x[K + N / 2] = csub(e[K], t);
    }
    template <int... K>
    static CUPSS_HD void combine(float2 (&x)[N], const float2 (&e)[N / 2], const float2 (&o)[N / 2],
                                 std::integer_sequence<int, K...>) {
        (comb1<K>(x, e, o), ...);
    }
    static CUPSS_HD void run(float2 (&x)[N]) {
        float2 e[N / 2], o[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; ++i) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; }
        Dft<N / 2, DIR>::run(e);
        Dft<N / 2, DIR>::run(o);
        combine(x, e, o, std::make_integer_sequence<int, N / 2>{});
    }
};
template <int DIR>
struct Dft<1, DIR> { static CUPSS_HD void run(float2 (&)[1]) {} };
template <int DIR>
struct Dft<2, DIR> {
    static CUPSS_HD void run(float2 (&x)[2]) {
        float2 a = x[0], b = x[1];
        x[0] = cadd(a, b); x[1] = csub(a, b);
    }
};
template <int DIR>
struct Dft<4, DIR> {
    static CUPSS_HD void run(float2 (&x)[4]) {
        float2 t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
        float2 t2 = cadd(x[1], x[3]), t3 = twiddle_mul<1, 4, DIR>(csub(x[1], x[3]));
        x[0] = cadd(t0, t2); x[1] = cadd(t1, t3); x[2] = csub(t0, t2); x[3] = csub(t1, t3);
    }
};

// ---------------------------------------------------------------- one Stockham pass on canonical registers
// Line of L points, E per thread (T = L/E threads), radix R, NS = product of earlier radices.
// Register e = m + r*(E/R) is input r of butterfly j = t + T*m (point j + r*L/R); after the
// call it holds output r of that butterfly, which belongs at point stockham_out_index().
// `tw` is the forward table exp(-2*pi*i*k/L), k = 0..L-1 (conjugated on the fly for DIR=+1).
template <int L, int E, int R, int NS, int DIR>
CUPSS_HD void stockham_pass(float2 (&v)[E], int t, const float2* __restrict__ tw) {
    constexpr int T = L / E, M = E / R;
    static_assert(E % R == 0 && L % E == 0, "bad FFT factorisation");
#pragma unroll
    for (int m = 0; m < M; ++m) {
        float2 a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = v[m + r * M];
        if constexpr (NS > 1) {
            const int k = (t + T * m) & (NS - 1);
            constexpr int step = L / (NS * R);
#pragma unroll
            for (int r = 1; r < R; ++r) {
#ifdef __CUDA_ARCH__
                float2 w = __ldg(tw + r * k * step);
#else
                float2 w = tw[r * k * step];
#endif
                if (DIR > 0) w.y = -w.y;
                a[r] = cmul(a[r], w);
            }
        }
        Dft<R, DIR>::run(a);
#pragma unroll
        for (int r = 0; r < R; ++r) v[m + r * M] = a[r];
    }
}

// Point index where register e (of thread t) must be stored after a radix-R pass with stride NS.
template <int L, int E, int R, int NS>
CUPSS_HD int stockham_out_index(int t, int e) {
    constexpr int T = L / E, M = E / R;
    const int m = e % M, r = e / M;
    const int j = t + T * m;
    return (j / NS) * NS * R + (j & (NS - 1)) + r * NS;
}

// ---------------------------------------------------------------- plan selection: L -> (E, R0, R1, R2)
template <int L> struct FftPlan;
#define CUPSS_FFTPLAN(L_, E_, R0_, R1_, R2_) \
    template <> struct FftPlan<L_> { static constexpr int E = E_, R0 = R0_, R1 = R1_, R2 = R2_, T = L_ / E_; };
CUPSS_FFTPLAN(1, 1, 1, 1, 1)
CUPSS_FFTPLAN(2, 2, 2, 1, 1)
CUPSS_FFTPLAN(4, 4, 4, 1, 1)
CUPSS_FFTPLAN(8, 8, 8, 1, 1)
CUPSS_FFTPLAN(16, 16, 16, 1, 1)
CUPSS_FFTPLAN(32, 32, 32, 1, 1)
CUPSS_FFTPLAN(64, 16, 8, 8, 1)
CUPSS_FFTPLAN(128, 16, 16, 8, 1)
CUPSS_FFTPLAN(256, 16, 16, 16, 1)
CUPSS_FFTPLAN(512, 32, 32, 16, 1)
CUPSS_FFTPLAN(1024, 32, 32, 32, 1)
CUPSS_FFTPLAN(2048, 32, 16, 16, 8)
CUPSS_FFTPLAN(4096, 32, 16, 16, 16)
CUPSS_FFTPLAN(8192, 32, 32, 16, 16)
#undef CUPSS_FFTPLAN

// Full line FFT on canonical registers.  `Ex` provides the shared-memory exchange:
//   ex.st(point, value), ex.ld(point), ex.sync().
// On entry the exchange buffer must be free; on exit other threads may still be reading it,
// so callers sync before re-using it.  Result is canonical (thread t holds points t + T*e).
template <int L, int DIR, class Ex>
CUPSS_HD void fft_line(float2 (&v)[FftPlan<L>::E], int t, const float2* __restrict__ tw, Ex& ex) {
    using P = FftPlan<L>;
    constexpr int E = P::E, T = P::T, R0 = P::R0, R1 = P::R1, R2 = P::R2;
    stockham_pass<L, E, R0, 1, DIR>(v, t, tw);
    if constexpr (R1 > 1) {
#pragma unroll
        for (int e = 0; e < E; ++e) ex.st(stockham_out_index<L, E, R0, 1>(t, e), v[e]);
        ex.sync();
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = ex.ld(t + T * e);
        stockham_pass<L, E, R1, R0, DIR>(v, t, tw);
        if constexpr (R2 > 1) {
            ex.sync();
#pragma unroll
            for (int e = 0; e < E; ++e) ex.st(stockham_out_index<L, E, R1, R0>(t, e), v[e]);
            ex.sync();
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = ex.ld(t + T * e);
            stockham_pass<L, E, R2, R0 * R1, DIR>(v, t, tw);
        }
    }
}

}  // namespace cupss
