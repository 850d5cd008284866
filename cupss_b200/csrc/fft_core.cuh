// fft_core.cuh -- FFT building blocks of the sm_100a step kernels.
//
// Replaces the cuFFT C2C executions on the reference's hot path
// (/root/reference/src/field.cpp:247-273 toReal/toComp, src/term.cpp:48-64 term::toComp).
//
// Every transform is an IN-PLACE mixed-radix FFT over a power-of-two line of L points that lives in
// shared memory between passes ("levels").  With radices R1..Rn (L = R1*...*Rn) level l works on
// independent blocks of N_l = L/(R1..R_{l-1}) points; one "virtual thread" v owns the R_l points
//      row(q) = blk*N_l + j + M_l*q,   M_l = N_l/R_l,  blk = v / M_l,  j = v % M_l,  q = 0..R_l-1
// reads them, does one radix-R_l butterfly in registers and writes the SAME positions back, so there
// is no ping-pong buffer and no hazard inside a level (one __syncthreads between levels).
//   decimation in frequency (DIF), levels 1..n : butterfly, then twiddle w_{N_l}^{j q}
//              natural order in -> digit-reversed order out
//   decimation in time (DIT),      levels n..1 : twiddle w_{N_l}^{j q}, then butterfly
//              digit-reversed order in -> natural order out
// for either sign of the exponent.  The strided-axis kernels run forward = DIF, inverse = DIT; the x pass
// runs inverse = DIF, forward = DIT (its real-space stage sits in the middle and is pointwise).
// Position p = q1*(L/R1) + q2*(L/(R1 R2)) + ... + qn holds frequency k = q1 + R1 q2 + R1 R2 q3 + ...
// The permutation is free where it matters: the strided-axis kernels move whole 128-byte row segments
// between global and shared memory, so a permuted ROW order costs nothing, and the k-space stage /
// the real-space products are pointwise and do not care about the order at all.
// A real thread loops over several virtual threads per level (rolled loop): few registers, a small
// instruction footprint (the unrolled register-resident design this replaces overflowed the
// instruction cache: 24 % "no instruction" stalls in profiles/r01a) and high occupancy.
// Complex add/sub use the packed FP32 instructions of sm_100 (FADD2 / FFMA2): one issue slot per
// complex operation -- these kernels are bound by issue slots and HBM, not by FP32 throughput.
//
// Everything is __host__ __device__: tests/host/fft_core_check.cu runs the very same level arithmetic
// on the CPU (there is no GPU in the build container).
#pragma once
#ifdef __CUDACC_RTC__
// run-time compilation (plan-specialised k stage, engine.cu): no host headers
#include <cuda/std/utility>
namespace cupss_std = cuda::std;
#else
#include <cuda_runtime.h>
#include <utility>
namespace cupss_std = std;
#endif

#define CUPSS_HD __host__ __device__ __forceinline__

namespace cupss {

// ---------------------------------------------------------------- complex helpers
CUPSS_HD float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}
CUPSS_HD float2 cmul_conj(float2 a, float2 b) {   // a * conj(b)
    return make_float2(fmaf(a.y, b.y, a.x * b.x), fmaf(a.y, b.x, -(a.x * b.y)));
}
CUPSS_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
CUPSS_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// packed FP32x2 (one instruction per complex add / sub / scaled add on sm_100)
CUPSS_HD float2 cadd(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
CUPSS_HD float2 csub(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);   // exact product, one rounding: == a - b
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
CUPSS_HD float2 cfma_s(float2 u, float s, float2 c) {   // u*s + c, componentwise
#ifdef __CUDA_ARCH__
    return __ffma2_rn(u, make_float2(s, s), c);
#else
    return make_float2(fmaf(u.x, s, c.x), fmaf(u.y, s, c.y));
#endif
}

CUPSS_HD float2 cmul2(float2 a, float2 b) {   // componentwise product (FMUL2), NOT the complex product
#ifdef __CUDA_ARCH__
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}

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

// ---------------------------------------------------------------- in-register DFT of N points, natural order in/out
// Recursive even/odd split; the combine step x[K] = e[K] + w^K o[K], x[K+N/2] = e[K] - w^K o[K]
// (w = exp(DIR*2*pi*i/N)) folds the trivial twiddles into the adds.
template <int N, int DIR>
struct Dft {
    template <int K>
    static CUPSS_HD void comb1(float2 (&x)[N], const float2 (&e)[N / 2], const float2 (&o)[N / 2]) {
        constexpr float h = 0.70710678118654752440f;
        const float2 E = e[K], O = o[K];
        if constexpr (K == 0) {
            x[K] = cadd(E, O);
            x[K + N / 2] = csub(E, O);
        } else if constexpr (4 * K == N) {          // w^K = DIR*i
            if constexpr (DIR > 0) {
                x[K] = make_float2(E.x - O.y, E.y + O.x);
                x[K + N / 2] = make_float2(E.x + O.y, E.y - O.x);
            } else {
                x[K] = make_float2(E.x + O.y, E.y - O.x);
                x[K + N / 2] = make_float2(E.x - O.y, E.y + O.x);
            }
        } else if constexpr (8 * K == N) {          // w^K = (1 + DIR*i)/sqrt2
            const float2 u = DIR > 0 ? make_float2(O.x - O.y, O.x + O.y) : make_float2(O.x + O.y, O.y - O.x);
            x[K] = cfma_s(u, h, E);
            x[K + N / 2] = cfma_s(u, -h, E);
        } else if constexpr (8 * K == 3 * N) {      // w^K = (-1 + DIR*i)/sqrt2
            const float2 u = DIR > 0 ? make_float2(O.x + O.y, O.y - O.x) : make_float2(O.x - O.y, O.x + O.y);
            x[K] = cfma_s(u, -h, E);
            x[K + N / 2] = cfma_s(u, h, E);
        } else {
            constexpr float c = (float)cx_cos(2.0 * kPi * K / N);
            constexpr float s = (float)(DIR * cx_sin(2.0 * kPi * K / N));
            const float2 t = make_float2(fmaf(-O.y, s, O.x * c), fmaf(O.x, s, O.y * c));
            x[K] = cadd(E, t);
            x[K + N / 2] = csub(E, t);
        }
    }
    template <int... K>
    static CUPSS_HD void combine(float2 (&x)[N], const float2 (&e)[N / 2], const float2 (&o)[N / 2],
                                 cupss_std::integer_sequence<int, K...>) {
        (comb1<K>(x, e, o), ...);
    }
    static CUPSS_HD void run(float2 (&x)[N]) {
        float2 e[N / 2], o[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; ++i) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; }
        Dft<N / 2, DIR>::run(e);
        Dft<N / 2, DIR>::run(o);
        combine(x, e, o, cupss_std::make_integer_sequence<int, N / 2>{});
    }
};
template <int DIR>
struct Dft<1, DIR> { static CUPSS_HD void run(float2 (&)[1]) {} };
template <int DIR>
struct Dft<2, DIR> {
    static CUPSS_HD void run(float2 (&x)[2]) {
        const float2 a = x[0], b = x[1];
        x[0] = cadd(a, b); x[1] = csub(a, b);
    }
};

// ---------------------------------------------------------------- band-limited input: pruned DFT of the x pass (kernels_x3.cu)
// a * conj(t), rotated by (-i)^ROT -- the four sign patterns of the same two multiplies and two fused multiply-adds
template <int ROT>
CUPSS_HD float2 cmul_conj_rot(float2 y, float2 t) {
    if constexpr (ROT == 0) return make_float2(fmaf(y.y, t.y, y.x * t.x), fmaf(y.y, t.x, -(y.x * t.y)));
    else if constexpr (ROT == 1) return make_float2(fmaf(y.y, t.x, -(y.x * t.y)), fmaf(-y.y, t.y, -(y.x * t.x)));
    else if constexpr (ROT == 2) return make_float2(fmaf(-y.y, t.y, -(y.x * t.x)), fmaf(-y.y, t.x, y.x * t.y));
    else return make_float2(fmaf(-y.y, t.x, y.x * t.y), fmaf(y.y, t.y, y.x * t.x));
}

// x[q] <- (-i)^q x[q] conj(tw[(q-1) M]) for q >= 1 (M = 2 R0: row stride of the strided level's twiddle table)
template <int R0, int... Q>
CUPSS_HD void x3_twiddle_rot(float2 (&x)[R0], const float2* tw, cupss_std::integer_sequence<int, Q...>) {
    ((x[Q] = Q == 0 ? x[Q] : cmul_conj_rot<Q % 4>(x[Q], tw[(Q > 0 ? Q - 1 : 0) * 2 * R0])), ...);
}

// Inverse R0-point DFT (exponent +) of a vector whose entries Q .. R0-Q-1 are zero (Q = R0/4): the strided level of the x
// pass when the input is band-limited to |k| <= sx/4 -- the dealiased field of a cubic term.  With x'[q'] = x[q' - Q] (indices
// mod R0, q' = 0 .. R0/2-1):  X[p] = (-i)^p Y[p],  Y[2r] = DFT_{R0/2}(x')[r],  Y[2r+1] = DFT_{R0/2}(x'[q'] w^q')[r]  (w = e^{2 pi i / R0}).
// A full butterfly stage, half of the loads and their mirror-pair bookkeeping disappear.  Returns Y in natural order; the caller
// folds (-i)^p into the twiddle multiplication that follows (cmul_conj_rot).
template <int R0>
struct PrunedDft {
    static constexpr int Q = R0 / 4, N2 = R0 / 2;
    template <int I>
    static CUPSS_HD float2 tw1(float2 v) {   // v * w^I, the constant folded at compile time
        constexpr float c = (float)cx_cos(2.0 * kPi * I / R0), sn = (float)cx_sin(2.0 * kPi * I / R0);
        return I == 0 ? v : make_float2(fmaf(-v.y, sn, v.x * c), fmaf(v.x, sn, v.y * c));
    }
    template <int... I>
    static CUPSS_HD void twiddle(float2 (&o)[N2], cupss_std::integer_sequence<int, I...>) {
        ((o[I] = tw1<I>(o[I])), ...);
    }
    // lo = x[0 .. Q-1], hi = x[R0-Q .. R0-1]
    static CUPSS_HD void run(const float2 (&lo)[Q], const float2 (&hi)[Q], float2 (&y)[R0]) {
        float2 e[N2], o[N2];
#pragma unroll
        for (int i = 0; i < Q; ++i) { e[i] = hi[i]; e[Q + i] = lo[i]; }
#pragma unroll
        for (int i = 0; i < N2; ++i) o[i] = e[i];
        twiddle(o, cupss_std::make_integer_sequence<int, N2>{});
        Dft<N2, +1>::run(e);
        Dft<N2, +1>::run(o);
#pragma unroll
        for (int r = 0; r < N2; ++r) { y[2 * r] = e[r]; y[2 * r + 1] = o[r]; }
    }
};

// ---------------------------------------------------------------- level plan: L -> radices
template <int L> struct FftLevels;
#define CUPSS_LEVELS(L_, N_, A_, B_, C_, D_)                                              \
    template <> struct FftLevels<L_> {                                                    \
        static constexpr int n = N_;                                                      \
        __host__ __device__ static constexpr int rad(int l) { return l == 0 ? A_ : (l == 1 ? B_ : (l == 2 ? C_ : D_)); } \
        __host__ __device__ static constexpr int min_rad() {                                                          \
            int m = A_;                                                                                               \
            if (N_ > 1 && B_ < m) m = B_;                                                                             \
            if (N_ > 2 && C_ < m) m = C_;                                                                             \
            if (N_ > 3 && D_ < m) m = D_;                                                                             \
            return m;                                                                                                 \
        }                                                                                                             \
        __host__ __device__ static constexpr int max_rad() {                                                          \
            int m = A_;                                                                                               \
            if (N_ > 1 && B_ > m) m = B_;                                                                             \
            if (N_ > 2 && C_ > m) m = C_;                                                                             \
            if (N_ > 3 && D_ > m) m = D_;                                                                             \
            return m;                                                                                                 \
        }                                                                                                             \
    };
CUPSS_LEVELS(1, 1, 1, 1, 1, 1)
CUPSS_LEVELS(2, 1, 2, 1, 1, 1)
CUPSS_LEVELS(4, 1, 4, 1, 1, 1)
CUPSS_LEVELS(8, 1, 8, 1, 1, 1)
CUPSS_LEVELS(16, 1, 16, 1, 1, 1)
CUPSS_LEVELS(32, 2, 4, 8, 1, 1)
CUPSS_LEVELS(64, 2, 8, 8, 1, 1)
CUPSS_LEVELS(128, 2, 16, 8, 1, 1)
CUPSS_LEVELS(256, 2, 16, 16, 1, 1)
CUPSS_LEVELS(512, 3, 8, 8, 8, 1)
CUPSS_LEVELS(1024, 3, 16, 8, 8, 1)
CUPSS_LEVELS(2048, 3, 16, 16, 8, 1)
CUPSS_LEVELS(4096, 3, 16, 16, 16, 1)
CUPSS_LEVELS(8192, 4, 16, 8, 8, 8)
#undef CUPSS_LEVELS

template <int L, int LV> struct LevelGeom {
    static constexpr int R = FftLevels<L>::rad(LV);
    static constexpr int NB = LV == 0 ? 1 : (LV == 1 ? FftLevels<L>::rad(0) : (LV == 2 ? FftLevels<L>::rad(0) * FftLevels<L>::rad(1)
                                                                                      : FftLevels<L>::rad(0) * FftLevels<L>::rad(1) * FftLevels<L>::rad(2)));
    static constexpr int N = L / NB;     // block length at this level
    static constexpr int M = N / R;      // stride between the R points of one butterfly
    static constexpr int NV = L / R;     // virtual threads of this level
    static constexpr int TWS = L / N;    // twiddle w_N^(j q) = exp(-2*pi*i * TWS*j*q / L)
    // Level twiddle table (built on the host, level_twiddles()): entry TWOFF + (q-1)*M + j holds w_N^(j q),
    // q = 1..R-1, j = 0..M-1 -- consecutive j are consecutive words, so a warp's loads are conflict-free and every
    // address is "j + compile-time offset".  Levels with M == 1 have no twiddles.
    static constexpr int TWCNT = M > 1 ? (R - 1) * M : 0;
    static constexpr int TWOFF = LV == 0 ? 0 : LevelGeom<L, (LV > 0 ? LV - 1 : 0)>::TWOFF + LevelGeom<L, (LV > 0 ? LV - 1 : 0)>::TWCNT;
};
template <int L> struct TwTable {
    static constexpr int LEN_ = LevelGeom<L, 3>::TWOFF + LevelGeom<L, 3>::TWCNT;   // levels beyond n have R = 1: TWCNT = 0
    static constexpr int LEN = LEN_ > 0 ? LEN_ : 1;
};

// Frequency held at position p after the forward (digit-reversed) transform; its own inverse for symmetric
// radix lists, and in general the map the inverse transform expects on input.
template <int L>
CUPSS_HD unsigned freq_of_pos(unsigned p) {
    using F = FftLevels<L>;
    unsigned k = 0, w = 1, rem = p, n = L;
#pragma unroll
    for (int l = 0; l < F::n; ++l) {
        n /= F::rad(l);
        const unsigned q = rem / n;
        rem -= q * n;
        k += q * w;
        w *= F::rad(l);
    }
    return k;
}

#ifndef __CUDACC_RTC__
// Host: fill the level twiddle table of L (TwTable<L>::LEN entries).
template <int L, int LV>
inline void fill_level_twiddles(float2* out) {
    if constexpr (LV < FftLevels<L>::n) {
        using G = LevelGeom<L, LV>;
        if constexpr (G::M > 1) {
            for (int q = 1; q < G::R; ++q)
                for (int j = 0; j < G::M; ++j) {
                    const long long k = ((long long)G::TWS * j * q) % L;
                    const double a = -2.0 * kPi * (double)k / (double)L;
                    out[G::TWOFF + (q - 1) * G::M + j] = make_float2((float)__builtin_cos(a), (float)__builtin_sin(a));
                }
        }
        fill_level_twiddles<L, LV + 1>(out);
    }
}
#endif

// Position that holds frequency k after the forward transform (inverse map of freq_of_pos).
template <int L>
CUPSS_HD unsigned pos_of_freq(unsigned k) {
    using F = FftLevels<L>;
    unsigned p = 0, n = L, rem = k;
#pragma unroll
    for (int l = 0; l < F::n; ++l) {
        const unsigned r = F::rad(l);
        n /= r;
        p += (rem % r) * n;
        rem /= r;
    }
    return p;
}

// One level on one virtual thread; `x` holds the R points in butterfly order q = 0..R-1.
//   SIGN: sign of the exponent (-1 forward, +1 inverse).
//   DIF : true  -> butterfly, then twiddle exp(SIGN*2*pi*i*j*q/N)   (natural order in, digit-reversed out)
//         false -> twiddle, then butterfly                           (digit-reversed in, natural order out)
// `tw` is the level twiddle table of L (forward sign; conjugated on the fly for SIGN > 0).
template <int L, int LV, int SIGN, bool DIF>
CUPSS_HD void level_butterfly(float2 (&x)[LevelGeom<L, LV>::R], unsigned j, const float2* __restrict__ tw) {
    using G = LevelGeom<L, LV>;
    constexpr int R = G::R;
    if constexpr (!DIF && G::M > 1) {
        const float2* t = tw + G::TWOFF + j;
#pragma unroll
        for (int q = 1; q < R; ++q) x[q] = SIGN > 0 ? cmul_conj(x[q], t[(q - 1) * G::M]) : cmul(x[q], t[(q - 1) * G::M]);
    }
    Dft<R, SIGN>::run(x);
    if constexpr (DIF && G::M > 1) {
        const float2* t = tw + G::TWOFF + j;
#pragma unroll
        for (int q = 1; q < R; ++q) x[q] = SIGN > 0 ? cmul_conj(x[q], t[(q - 1) * G::M]) : cmul(x[q], t[(q - 1) * G::M]);
    }
}

// Same for two independent lines (the two columns / two jobs a thread carries): every twiddle is loaded once.
template <int L, int LV, int SIGN, bool DIF>
CUPSS_HD void level_butterfly2(float2 (&x)[LevelGeom<L, LV>::R], float2 (&y)[LevelGeom<L, LV>::R], unsigned j, const float2* __restrict__ tw) {
    using G = LevelGeom<L, LV>;
    constexpr int R = G::R;
    if constexpr (!DIF && G::M > 1) {
        const float2* t = tw + G::TWOFF + j;
#pragma unroll
        for (int q = 1; q < R; ++q) {
            const float2 w = t[(q - 1) * G::M];
            x[q] = SIGN > 0 ? cmul_conj(x[q], w) : cmul(x[q], w);
            y[q] = SIGN > 0 ? cmul_conj(y[q], w) : cmul(y[q], w);
        }
    }
    Dft<R, SIGN>::run(x);
    Dft<R, SIGN>::run(y);
    if constexpr (DIF && G::M > 1) {
        const float2* t = tw + G::TWOFF + j;
#pragma unroll
        for (int q = 1; q < R; ++q) {
            const float2 w = t[(q - 1) * G::M];
            x[q] = SIGN > 0 ? cmul_conj(x[q], w) : cmul(x[q], w);
            y[q] = SIGN > 0 ? cmul_conj(y[q], w) : cmul(y[q], w);
        }
    }
}

}  // namespace cupss
