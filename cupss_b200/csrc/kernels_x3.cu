// kernels_x3.cu -- two-level variant of the contiguous-axis pass for sx = R0 * 2R0 (512 = 16*32, 128 = 8*16), used when
// the real-space stage is one output of one input (the Cahn-Hilliard / Allen-Cahn / Swift-Hohenberg class).
//
// Same algorithm as kernels_x.cu (two real lines as one complex line, inverse as decimation in frequency, pointwise
// real-space stage, forward as decimation in time), restructured around the resource that bounds that kernel --
// shared-memory bandwidth (profiles/r01d: l1tex 78 %, 5 shared-memory round trips per point):
//   * two levels instead of three: radix R0 on the strided level, radix R1 = 2*R0 on the innermost (contiguous) level,
//     whose inverse butterfly, normalisation, product and forward butterfly are fused on registers;
//   * a thread owns BOTH members j and M-j of every mirror pair of the strided level, so C[k] and C[sx-k] are formed
//     from one load of A[k], B[k] on the way in and untangled on registers on the way out: no exchange for the untangle
//     and half the global load instructions;
//   => 2 shared-memory round trips per point.
// Replaces the same reference code as kernels_x.cu (/root/reference/src/field.cpp:247-298, src/term.cpp:48-102).
#include <cstdlib>

#include "kernels.h"

namespace cupss {

template <int SX, int NT> struct X3Cfg {
    static constexpr int R0 = SX == 512 ? 16 : (SX == 128 ? 8 : 0);
    static constexpr int R1 = 2 * R0;             // innermost radix = stride M of the strided level
    static constexpr int M = R1;
    static constexpr int TL = R0;                 // threads per job (complex line): M/2 mirror pairs == R0 innermost blocks
    static constexpr int THREADS = NT;            // 16 warps per SM either way (the kernel needs ~120 registers): NT = 256 -> 2 CTAs
    static constexpr int CTAS = 512 / NT;        // of 8 warps, 128 -> 4 CTAs of 4 warps (finer barriers, more staggered phases)
    static constexpr int JOBS = THREADS / TL;
    static constexpr int LB = SX + 2 * R0;        // padded line, float2 elements: two pad elements after every R1
    static constexpr int TWLEN = (R0 - 1) * M;    // strided-level twiddles: entry (q-1)*M + j = exp(-2*pi*i*j*q/SX)
    static constexpr size_t SMEM = ((size_t)TWLEN + (size_t)JOBS * LB) * sizeof(float2);
};

template <int SX>
__device__ __forceinline__ unsigned x3pad(unsigned idx) { return idx + 2u * (idx / (unsigned)X3Cfg<SX, 256>::R1); }

// POLY: the real-space stage is c0*r^p0 + c1*r^p1 of the one input (powers up to 4, c1 = 0 for a single monomial) instead of
// the straight-line c*r^2 / c*r^3 of the headline class.
// PRUNE: the input is known (launcher) to be band-limited to kx <= SX/4; the inverse strided level runs PrunedDft.
template <int SX, int NT, int MINB = X3Cfg<SX, NT>::CTAS, bool POLY = false, bool PRUNE = false>
__global__ void __launch_bounds__(NT, MINB) xpass3_kernel(const __grid_constant__ XArgs a) {
    using Cfg = X3Cfg<SX, NT>;
    constexpr int R0 = Cfg::R0, R1 = Cfg::R1, M = Cfg::M, TL = Cfg::TL, LB = Cfg::LB, H = R0 / 2;
    extern __shared__ float2 smem2[];
    float2* twS = smem2;
    const unsigned tid = threadIdx.x, job = tid / TL, t = tid % TL;
    float2* xb = smem2 + Cfg::TWLEN + job * LB;

    for (unsigned i = tid; i < (unsigned)Cfg::TWLEN; i += Cfg::THREADS) twS[i] = __ldg(a.tw3 + i);

    const long long jg = (long long)blockIdx.x * Cfg::JOBS + job;
    const long long lA = 2 * jg, lB = lA + 1;
    const bool hasA = lA < a.nlines, hasB = lB < a.nlines;
    const bool t0 = t == 0;
    const unsigned jA = t0 ? 0u : t, jB = t0 ? (unsigned)(M / 2) : (unsigned)M - t;   // the thread's two rows of the strided level
    const float2 z = make_float2(0.0f, 0.0f);

    float2 xA[R0], xB[R0];

    // ------------------------------------------------ inverse, strided level: form C from the half-spectrum lines
    if constexpr (PRUNE) {
        constexpr int Q = R0 / 4;   // live: k = j + M q with q < Q, and k = M Q = SX/4 on row 0
        const float2* pa = a.in[0] + lA * a.pitch;
        const float2* pb = a.in[0] + lB * a.pitch;
        auto form = [&](unsigned k, float2& c, float2& m) {   // c = A[k] + i B[k],  m = conj(A[k]) + i conj(B[k]) = C[sx - k]
            float2 A = hasA ? __ldg(pa + k) : z;
            float2 B = hasB ? __ldg(pb + k) : z;
            if (k == 0) { A.y = 0.0f; B.y = 0.0f; }   // real-part projection of the self-conjugate bin
            c = make_float2(A.x - B.y, A.y + B.x);
            m = make_float2(A.x + B.y, B.x - A.y);
        };
        float2 cA[Q], mA[Q], cB[Q], mB[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            form(jA + M * q, cA[q], mA[q]);
            form(jB + M * q, cB[q], mB[q]);
        }
        // k = SX/4 (row 0 only): predicated loads in the same burst as the others, zeros elsewhere
        float2 ev, od;
        {
            const float2 A = (t0 && hasA) ? __ldg(pa + SX / 4) : z;
            const float2 B = (t0 && hasB) ? __ldg(pb + SX / 4) : z;
            const float2 c = make_float2(A.x - B.y, A.y + B.x), m = make_float2(A.x + B.y, B.x - A.y);
            ev = cadd(m, c); od = csub(m, c);   // Y[p] += m + (-1)^p c
        }
        // upper ends x[R0-Q .. R0-1] of the two rows: the mirrors (same placement as the general branch below, zeros dropped)
        float2 hiA[Q], hiB[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) hiB[i] = t0 ? mB[Q - 1 - i] : mA[Q - 1 - i];
        hiA[0] = t0 ? z : mB[Q - 1];   // row 0: x[R0-Q] is the mirror of k = SX/4, added below
#pragma unroll
        for (int i = 1; i < Q; ++i) hiA[i] = t0 ? mA[Q - i] : mB[Q - 1 - i];
        PrunedDft<R0>::run(cA, hiA, xA);
        PrunedDft<R0>::run(cB, hiB, xB);
        if (t0) {   // (two lanes of the warp: the other lanes hold zeros in ev / od)
#pragma unroll
            for (int p2 = 0; p2 < R0; p2 += 2) { xA[p2] = cadd(xA[p2], ev); xA[p2 + 1] = cadd(xA[p2 + 1], od); }
        }
    } else {
        const float2* pa = a.in[0] + lA * a.pitch;
        const float2* pb = a.in[0] + lB * a.pitch;
        const int kmax = a.kmax[0];
        float2 mA[H], mB[H];
        auto form = [&](unsigned k, float2& c, float2& m) {   // c = A[k] + i B[k],  m = conj(A[k]) + i conj(B[k]) = C[sx - k]
            const bool live = (int)k <= kmax;
            float2 A = (live && hasA) ? __ldg(pa + k) : z;
            float2 B = (live && hasB) ? __ldg(pb + k) : z;
            if (k == 0 || 2 * k == SX) { A.y = 0.0f; B.y = 0.0f; }   // real-part projection of self-conjugate bins
            c = make_float2(A.x - B.y, A.y + B.x);
            m = make_float2(A.x + B.y, B.x - A.y);
        };
#pragma unroll
        for (int q = 0; q < H; ++q) {
            form(jA + M * q, xA[q], mA[q]);
            form(jB + M * q, xB[q], mB[q]);
        }
        float2 cMid = z, mMid;
        if (t0) form(SX / 2, cMid, mMid);   // k = sx/2 belongs to row 0 (register H of the thread that owns rows 0 and M/2)
        // mirrors: rows (j, M - j) for t >= 1; row 0 mirrors onto itself and so does row M/2 for t == 0
#pragma unroll
        for (int i = 0; i < H; ++i) xB[H + i] = t0 ? mB[H - 1 - i] : mA[H - 1 - i];
        xA[H] = t0 ? cMid : mB[H - 1];
#pragma unroll
        for (int i = 1; i < H; ++i) xA[H + i] = t0 ? mA[H - i] : mB[H - 1 - i];
        Dft<R0, +1>::run(xA);
        Dft<R0, +1>::run(xB);
    }
    __syncthreads();   // twiddle table complete
    if constexpr (PRUNE) {   // X[q] = (-i)^q Y[q]: the rotation rides on the twiddle multiplication
        x3_twiddle_rot(xA, twS + jA, cupss_std::make_integer_sequence<int, R0>{});
        x3_twiddle_rot(xB, twS + jB, cupss_std::make_integer_sequence<int, R0>{});
    } else {
#pragma unroll
        for (int q = 1; q < R0; ++q) {
            xA[q] = cmul_conj(xA[q], twS[(q - 1) * M + jA]);
            xB[q] = cmul_conj(xB[q], twS[(q - 1) * M + jB]);
        }
    }
#pragma unroll
    for (int q = 0; q < R0; ++q) {
        xb[x3pad<SX>(jA + M * q)] = xA[q];
        xb[x3pad<SX>(jB + M * q)] = xB[q];
    }
    // a job's TL threads sit in ONE warp (TL divides 32) and its line is touched by nobody else: warp-level barriers are
    // enough between the levels, so the warps of a CTA drift apart and their load / butterfly / store phases interleave
    static_assert(32 % TL == 0, "a job must not straddle warps");
    __syncwarp();

    // ------------------------------------------------ innermost level: inverse butterfly, product, forward butterfly
    {
        float2 y[R1];
        float4* blk = reinterpret_cast<float4*>(xb + t * (R1 + 2));   // block t: R1 contiguous points (16-byte aligned)
#pragma unroll
        for (int i = 0; i < R1 / 2; ++i) {
            const float4 v = blk[i];
            y[2 * i] = make_float2(v.x, v.y); y[2 * i + 1] = make_float2(v.z, v.w);
        }
        Dft<R1, +1>::run(y);
        // both real lines of the job at once: (rx, ry) = y * norm, then the left-to-right product (r*r)*r of
        // computeProduct (src/term.cpp:85-92) and the coefficient, as packed FMUL2 (same IEEE operations per component)
        const float2 norm2 = make_float2(a.norm, a.norm);
        const float2 c02 = make_float2(a.mono[0].coef, a.mono[0].coef);
        const bool cube = a.mono[0].nfac == 3;   // the launcher only sends single monomials c*r^2 / c*r^3 here (warp-uniform)
        if constexpr (POLY) {
            const int p0 = a.mono[0].nfac, p1 = a.nMono > 1 ? a.mono[1].nfac : 0;   // warp-uniform
            const float c1 = a.nMono > 1 ? a.mono[1].coef : 0.0f;
            const float2 c12 = make_float2(c1, c1);
#pragma unroll
            for (int i = 0; i < R1; ++i) {
                const float2 r = cmul2(y[i], norm2);
                const float2 r2 = cmul2(r, r), r3 = cmul2(r2, r), r4 = cmul2(r3, r);
                const float2 one = make_float2(1.0f, 1.0f);
                const float2 w0 = p0 == 1 ? r : (p0 == 2 ? r2 : (p0 == 3 ? r3 : r4));
                const float2 w1 = p1 == 0 ? one : (p1 == 1 ? r : (p1 == 2 ? r2 : (p1 == 3 ? r3 : r4)));
                y[i] = cadd(cmul2(c02, w0), cmul2(c12, w1));
            }
        } else if (cube) {
#pragma unroll
            for (int i = 0; i < R1; ++i) {
                const float2 r = cmul2(y[i], norm2);
                y[i] = cmul2(c02, cmul2(cmul2(r, r), r));
            }
        } else {
#pragma unroll
            for (int i = 0; i < R1; ++i) {
                const float2 r = cmul2(y[i], norm2);
                y[i] = cmul2(c02, cmul2(r, r));
            }
        }
        Dft<R1, -1>::run(y);
#pragma unroll
        for (int i = 0; i < R1 / 2; ++i) blk[i] = make_float4(y[2 * i].x, y[2 * i].y, y[2 * i + 1].x, y[2 * i + 1].y);
    }
    __syncwarp();

    // ------------------------------------------------ forward, strided level (twiddle, butterfly) + untangle on registers
#pragma unroll
    for (int q = 0; q < R0; ++q) {
        xA[q] = xb[x3pad<SX>(jA + M * q)];
        xB[q] = xb[x3pad<SX>(jB + M * q)];
    }
#pragma unroll
    for (int q = 1; q < R0; ++q) {
        xA[q] = cmul(xA[q], twS[(q - 1) * M + jA]);
        xB[q] = cmul(xB[q], twS[(q - 1) * M + jB]);
    }
    Dft<R0, -1>::run(xA);
    Dft<R0, -1>::run(xB);
    // xA[r] = C[jA + M r], xB[r] = C[jB + M r];  A[k] = (C[k] + conj C[sx-k]) / 2,  B[k] = (C[k] - conj C[sx-k]) / (2i)
    float2* qa = a.out[0] + lA * a.pitch;
    float2* qb = a.out[0] + lB * a.pitch;
    // 0.5*(u +- v) as fma(+-0.5, v, 0.5*u): the halvings are exact, so the single rounding is that of u +- v
    const float2 half2 = make_float2(0.5f, 0.5f);
    auto emit = [&](unsigned k, float2 Ck, float2 Cm) {
        const float2 h = cmul2(Ck, half2);
        if (hasA) qa[k] = make_float2(fmaf(0.5f, Cm.x, h.x), fmaf(-0.5f, Cm.y, h.y));
        if (hasB) qb[k] = make_float2(fmaf(0.5f, Cm.y, h.y), fmaf(0.5f, Cm.x, -h.x));
    };
#pragma unroll
    for (int r = 0; r < H; ++r) {
        const float2 CmA = t0 ? xA[(R0 - r) % R0] : xB[R0 - 1 - r];
        const float2 CmB = t0 ? xB[R0 - 1 - r] : xA[R0 - 1 - r];
        emit(jA + M * r, xA[r], CmA);
        emit(jB + M * r, xB[r], CmB);
    }
    if (t0) emit(SX / 2, xA[H], xA[H]);
}

// Sum of powers (several inputs, one output, every monomial c * r_g^p with p <= 4 of ONE input -- the three squared
// gradients of KPZ): the inverse levels run input by input and the monomials are accumulated on registers at the
// innermost level, in monomial order like the generic kernels; one forward transform at the end.  128-thread CTAs,
// three per SM (the accumulators need ~160 registers).
template <int SX>
__global__ void __launch_bounds__(128, 3) xpass3s_kernel(const __grid_constant__ XArgs a) {
    using Cfg = X3Cfg<SX, 128>;
    constexpr int R0 = Cfg::R0, R1 = Cfg::R1, M = Cfg::M, TL = Cfg::TL, LB = Cfg::LB, H = R0 / 2;
    extern __shared__ float2 smem2[];
    float2* twS = smem2;
    const unsigned tid = threadIdx.x, job = tid / TL, t = tid % TL;
    float2* xb = smem2 + Cfg::TWLEN + job * LB;

    for (unsigned i = tid; i < (unsigned)Cfg::TWLEN; i += Cfg::THREADS) twS[i] = __ldg(a.tw3 + i);

    const long long jg = (long long)blockIdx.x * Cfg::JOBS + job;
    const long long lA = 2 * jg, lB = lA + 1;
    const bool hasA = lA < a.nlines, hasB = lB < a.nlines;
    const bool t0 = t == 0;
    const unsigned jA = t0 ? 0u : t, jB = t0 ? (unsigned)(M / 2) : (unsigned)M - t;   // the thread's two rows of the strided level
    const float2 z = make_float2(0.0f, 0.0f);

    float2 xA[R0], xB[R0];
    // accumulator of the real-space term between inputs: a second line buffer per job, private to the thread that owns
    // the innermost block (registers cannot hold it next to the strided level's 2 x R0 points without spilling)
    float4* ab = reinterpret_cast<float4*>(smem2 + Cfg::TWLEN + Cfg::JOBS * LB + job * LB + t * (R1 + 2));
    __syncthreads();   // twiddle table complete

#pragma unroll 1
    for (int g = 0; g < a.nIn; ++g) {
    // ------------------------------------------------ inverse, strided level: form C from the half-spectrum lines
    {
        const float2* pa = a.in[g] + lA * a.pitch;
        const float2* pb = a.in[g] + lB * a.pitch;
        const int kmax = a.kmax[g];
        float2 mA[H], mB[H];
        auto form = [&](unsigned k, float2& c, float2& m) {   // c = A[k] + i B[k],  m = conj(A[k]) + i conj(B[k]) = C[sx - k]
            const bool live = (int)k <= kmax;
            float2 A = (live && hasA) ? __ldg(pa + k) : z;
            float2 B = (live && hasB) ? __ldg(pb + k) : z;
            if (k == 0 || 2 * k == SX) { A.y = 0.0f; B.y = 0.0f; }   // real-part projection of self-conjugate bins
            c = make_float2(A.x - B.y, A.y + B.x);
            m = make_float2(A.x + B.y, B.x - A.y);
        };
#pragma unroll
        for (int q = 0; q < H; ++q) {
            form(jA + M * q, xA[q], mA[q]);
            form(jB + M * q, xB[q], mB[q]);
        }
        float2 cMid = z, mMid;
        if (t0) form(SX / 2, cMid, mMid);   // k = sx/2 belongs to row 0 (register H of the thread that owns rows 0 and M/2)
        // mirrors: rows (j, M - j) for t >= 1; row 0 mirrors onto itself and so does row M/2 for t == 0
#pragma unroll
        for (int i = 0; i < H; ++i) xB[H + i] = t0 ? mB[H - 1 - i] : mA[H - 1 - i];
        xA[H] = t0 ? cMid : mB[H - 1];
#pragma unroll
        for (int i = 1; i < H; ++i) xA[H + i] = t0 ? mA[H - i] : mB[H - 1 - i];
    }
    Dft<R0, +1>::run(xA);
    Dft<R0, +1>::run(xB);
#pragma unroll
    for (int q = 1; q < R0; ++q) {
        xA[q] = cmul_conj(xA[q], twS[(q - 1) * M + jA]);
        xB[q] = cmul_conj(xB[q], twS[(q - 1) * M + jB]);
    }
#pragma unroll
    for (int q = 0; q < R0; ++q) {
        xb[x3pad<SX>(jA + M * q)] = xA[q];
        xb[x3pad<SX>(jB + M * q)] = xB[q];
    }
    // a job's TL threads sit in ONE warp (TL divides 32) and its line is touched by nobody else: warp-level barriers are
    // enough between the levels, so the warps of a CTA drift apart and their load / butterfly / store phases interleave
    static_assert(32 % TL == 0, "a job must not straddle warps");
    __syncwarp();

    // ------------------------------------------------ innermost level: inverse butterfly, this input's monomials
    {
        float2 y[R1];
        const float4* blk = reinterpret_cast<const float4*>(xb + t * (R1 + 2));   // block t: R1 contiguous points (16-byte aligned)
#pragma unroll
        for (int i = 0; i < R1 / 2; ++i) {
            const float4 v = blk[i];
            y[2 * i] = make_float2(v.x, v.y); y[2 * i + 1] = make_float2(v.z, v.w);
        }
        Dft<R1, +1>::run(y);
        const float2 norm2 = make_float2(a.norm, a.norm);
#pragma unroll
        for (int i = 0; i < R1; ++i) y[i] = cmul2(y[i], norm2);
        // r^p is the left-to-right product ((r*r)*r)*r of computeProduct (src/term.cpp:85-92); both lines of the job at once.
        // The launcher sends exactly one monomial per input, in input order (mono[g] belongs to input g).
        const float2 cm = make_float2(a.mono[g].coef, a.mono[g].coef);
        const int pm = a.mono[g].nfac;   // warp-uniform
#pragma unroll
        for (int i = 0; i < R1; ++i) {
            float2 pw = y[i];
            if (pm >= 2) pw = cmul2(pw, y[i]);
            if (pm >= 3) pw = cmul2(pw, y[i]);
            if (pm >= 4) pw = cmul2(pw, y[i]);
            y[i] = cmul2(cm, pw);
        }
        if (g > 0) {   // add what the earlier inputs left, in input (= monomial) order
#pragma unroll
            for (int i = 0; i < R1 / 2; ++i) {
                const float4 v = ab[i];
                y[2 * i] = cadd(make_float2(v.x, v.y), y[2 * i]); y[2 * i + 1] = cadd(make_float2(v.z, v.w), y[2 * i + 1]);
            }
        }
        if (g + 1 < a.nIn) {
#pragma unroll
            for (int i = 0; i < R1 / 2; ++i) ab[i] = make_float4(y[2 * i].x, y[2 * i].y, y[2 * i + 1].x, y[2 * i + 1].y);
        } else {
            // last input: forward butterfly of the accumulated real-space term, straight from registers
            Dft<R1, -1>::run(y);
            float4* wb = reinterpret_cast<float4*>(xb + t * (R1 + 2));
#pragma unroll
            for (int i = 0; i < R1 / 2; ++i) wb[i] = make_float4(y[2 * i].x, y[2 * i].y, y[2 * i + 1].x, y[2 * i + 1].y);
        }
    }
    __syncwarp();   // the line buffer is re-used by the next input / read by the strided forward level
    }   // inputs

    // ------------------------------------------------ forward, strided level (twiddle, butterfly) + untangle on registers
#pragma unroll
    for (int q = 0; q < R0; ++q) {
        xA[q] = xb[x3pad<SX>(jA + M * q)];
        xB[q] = xb[x3pad<SX>(jB + M * q)];
    }
#pragma unroll
    for (int q = 1; q < R0; ++q) {
        xA[q] = cmul(xA[q], twS[(q - 1) * M + jA]);
        xB[q] = cmul(xB[q], twS[(q - 1) * M + jB]);
    }
    Dft<R0, -1>::run(xA);
    Dft<R0, -1>::run(xB);
    // xA[r] = C[jA + M r], xB[r] = C[jB + M r];  A[k] = (C[k] + conj C[sx-k]) / 2,  B[k] = (C[k] - conj C[sx-k]) / (2i)
    float2* qa = a.out[0] + lA * a.pitch;
    float2* qb = a.out[0] + lB * a.pitch;
    // 0.5*(u +- v) as fma(+-0.5, v, 0.5*u): the halvings are exact, so the single rounding is that of u +- v
    const float2 half2 = make_float2(0.5f, 0.5f);
    auto emit = [&](unsigned k, float2 Ck, float2 Cm) {
        const float2 h = cmul2(Ck, half2);
        if (hasA) qa[k] = make_float2(fmaf(0.5f, Cm.x, h.x), fmaf(-0.5f, Cm.y, h.y));
        if (hasB) qb[k] = make_float2(fmaf(0.5f, Cm.y, h.y), fmaf(0.5f, Cm.x, -h.x));
    };
#pragma unroll
    for (int r = 0; r < H; ++r) {
        const float2 CmA = t0 ? xA[(R0 - r) % R0] : xB[R0 - 1 - r];
        const float2 CmB = t0 ? xB[R0 - 1 - r] : xA[R0 - 1 - r];
        emit(jA + M * r, xA[r], CmA);
        emit(jB + M * r, xB[r], CmB);
    }
    if (t0) emit(SX / 2, xA[H], xA[H]);
}

template <int SX, int NT, int MINB = X3Cfg<SX, NT>::CTAS, bool POLY = false, bool PRUNE = false>
static cudaError_t launch_x3_size(XArgs& a, cudaStream_t st) {
    using Cfg = X3Cfg<SX, NT>;
    static bool attr = false;
    if (!attr) {
        if (Cfg::SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(xpass3_kernel<SX, NT, MINB, POLY, PRUNE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
            if (e != cudaSuccess) return e;
        }
        attr = true;
    }
    const long long njobs = (a.nlines + 1) / 2;
    const unsigned grid = (unsigned)((njobs + Cfg::JOBS - 1) / Cfg::JOBS);
    xpass3_kernel<SX, NT, MINB, POLY, PRUNE><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a);
    return cudaGetLastError();
}

template <int SX>
static cudaError_t launch_x3s_size(XArgs& a, cudaStream_t st) {
    using Cfg = X3Cfg<SX, 128>;
    constexpr size_t SMEM = Cfg::SMEM + (size_t)Cfg::JOBS * Cfg::LB * sizeof(float2);   // + the accumulator lines
    static bool attr = false;
    if (!attr) {
        if (SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(xpass3s_kernel<SX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
            if (e != cudaSuccess) return e;
        }
        attr = true;
    }
    const long long njobs = (a.nlines + 1) / 2;
    const unsigned grid = (unsigned)((njobs + Cfg::JOBS - 1) / Cfg::JOBS);
    xpass3s_kernel<SX><<<grid, Cfg::THREADS, SMEM, st>>>(a);
    return cudaGetLastError();
}

// sum-of-powers variants: the two-level kernel for sx = 512 and 128, the three-level one (kernels_x4.cu) for its sizes
bool xpass3_sumpow_supported(int sx) {
    static const bool off = [] { const char* e = getenv("CUPSS_B200_NO_X3S"); return e && e[0] == '1'; }();
    return !off && (sx == 512 || sx == 128 || xpass4_supported(sx));
}
cudaError_t launch_xpass3_sumpow(int sx, XArgs& a, cudaStream_t st) {
    if (xpass4_supported(sx)) return launch_xpass4_sumpow(sx, a, st);
    if (sx == 512) return launch_x3s_size<512>(a, st);
    if (sx == 128) return launch_x3s_size<128>(a, st);
    return cudaErrorInvalidValue;
}

bool xpass3_supported(int sx) { return sx == 512 || sx == 128 || xpass4_supported(sx); }

cudaError_t launch_xpass3(int sx, XArgs& a, cudaStream_t st) {
    if (xpass4_supported(sx)) return launch_xpass4(sx, a, st);   // long lines: the three-level kernel (kernels_x4.cu)
    const bool straight = a.nMono == 1 && (a.mono[0].nfac == 2 || a.mono[0].nfac == 3);
    if (!straight) {   // two monomials / other powers of the one input
        if (sx == 512) return launch_x3_size<512, 128, 4, true>(a, st);
        if (sx == 128) return launch_x3_size<128, 256, 2, true>(a, st);
        return cudaErrorInvalidValue;
    }
    static const int nt = [] { const char* e = getenv("CUPSS_B200_X3_THREADS"); return e ? atoi(e) : 128; }();   // measured (profiles/README.md): 0.202 ms at 128, 0.208 at 64, 0.229 at 256
    static const bool noPrune = [] { const char* e = getenv("CUPSS_B200_X3_NOPRUNE"); return e && e[0] == '1'; }();
    // the pruned instantiation fits 96 registers (4 bytes of spills): five 4-warp CTAs per SM, 0.199 -> 0.192 ms (profiles/README.md, r2j)
    static const int minb = [] { const char* e = getenv("CUPSS_B200_X3_MINB"); return e ? atoi(e) : 5; }();
    if (sx == 512 && a.kmax[0] == 128 && nt == 128 && !noPrune)   // band-limited to sx/4: pruned strided level
        return minb == 5 ? launch_x3_size<512, 128, 5, false, true>(a, st) : launch_x3_size<512, 128, 4, false, true>(a, st);
    if (sx == 512) return nt == 256 ? launch_x3_size<512, 256>(a, st) : (nt == 64 ? launch_x3_size<512, 64>(a, st) : launch_x3_size<512, 128>(a, st));
    if (sx == 128) return launch_x3_size<128, 256>(a, st);
    return cudaErrorInvalidValue;
}

int host_x3_twiddles(int sx, float2* out) {
    if (xpass4_supported(sx)) return host_x4_twiddles(sx, out);
    if (!xpass3_supported(sx)) return 0;
    const int R0 = sx == 512 ? 16 : 8, M = 2 * R0;
    for (int q = 1; q < R0; ++q)
        for (int j = 0; j < M; ++j) {
            const double ang = -2.0 * kPi * (double)((j * q) % sx) / (double)sx;
            out[(q - 1) * M + j] = make_float2((float)__builtin_cos(ang), (float)__builtin_sin(ang));
        }
    return (R0 - 1) * M;
}

}  // namespace cupss
