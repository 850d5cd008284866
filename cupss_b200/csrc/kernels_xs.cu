// kernels_xs.cu -- one-job variant of the generic ("stash") contiguous-axis pass for long lines (sx = 1024, 2048, 4096).
//
// Same algorithm and the same arithmetic per line as the stash path of kernels_x.cu (two real lines as one complex line,
// inverse as decimation in frequency, the real fields of every input parked in shared memory, the monomials of every output
// evaluated point by point, forward as decimation in time) -- results are bit-identical -- but a slot is ONE job (a pair of
// lines, float2 elements) instead of two interleaved jobs (float4): at sx = 2048 a slot of the two-job kernel with its stash
// fills an SM's shared memory (37 KB per buffer, 1 + nIn buffers: one 256-thread CTA per SM, every phase exposed -- Model H
// 2048^2: 11-14 % of the HBM roofline, profiles/README.md).  Here a buffer is 17 KB, the transforms of all inputs run side
// by side in their stash buffers, a CTA has SX / 8 threads and two or three CTAs share an SM, so the global loads of one
// overlap the butterflies of the others.  Replaces the same reference code as kernels_x.cu (/root/reference/src/field.cpp:247-298, src/term.cpp:48-102,
// src/term_kernels.cu:48-70).
#include <cstdlib>

#include "kernels.h"

namespace cupss {

template <int SX> struct XsCfg {
    using F = FftLevels<SX>;
    static_assert(F::n == 3, "three-level line lengths only");
    static constexpr int RL = F::rad(2);
    static constexpr int XB = SX + SX / 16 + 1;              // padded line, float2 elements
    static constexpr int TW = TwTable<SX>::LEN;
    static constexpr int NT = SX / 16 > 256 ? 256 : SX / 16;  // default threads per CTA = virtual threads of a radix-16 level
};
// one pad element per 16: consecutive elements (outer levels), stride RL (innermost level) and the 8-wide groups of the
// middle level of 2048 = 16*16*8 all hit 16 different 64-bit banks per half warp
__device__ __forceinline__ unsigned xspad(unsigned idx) { return idx + (idx >> 4); }

// NT: threads per CTA.  TWG: the level twiddles are read from global memory through L1 instead of a shared-memory copy
// (16 KB at sx = 2048: one more CTA per SM when the stash is small).
// The inverse transforms of ALL inputs run side by side, each in place in its own stash buffer (a level is one loop over
// nIn x NV virtual threads and one barrier, not nIn of each), and so do the forward transforms of up to a.jobsPerCta outputs,
// each in a line buffer of its own: the CTA's threads stay busy on the radix-16 levels (SX / 16 virtual threads per line).
template <int SX, int NT, bool TWG>
__global__ void __launch_bounds__(NT, (SX >= 4096 ? 512 : 768) / NT) xstash1_kernel(const __grid_constant__ XArgs a) {
    using Cfg = XsCfg<SX>;
    constexpr unsigned XB = Cfg::XB, RL = Cfg::RL;
    using G0 = LevelGeom<SX, 0>;
    using G2 = LevelGeom<SX, 2>;
    extern __shared__ float2 smemS[];
    const float2* twS = TWG ? a.tw : smemS;
    float2* stash = smemS + (TWG ? 0 : Cfg::TW);      // real fields of the inputs: (line A, line B) per point, buffer g
    float2* lineBufs = stash + (unsigned)a.nIn * XB;    // a.jobsPerCta line buffers for the forward transforms
    const unsigned tid = threadIdx.x;
    const unsigned nIn = (unsigned)a.nIn, OB = (unsigned)a.jobsPerCta;

    if constexpr (!TWG) {
        for (unsigned i = tid; i < (unsigned)Cfg::TW; i += NT) smemS[i] = __ldg(a.tw + i);
        __syncthreads();
    }

    const long long l0 = 2ll * blockIdx.x;                 // lines A = l0, B = l0 + 1
    const bool vA = l0 < a.nlines, vB = l0 + 1 < a.nlines;
    const float2 z = make_float2(0.0f, 0.0f);

    // one shared -> shared level over `nb` buffers starting at `bufs`
    auto level_ss = [&](auto lvTag, auto signTag, auto difTag, float2* bufs, unsigned nb) {
        constexpr int LV = decltype(lvTag)::value;
        constexpr int SIGN = decltype(signTag)::value;
        constexpr bool DIF = decltype(difTag)::value;
        using G = LevelGeom<SX, LV>;
        constexpr unsigned R = G::R, M = G::M, N = G::N, NV = G::NV;
#pragma unroll 1
        for (unsigned w = tid; w < nb * NV; w += NT) {
            const unsigned b = w / NV, v = w % NV;
            float2* xb = bufs + b * XB;
            const unsigned blk = v / M, j = v % M, row0 = blk * N + j;
            // xspad(row0 + M q) = xspad(row0) + xspad(M q): M q is a multiple of 8 and row0 mod 16 < 8 whenever M = 8 (j < 8, N a
            // multiple of 16), so the low four bits never carry -- one padded base and compile-time offsets instead of a shift
            // and an add per access
            static_assert(M % 8 == 0 && N % 16 == 0, "padded offsets are folded at compile time");
            float2* xr = xb + xspad(row0);
            float2 x[R];
#pragma unroll
            for (unsigned q = 0; q < R; ++q) x[q] = xr[M * q + ((M * q) >> 4)];
            level_butterfly<SX, LV, SIGN, DIF>(x, j, twS);
#pragma unroll
            for (unsigned q = 0; q < R; ++q) xr[M * q + ((M * q) >> 4)] = x[q];
        }
    };
    using std::integral_constant;
    using L0 = integral_constant<int, 0>;
    using L1 = integral_constant<int, 1>;
    using Plus = integral_constant<int, 1>;
    using Minus = integral_constant<int, -1>;
    using Dif = integral_constant<bool, true>;
    using Dit = integral_constant<bool, false>;

    // ------------------------------------------------ inverse part: C2R of every input, in place in its stash buffer
    {
        constexpr unsigned R = G0::R, M = G0::M, NV = G0::NV;
#pragma unroll 1
        for (unsigned w = tid; w < nIn * NV; w += NT) {
            const unsigned g = w / NV, v = w % NV;
            const float2* pA = a.in[g] + l0 * a.pitch;
            const float2* pB = pA + a.pitch;
            const int kmax = a.kmax[g];
            float2* xb = stash + g * XB;
            float2 x[R];
#pragma unroll
            for (unsigned q = 0; q < R; ++q) {
                // C[idx] = A[k] + i B[k] (idx <= sx/2, k = idx)  |  conj(A[k]) + i conj(B[k]) (idx > sx/2, k = sx - idx);
                // v < M, so the half is known per q except on the row M*q == sx/2
                const unsigned idx = v + M * q;
                const bool upper = M * q > SX / 2u || (M * q == SX / 2u && v > 0);
                const unsigned k = upper ? SX - idx : idx;
                const bool live = (int)k <= kmax;
                float2 A = (live && vA) ? __ldg(pA + k) : z;
                float2 B = (live && vB) ? __ldg(pB + k) : z;
                if (k == 0 || 2 * k == SX) { A.y = 0.0f; B.y = 0.0f; }   // real-part projection of self-conjugate bins
                x[q] = upper ? make_float2(A.x + B.y, B.x - A.y) : make_float2(A.x - B.y, A.y + B.x);
            }
            level_butterfly<SX, 0, +1, true>(x, v, twS);
            float2* xr = xb + xspad(v);   // M is a multiple of 16 here: xspad(v + M q) = xspad(v) + xspad(M q)
            static_assert(M % 16 == 0, "padded offsets are folded at compile time");
#pragma unroll
            for (unsigned q = 0; q < R; ++q) xr[M * q + ((M * q) >> 4)] = x[q];
        }
    }
    __syncthreads();
    level_ss(L1{}, Plus{}, Dif{}, stash, nIn);
    __syncthreads();
#pragma unroll 1
    for (unsigned w = tid; w < nIn * (unsigned)G2::NV; w += NT) {
        const unsigned g = w / G2::NV, v = w % G2::NV;
        float2* sg = stash + g * XB + xspad(v * RL);   // the RL points of a virtual thread are contiguous (RL <= 16)
        float2 x[RL];
#pragma unroll
        for (unsigned q = 0; q < RL; ++q) x[q] = sg[q];
        level_butterfly<SX, 2, +1, true>(x, 0, twS);
#pragma unroll
        for (unsigned q = 0; q < RL; ++q) sg[q] = cscale(x[q], a.norm);
    }

    // ------------------------------------------------ products + forward part (R2C of the outputs, OB at a time)
    for (unsigned o0 = 0; o0 < (unsigned)a.nOut; o0 += OB) {
        const unsigned nb = min(OB, (unsigned)a.nOut - o0);
        __syncthreads();   // stash complete / the previous batch's untangle has read the line buffers
#pragma unroll 1
        for (unsigned w = tid; w < nb * (unsigned)G2::NV; w += NT) {
            const unsigned ob = w / G2::NV, v = w % G2::NV;
            const int o = (int)(o0 + ob);
            const unsigned base = xspad(v * RL);
            float2 x[RL];
#pragma unroll
            for (unsigned q = 0; q < RL; ++q) x[q] = z;
            // every monomial of this output: coef * r_f0 * r_f1 * ... multiplied left to right (computeProduct,
            // src/term.cpp:85-92), the monomials added in their order -- the same operations as kernels_x.cu
            for (int m = 0; m < a.nMono; ++m) {
                if (a.mono[m].out != o) continue;
                const float c = a.mono[m].coef;
                const int nf = a.mono[m].nfac;
                float2 pr[RL];
#pragma unroll
                for (unsigned q = 0; q < RL; ++q) pr[q] = make_float2(c, c);
                for (int f = 0; f < nf; ++f) {
                    const float2* sf = stash + (unsigned)a.mono[m].fac[f] * XB + base;
#pragma unroll
                    for (unsigned q = 0; q < RL; ++q) pr[q] = cmul2(pr[q], sf[q]);
                }
#pragma unroll
                for (unsigned q = 0; q < RL; ++q) x[q] = cadd(x[q], pr[q]);
            }
            level_butterfly<SX, 2, -1, false>(x, 0, twS);
            float2* xb = lineBufs + ob * XB + base;
#pragma unroll
            for (unsigned q = 0; q < RL; ++q) xb[q] = x[q];
        }
        __syncthreads();
        level_ss(L1{}, Minus{}, Dit{}, lineBufs, nb);
        __syncthreads();
        level_ss(L0{}, Minus{}, Dit{}, lineBufs, nb);
        __syncthreads();
        // untangle the two real lines: A[k] = (C[k] + conj C[sx-k]) / 2, B[k] = (C[k] - conj C[sx-k]) / (2i); two
        // neighbouring k per thread (128-bit stores)
        constexpr unsigned NP = SX / 4 + 1;   // k = 0, 2, ..., sx/2
#pragma unroll 1
        for (unsigned w = tid; w < nb * NP; w += NT) {
            const unsigned ob = w / NP, k = 2 * (w % NP);
            const float2* xb = lineBufs + ob * XB;
            float2* qA = a.out[o0 + ob] + l0 * a.pitch;
            float2* qB = qA + a.pitch;
            auto at = [&](unsigned kk) -> float2 { return xb[xspad(kk & (SX - 1))]; };
            auto split = [&](float2 Ck, float2 Cm, float2& A, float2& B) {
                A = make_float2(0.5f * (Ck.x + Cm.x), 0.5f * (Ck.y - Cm.y));
                B = make_float2(0.5f * (Ck.y + Cm.y), -0.5f * (Ck.x - Cm.x));
            };
            float2 A, B, An, Bn;
            split(at(k), at(SX - k), A, B);
            if (k + 1 <= SX / 2) {
                split(at(k + 1), at(SX - k - 1), An, Bn);
                if (vA) *reinterpret_cast<float4*>(qA + k) = make_float4(A.x, A.y, An.x, An.y);
                if (vB) *reinterpret_cast<float4*>(qB + k) = make_float4(B.x, B.y, Bn.x, Bn.y);
            } else {
                if (vA) qA[k] = A;
                if (vB) qB[k] = B;
            }
        }
    }
}

// ---------------------------------------------------------------- dispatch
static bool xs_disabled() {
    static const bool off = getenv("CUPSS_B200_NO_XS1") != nullptr;
    return off;
}
bool xstash1_supported(int sx) { return !xs_disabled() && (sx == 1024 || sx == 2048 || sx == 4096); }

static int xs_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
template <int SX> static size_t xs_smem(int nIn, int ob, bool twg) {
    return ((size_t)(twg ? 0 : XsCfg<SX>::TW) + (size_t)(ob + nIn) * XsCfg<SX>::XB) * sizeof(float2);
}
// inputs one launch can take: the line buffer and one stash buffer per input within the 227 KB a CTA may have
int xstash1_max_inputs(int sx) {
    size_t xb = 0;
    switch (sx) {
#define X(N) case N: xb = (size_t)XsCfg<N>::XB * sizeof(float2); break;
        X(1024) X(2048) X(4096)
#undef X
        default: return 0;
    }
    const long long fit = (long long)((226 * 1024) / xb) - 1;
    return (int)(fit < 1 ? 1 : (fit > XP_MAX_IN ? XP_MAX_IN : fit));
}

template <int SX, int NT, bool TWG>
static cudaError_t launch_xs2(XArgs& a, cudaStream_t st) {
    const size_t smem = xs_smem<SX>(a.nIn, a.jobsPerCta, TWG);
    if (smem > 226 * 1024) return cudaErrorInvalidValue;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(xstash1_kernel<SX, NT, TWG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    const unsigned grid = (unsigned)((a.nlines + 1) / 2);
    xstash1_kernel<SX, NT, TWG><<<grid, NT, smem, st>>>(a);
    return cudaGetLastError();
}
template <int SX>
static cudaError_t launch_xs(XArgs& a, cudaStream_t st) {
    static const int nt = xs_env("CUPSS_B200_XS_NT", XsCfg<SX>::NT >= 256 ? XsCfg<SX>::NT : 2 * XsCfg<SX>::NT);
    static const int twgEnv = xs_env("CUPSS_B200_XS_TWG", -1);
    static const int obEnv = xs_env("CUPSS_B200_XS_OB", 1);   // measured (Model H 2048^2): 2 buffers cost a CTA per SM and gain nothing
    // line buffers: outputs transformed side by side, as long as they fit
    int ob = a.nOut < obEnv ? a.nOut : obEnv;
    while (ob > 1 && xs_smem<SX>(a.nIn, ob, true) + 1024 > 226 * 1024) --ob;
    if (ob < 1) ob = 1;
    a.jobsPerCta = ob;
    // twiddles through L1 by default: the shared-memory copy costs a prologue per CTA and, with a small stash, a CTA per SM
    // (Model H 2048^2: x_con 0.050 -> 0.042 ms, x_dyn 0.059 -> 0.057 ms)
    bool twg = twgEnv >= 0 ? twgEnv != 0 : true;
    if (!twg && xs_smem<SX>(a.nIn, ob, false) > 226 * 1024) twg = true;   // the copy does not fit next to this many buffers
    if (nt == 2 * XsCfg<SX>::NT) return twg ? launch_xs2<SX, 2 * XsCfg<SX>::NT, true>(a, st) : launch_xs2<SX, 2 * XsCfg<SX>::NT, false>(a, st);
    return twg ? launch_xs2<SX, XsCfg<SX>::NT, true>(a, st) : launch_xs2<SX, XsCfg<SX>::NT, false>(a, st);
}

cudaError_t launch_xstash1(int sx, XArgs& a, cudaStream_t st) {
    switch (sx) {
        case 1024: return launch_xs<1024>(a, st);
        case 2048: return launch_xs<2048>(a, st);
        case 4096: return launch_xs<4096>(a, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cupss
