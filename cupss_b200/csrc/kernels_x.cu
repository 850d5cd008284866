// kernels_x.cu -- the contiguous-axis pass: C2R -> real-space products -> R2C in one kernel.
//
// Replaces, per product term, the x part of cufftExecC2C (inverse and forward), normalize_k and
// computeProduct_k of the reference (/root/reference/src/field.cpp:247-298, src/term.cpp:48-102,
// src/term_kernels.cu:48-70, src/field_kernels.cu:115-128): the dealiased real fields never exist in HBM.
//
// A "job" is a PAIR of neighbouring x lines (same array, rows 2j and 2j+1).  Two real lines a, b are
// transformed as ONE complex line c = a + i b of sx points (no half-length untangling twiddles):
//   inverse: C[k] = A[k] + i B[k] (k <= sx/2),  C[sx-k] = conj(A[k]) + i conj(B[k])
//   forward: A[k] = (C[k] + conj C[sx-k]) / 2,   B[k] = (C[k] - conj C[sx-k]) / (2i)
// T = sx/E threads per job, several jobs per CTA; consecutive lanes own consecutive points so every
// global access is a contiguous run of T float2.
#include "kernels.h"

namespace cupss {

template <int SX>
struct LineEx {
    float2* buf;
    __device__ __forceinline__ static int p(int idx) {
        constexpr int R0 = FftPlan<SX>::R0;
        return idx + idx / R0;   // stride R0+1 float2 between butterflies of the first pass: conflict-free
    }
    __device__ __forceinline__ void st(int idx, float2 v) { buf[p(idx)] = v; }
    __device__ __forceinline__ float2 ld(int idx) const { return buf[p(idx)]; }
    __device__ __forceinline__ void sync() { __syncthreads(); }
};

template <int SX> struct XCfg {
    static constexpr int XB = SX + SX / FftPlan<SX>::R0 + 1;   // padded exchange line (float2)
};

template <int SX, int MODE, bool FAST>
__global__ void __launch_bounds__(256) xpass_kernel(const __grid_constant__ XArgs a) {
    using P = FftPlan<SX>;
    constexpr int E = P::E, T = P::T, XB = XCfg<SX>::XB;
    extern __shared__ float2 smem[];
    const int jl = threadIdx.x / T, t = threadIdx.x % T;
    const long long job = (long long)blockIdx.x * a.jobsPerCta + jl;
    const long long lineA = 2 * job, lineB = 2 * job + 1;
    const bool hasA = lineA < a.nlines, hasB = lineB < a.nlines;
    float2* xb = smem + (size_t)jl * a.perJobFloat2;
    float2* stash = xb + XB;
    LineEx<SX> ex{xb};

    float2 v[E];

    // ------------------------------------------------ inverse part (C2R of every input)
    if (MODE != X_R2C_ONLY) {
        for (int g = 0; g < a.nIn; ++g) {
            const float2* pa = a.in[g] + lineA * a.pitch;
            const float2* pb = a.in[g] + lineB * a.pitch;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int idx = t + T * e;
                const int k = idx <= SX / 2 ? idx : SX - idx;
                const bool live = k <= a.kmax[g];
                float2 A = (hasA && live) ? __ldg(pa + k) : make_float2(0.0f, 0.0f);
                float2 B = (hasB && live) ? __ldg(pb + k) : make_float2(0.0f, 0.0f);
                if (k == 0 || 2 * k == SX) { A.y = 0.0f; B.y = 0.0f; }   // real-part projection of self-conjugate bins
                if (idx > SX / 2) { A.y = -A.y; B.y = -B.y; }
                v[e] = make_float2(A.x - B.y, A.y + B.x);
            }
            if (g > 0 && P::R1 > 1) __syncthreads();
            fft_line<SX, +1>(v, t, a.tw, ex);
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cscale(v[e], a.norm);
            if (MODE == X_C2R_ONLY) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int x = t + T * e;
                    if (hasA) a.realOut[lineA * SX + x] = v[e].x;
                    if (hasB) a.realOut[lineB * SX + x] = v[e].y;
                }
            } else if (!FAST) {
#pragma unroll
                for (int e = 0; e < E; ++e) stash[(size_t)g * SX + t + T * e] = v[e];
            }
        }
        if (MODE == X_C2R_ONLY) return;
    }

    // ------------------------------------------------ products + forward part (R2C of every output)
    const int nOut = MODE == X_R2C_ONLY ? 1 : a.nOut;
    for (int o = 0; o < nOut; ++o) {
        if (MODE == X_R2C_ONLY) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int x = t + T * e;
                v[e].x = hasA ? __ldg(a.realIn + lineA * SX + x) : 0.0f;
                v[e].y = hasB ? __ldg(a.realIn + lineB * SX + x) : 0.0f;
            }
        } else if (FAST) {
            // one input, one output, at most two monomials c*r^p with p <= 4: branch-free, straight from registers.
            // r^p is the left-to-right product ((r*r)*r)*r of computeProduct (src/term.cpp:85-92).
            const float c0 = a.mono[0].coef, c1 = a.nMono > 1 ? a.mono[1].coef : 0.0f;
            const int p0 = a.mono[0].nfac, p1 = a.nMono > 1 ? a.mono[1].nfac : 0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const float2 r = v[e];
                const float2 r2 = make_float2(r.x * r.x, r.y * r.y);
                const float2 r3 = make_float2(r2.x * r.x, r2.y * r.y);
                const float2 r4 = make_float2(r3.x * r.x, r3.y * r.y);
#define CUPSS_RP(p, c) ((p) == 0 ? 1.0f : ((p) == 1 ? r.c : ((p) == 2 ? r2.c : ((p) == 3 ? r3.c : r4.c))))
                float2 acc = make_float2(c0 * CUPSS_RP(p0, x), c0 * CUPSS_RP(p0, y));
                if (a.nMono > 1) { acc.x += c1 * CUPSS_RP(p1, x); acc.y += c1 * CUPSS_RP(p1, y); }
#undef CUPSS_RP
                v[e] = acc;
            }
        } else {
            // real fields of every input are in the stash, point-aligned with this thread's registers
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int x = t + T * e;
                float2 acc = make_float2(0.0f, 0.0f);
                for (int m = 0; m < a.nMono; ++m) {
                    if (a.mono[m].out != o) continue;
                    float px = a.mono[m].coef, py = px;
                    for (int f = 0; f < a.mono[m].nfac; ++f) {
                        const float2 r = stash[(size_t)a.mono[m].fac[f] * SX + x];
                        px *= r.x; py *= r.y;
                    }
                    acc.x += px; acc.y += py;
                }
                v[e] = acc;
            }
        }
        if (P::R1 > 1) __syncthreads();   // exchange buffer free (previous transform / previous untangle reads)
        fft_line<SX, -1>(v, t, a.tw, ex);
        // untangle the two real lines: needs C[sx-k], owned by another thread
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; ++e) ex.st(t + T * e, v[e]);
        __syncthreads();
        float2* qa = a.out[o] + lineA * a.pitch;
        float2* qb = a.out[o] + lineB * a.pitch;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int idx = t + T * e;
            if (idx <= SX / 2) {
                const float2 Cm = ex.ld((SX - idx) % SX);
                const float2 Ck = v[e];
                const float2 S = make_float2(Ck.x + Cm.x, Ck.y - Cm.y);   // C + conj(Cm)
                const float2 D = make_float2(Ck.x - Cm.x, Ck.y + Cm.y);   // C - conj(Cm)
                if (hasA) qa[idx] = make_float2(0.5f * S.x, 0.5f * S.y);
                if (hasB) qb[idx] = make_float2(0.5f * D.y, -0.5f * D.x);
            }
        }
    }
}

// ---------------------------------------------------------------- dispatch
template <int SX, int MODE, bool FAST>
static cudaError_t launch_x(XArgs& a, cudaStream_t st) {
    using P = FftPlan<SX>;
    constexpr int T = P::T, XB = XCfg<SX>::XB;
    const int stashLines = (MODE == X_HOT && !FAST) ? a.nIn : 0;
    const int perJob = XB + stashLines * SX;
    const size_t budget = 96 * 1024;
    int jobs = 256 / T;
    if (jobs < 1) jobs = 1;
    while (jobs > 1 && (size_t)jobs * perJob * sizeof(float2) > budget) jobs >>= 1;
    const size_t smem = (size_t)jobs * perJob * sizeof(float2);
    if (smem > 200 * 1024 || jobs * T > 256) return cudaErrorInvalidValue;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(xpass_kernel<SX, MODE, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    a.jobsPerCta = jobs;
    a.perJobFloat2 = perJob;
    const long long njobs = (a.nlines + 1) / 2;
    const unsigned grid = (unsigned)((njobs + jobs - 1) / jobs);
    xpass_kernel<SX, MODE, FAST><<<grid, jobs * T, smem, st>>>(a);
    return cudaGetLastError();
}

template <int SX>
static cudaError_t launch_x_mode(int mode, XArgs& a, cudaStream_t st) {
    if constexpr (FftPlan<SX>::T > 256) {
        return cudaErrorInvalidValue;
    } else {
        if (mode == X_C2R_ONLY) return launch_x<SX, X_C2R_ONLY, true>(a, st);
        if (mode == X_R2C_ONLY) return launch_x<SX, X_R2C_ONLY, true>(a, st);
        bool fast = a.nIn == 1 && a.nOut == 1 && a.nMono >= 1 && a.nMono <= 2;
        for (int m = 0; m < a.nMono && fast; ++m) fast = a.mono[m].nfac <= 4;
        if (fast) return launch_x<SX, X_HOT, true>(a, st);
        return launch_x<SX, X_HOT, false>(a, st);
    }
}

cudaError_t launch_xpass(int sx, int mode, XArgs& a, cudaStream_t st) {
    switch (sx) {
#define X(N) case N: return launch_x_mode<N>(mode, a, st);
        X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192)
#undef X
    }
    return cudaErrorInvalidValue;
}

}  // namespace cupss
