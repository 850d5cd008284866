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
// A "slot" is two jobs whose complex lines are interleaved element-wise in shared memory (float4 = the same
// point of both jobs): every shared access is 128 bits wide and every twiddle is loaded once for two lines.
// The transform runs in place on the slot's line, level by level (fft_core.cuh):
//   inverse as decimation in frequency (natural order in, digit-reversed out), the pointwise real-space stage
//   on the digit-reversed order (it does not care), forward as decimation in time (digit-reversed in, natural
//   out); the innermost level of both and the products are fused on registers.
#include "kernels.h"

namespace cupss {

template <int SX> struct XCfg {
    static constexpr int XB = SX + SX / 8 + 1;                      // padded line (float4 elements)
    static constexpr int NVMAX = SX / FftLevels<SX>::min_rad();     // most virtual threads per line any level has
    static constexpr int TWF4 = (TwTable<SX>::LEN + 1) / 2;         // level twiddle table, in float4 units
};
// one pad element per 8: the innermost level (stride 8 elements between lanes) and the outer levels are conflict-free
__device__ __forceinline__ unsigned xpad(unsigned idx) { return idx + (idx >> 3); }

constexpr int X_THREADS = 256;

// SLOTS > 0: slots per CTA known at compile time (no stash: the line buffer is the whole slot); 0: run-time (XArgs).
// SUMPOW (with FAST): several inputs, every monomial a power of one input -- accumulated on registers input by input.
template <int SX, int MODE, bool FAST, int SLOTS, bool SUMPOW = false>
__global__ void __launch_bounds__(X_THREADS, (FAST && !SUMPOW && FftLevels<SX>::max_rad() <= 8) ? 3 : 2) xpass_kernel(const __grid_constant__ XArgs a) {
    using F = FftLevels<SX>;
    constexpr int n = F::n, LAST = n - 1, XB = XCfg<SX>::XB;
    using GL = LevelGeom<SX, LAST>;
    constexpr unsigned RL = GL::R;
    extern __shared__ float4 smem4[];
    float2* twS = reinterpret_cast<float2*>(smem4);
    float4* bufs = smem4 + XCfg<SX>::TWF4;
    const unsigned S = SLOTS > 0 ? (unsigned)SLOTS : (unsigned)a.jobsPerCta;            // slots per CTA
    const unsigned per = SLOTS > 0 ? (unsigned)XB : (unsigned)a.perJobFloat2;           // float4 elements per slot (line + stash)
    const unsigned tid = threadIdx.x;
    const unsigned NT = SLOTS > 0 ? (unsigned)X_THREADS : blockDim.x;   // the stash path sizes its CTAs to the slots that fit

    for (unsigned i = tid; i < (unsigned)TwTable<SX>::LEN; i += NT) twS[i] = __ldg(a.tw + i);
    __syncthreads();

    using std::integral_constant;
    using Plus = integral_constant<int, 1>;
    using Minus = integral_constant<int, -1>;
    using Dif = integral_constant<bool, true>;
    using Dit = integral_constant<bool, false>;

    // lines of slot s: job0 = (A0, B0), job1 = (A1, B1)
    auto line0 = [&](unsigned s) -> long long { return 4ll * ((long long)blockIdx.x * S + s); };

    // Source lines of slot s for input g: pointers to the four half-spectrum lines (A0, B0, A1, B1) and which of them exist.
    struct SlotSrc { const float2* p[4]; bool v[4]; int kmax; };
    auto slot_src = [&](int g, unsigned s) -> SlotSrc {
        SlotSrc r;
        const long long l0 = line0(s);
        const float2* base = a.in[g] + l0 * a.pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) { r.v[i] = l0 + i < a.nlines; r.p[i] = base + (unsigned)(i * a.pitch); }
        r.kmax = a.kmax[g];
        return r;
    };
    // C[idx] of both jobs (float4: job0 in xy, job1 in zw).
    //   HALF = 0: idx <= sx/2 known (k = idx, no conjugation); 1: idx > sx/2 known (k = sx - idx, conjugated); 2: decide at run time.
    auto formC = [&](auto halfTag, const SlotSrc& src, unsigned idx) -> float4 {
        constexpr int HALF = decltype(halfTag)::value;
        const bool upper = HALF == 1 || (HALF == 2 && idx > SX / 2);
        const unsigned k = upper ? SX - idx : idx;
        const bool live = (int)k <= src.kmax;
        const float2 z = make_float2(0.0f, 0.0f);
        float2 A0 = (live && src.v[0]) ? __ldg(src.p[0] + k) : z;
        float2 B0 = (live && src.v[1]) ? __ldg(src.p[1] + k) : z;
        float2 A1 = (live && src.v[2]) ? __ldg(src.p[2] + k) : z;
        float2 B1 = (live && src.v[3]) ? __ldg(src.p[3] + k) : z;
        if (k == 0 || 2 * k == SX) { A0.y = 0.0f; B0.y = 0.0f; A1.y = 0.0f; B1.y = 0.0f; }   // real-part projection of self-conjugate bins
        if (upper) return make_float4(A0.x + B0.y, B0.x - A0.y, A1.x + B1.y, B1.x - A1.y);   // conj(A) + i conj(B)
        return make_float4(A0.x - B0.y, A0.y + B0.x, A1.x - B1.y, A1.y + B1.x);                // A + i B
    };

    // one in-place level over every slot of the CTA, shared -> shared
    auto level_ss = [&](auto lvTag, auto signTag, auto difTag) {
        constexpr int LV = decltype(lvTag)::value;
        constexpr int SIGN = decltype(signTag)::value;
        constexpr bool DIF = decltype(difTag)::value;
        using G = LevelGeom<SX, LV>;
        constexpr unsigned R = G::R, M = G::M, N = G::N, NV = G::NV;
#pragma unroll 1
        for (unsigned w = tid; w < S * NV; w += NT) {
            const unsigned s = w / NV, v = w % NV;
            const unsigned blk = v / M, j = v % M, row0 = blk * N + j;
            float4* xb = bufs + s * per;
            float2 x0[R], x1[R];
#pragma unroll
            for (unsigned q = 0; q < R; ++q) {
                const float4 t = xb[xpad(row0 + M * q)];
                x0[q] = make_float2(t.x, t.y); x1[q] = make_float2(t.z, t.w);
            }
            level_butterfly2<SX, LV, SIGN, DIF>(x0, x1, j, twS);
#pragma unroll
            for (unsigned q = 0; q < R; ++q) xb[xpad(row0 + M * q)] = make_float4(x0[q].x, x0[q].y, x1[q].x, x1[q].y);
        }
    };

    // ------------------------------------------------ R2C only (upload): real lines in, forward DIF, untangle from the digit-reversed order
    if constexpr (MODE == X_R2C_ONLY) {
        {
            using G = LevelGeom<SX, 0>;
            constexpr unsigned R = G::R, M = G::M, NV = G::NV;
#pragma unroll 1
            for (unsigned w = tid; w < S * NV; w += NT) {
                const unsigned s = w / NV, v = w % NV;
                const long long l0 = line0(s);
                float4* xb = bufs + s * per;
                float2 x0[R], x1[R];
#pragma unroll
                for (unsigned q = 0; q < R; ++q) {
                    const unsigned x = v + M * q;
                    // a.norm: 1 for a plain forward transform; the scale that turns real values back into the engine's
                    // un-normalised W2 convention when a real view is committed (a power of two: exact)
                    x0[q].x = l0 < a.nlines ? __ldg(a.realIn + l0 * SX + x) * a.norm : 0.0f;
                    x0[q].y = l0 + 1 < a.nlines ? __ldg(a.realIn + (l0 + 1) * SX + x) * a.norm : 0.0f;
                    x1[q].x = l0 + 2 < a.nlines ? __ldg(a.realIn + (l0 + 2) * SX + x) * a.norm : 0.0f;
                    x1[q].y = l0 + 3 < a.nlines ? __ldg(a.realIn + (l0 + 3) * SX + x) * a.norm : 0.0f;
                }
                level_butterfly2<SX, 0, -1, true>(x0, x1, v, twS);
#pragma unroll
                for (unsigned q = 0; q < R; ++q) xb[xpad(v + M * q)] = make_float4(x0[q].x, x0[q].y, x1[q].x, x1[q].y);
            }
        }
        __syncthreads();
        if constexpr (n >= 2) { level_ss(integral_constant<int, 1>{}, Minus{}, Dif{}); __syncthreads(); }
        if constexpr (n >= 3) { level_ss(integral_constant<int, 2>{}, Minus{}, Dif{}); __syncthreads(); }
        if constexpr (n >= 4) { level_ss(integral_constant<int, 3>{}, Minus{}, Dif{}); __syncthreads(); }
    }

    // ------------------------------------------------ inverse part (C2R of every input), decimation in frequency
    // FAST with several inputs ("sum of powers": every monomial is a power of ONE input, one output): the contributions
    // are accumulated on registers input by input, no stash.  Every thread runs the innermost loop at most once (the slot
    // count is chosen so), so the accumulators are plain registers across the inputs.
    float2 acc0[SUMPOW ? RL : 1], acc1[SUMPOW ? RL : 1];
    if constexpr (SUMPOW) {
#pragma unroll
        for (unsigned q = 0; q < RL; ++q) { acc0[q] = make_float2(0.0f, 0.0f); acc1[q] = make_float2(0.0f, 0.0f); }
    }
    if constexpr (MODE != X_R2C_ONLY) {
        for (int g = 0; g < a.nIn; ++g) {
            if (g > 0) __syncthreads();   // line buffers are re-used per input
            if constexpr (n > 1) {
                using G = LevelGeom<SX, 0>;
                constexpr unsigned R = G::R, M = G::M, NV = G::NV;
#pragma unroll 1
                for (unsigned w = tid; w < S * NV; w += NT) {
                    const unsigned s = w / NV, v = w % NV;
                    float4* xb = bufs + s * per;
                    const SlotSrc src = slot_src(g, s);
                    float2 x0[R], x1[R];
#pragma unroll
                    for (unsigned q = 0; q < R; ++q) {
                        // v < M here, so which half of the spectrum the point lies in is known per q except on the q*M == sx/2 row
                        float4 t;
                        if (M * q < SX / 2u) t = formC(integral_constant<int, 0>{}, src, v + M * q);
                        else if (M * q > SX / 2u) t = formC(integral_constant<int, 1>{}, src, v + M * q);
                        else t = formC(integral_constant<int, 2>{}, src, v + M * q);
                        x0[q] = make_float2(t.x, t.y); x1[q] = make_float2(t.z, t.w);
                    }
                    level_butterfly2<SX, 0, +1, true>(x0, x1, v, twS);
#pragma unroll
                    for (unsigned q = 0; q < R; ++q) xb[xpad(v + M * q)] = make_float4(x0[q].x, x0[q].y, x1[q].x, x1[q].y);
                }
                __syncthreads();
                if constexpr (n >= 3) { level_ss(integral_constant<int, 1>{}, Plus{}, Dif{}); __syncthreads(); }
                if constexpr (n >= 4) { level_ss(integral_constant<int, 2>{}, Plus{}, Dif{}); __syncthreads(); }
            }
            // innermost level: real values appear on registers (position p <-> real index freq_of_pos(p))
#pragma unroll 1
            for (unsigned w = tid; w < S * GL::NV; w += NT) {
                const unsigned s = w / GL::NV, v = w % GL::NV;
                float4* xb = bufs + s * per;
                SlotSrc src;
                if constexpr (n == 1) src = slot_src(g, s);
                float2 x0[RL], x1[RL];
#pragma unroll
                for (unsigned q = 0; q < RL; ++q) {
                    float4 t;
                    if constexpr (n > 1) t = xb[xpad(v * RL + q)];
                    else t = formC(integral_constant<int, 2>{}, src, q);
                    x0[q] = make_float2(t.x, t.y); x1[q] = make_float2(t.z, t.w);
                }
                level_butterfly2<SX, LAST, +1, true>(x0, x1, 0, twS);
#pragma unroll
                for (unsigned q = 0; q < RL; ++q) { x0[q] = cscale(x0[q], a.norm); x1[q] = cscale(x1[q], a.norm); }
                if constexpr (MODE == X_C2R_ONLY) {
                    // park the real values in place; written out in natural order below
#pragma unroll
                    for (unsigned q = 0; q < RL; ++q) xb[xpad(v * RL + q)] = make_float4(x0[q].x, x0[q].y, x1[q].x, x1[q].y);
                } else if constexpr (FAST) {
                    // one input, one output, at most two monomials c*r^p with p <= 4: branch-free, straight from registers.
                    // r^p is the left-to-right product ((r*r)*r)*r of computeProduct (src/term.cpp:85-92).
                    const float c0 = a.mono[0].coef, c1 = a.nMono > 1 ? a.mono[1].coef : 0.0f;
                    const int p0 = a.mono[0].nfac, p1 = a.nMono > 1 ? a.mono[1].nfac : 0;
                    auto powr = [](float r, int p) -> float {
                        const float r2 = r * r, r3 = r2 * r, r4 = r3 * r;
                        return p == 0 ? 1.0f : (p == 1 ? r : (p == 2 ? r2 : (p == 3 ? r3 : r4)));
                    };
                    auto apply = [&](auto f) {
#pragma unroll
                        for (unsigned q = 0; q < RL; ++q) {
                            x0[q] = make_float2(f(x0[q].x), f(x0[q].y));
                            x1[q] = make_float2(f(x1[q].x), f(x1[q].y));
                        }
                    };
                    // warp-uniform dispatch on the monomial pattern: the common single-monomial powers are straight-line code
                    if constexpr (!SUMPOW) {
                        if (a.nMono == 1 && p0 == 3) apply([&](float r) { return c0 * ((r * r) * r); });
                        else if (a.nMono == 1 && p0 == 2) apply([&](float r) { return c0 * (r * r); });
                        else if (a.nMono == 1) apply([&](float r) { return c0 * powr(r, p0); });
                        else apply([&](float r) { return c0 * powr(r, p0) + c1 * powr(r, p1); });
                    } else {
                        // sum of powers: add this input's monomials (in monomial order, like the stash path) to the accumulators
                        for (int m = 0; m < a.nMono; ++m) {
                            if (a.mono[m].fac[0] != g) continue;
                            const float cm = a.mono[m].coef;
                            const int pm = a.mono[m].nfac;
#pragma unroll
                            for (unsigned q = 0; q < RL; ++q) {
                                acc0[q].x += cm * powr(x0[q].x, pm); acc0[q].y += cm * powr(x0[q].y, pm);
                                acc1[q].x += cm * powr(x1[q].x, pm); acc1[q].y += cm * powr(x1[q].y, pm);
                            }
                        }
                        if (g + 1 < a.nIn) continue;   // more inputs to come: nothing to transform yet
#pragma unroll
                        for (unsigned q = 0; q < RL; ++q) { x0[q] = acc0[q]; x1[q] = acc1[q]; }
                    }
                    level_butterfly2<SX, LAST, -1, false>(x0, x1, 0, twS);
#pragma unroll
                    for (unsigned q = 0; q < RL; ++q) xb[xpad(v * RL + q)] = make_float4(x0[q].x, x0[q].y, x1[q].x, x1[q].y);
                } else {
                    float4* stash = xb + XB + (unsigned)g * XB;
#pragma unroll
                    for (unsigned q = 0; q < RL; ++q) stash[xpad(v * RL + q)] = make_float4(x0[q].x, x0[q].y, x1[q].x, x1[q].y);
                }
            }
        }
    }

    if constexpr (MODE == X_C2R_ONLY) {
        __syncthreads();
        const unsigned tps = NT / S;   // threads per slot (S divides X_THREADS)
        const unsigned s = tid / tps;
        const long long l0 = line0(s);
        const float4* xb = bufs + s * per;
        for (unsigned x = tid % tps; x < (unsigned)SX; x += tps) {
            const float4 t = xb[xpad(pos_of_freq<SX>(x))];
            if (l0 < a.nlines) a.realOut[l0 * SX + x] = t.x;
            if (l0 + 1 < a.nlines) a.realOut[(l0 + 1) * SX + x] = t.y;
            if (l0 + 2 < a.nlines) a.realOut[(l0 + 2) * SX + x] = t.z;
            if (l0 + 3 < a.nlines) a.realOut[(l0 + 3) * SX + x] = t.w;
        }
        return;
    }

    // ------------------------------------------------ products + forward part (R2C of every output), decimation in time
    const int nOut = (MODE == X_R2C_ONLY || FAST) ? 1 : a.nOut;
    for (int o = 0; o < nOut; ++o) {
        if constexpr (MODE == X_HOT && !FAST) {
            // real fields of every input are in the stash, position-aligned with this thread's registers
            __syncthreads();   // stash complete / previous output's untangle reads done
#pragma unroll 1
            for (unsigned w = tid; w < S * GL::NV; w += NT) {
                const unsigned s = w / GL::NV, v = w % GL::NV;
                float4* xb = bufs + s * per;
                const float4* stash = xb + XB;
                float2 x0[RL], x1[RL];
#pragma unroll
                for (unsigned q = 0; q < RL; ++q) {
                    const unsigned p = xpad(v * RL + q);
                    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    for (int m = 0; m < a.nMono; ++m) {
                        if (a.mono[m].out != o) continue;
                        float4 pr = make_float4(a.mono[m].coef, a.mono[m].coef, a.mono[m].coef, a.mono[m].coef);
                        for (int f = 0; f < a.mono[m].nfac; ++f) {
                            const float4 r = stash[(unsigned)a.mono[m].fac[f] * XB + p];
                            pr.x *= r.x; pr.y *= r.y; pr.z *= r.z; pr.w *= r.w;
                        }
                        acc.x += pr.x; acc.y += pr.y; acc.z += pr.z; acc.w += pr.w;
                    }
                    x0[q] = make_float2(acc.x, acc.y); x1[q] = make_float2(acc.z, acc.w);
                }
                level_butterfly2<SX, LAST, -1, false>(x0, x1, 0, twS);
#pragma unroll
                for (unsigned q = 0; q < RL; ++q) xb[xpad(v * RL + q)] = make_float4(x0[q].x, x0[q].y, x1[q].x, x1[q].y);
            }
        }
        if constexpr (MODE == X_HOT) {
            __syncthreads();
            if constexpr (n >= 4) { level_ss(integral_constant<int, 2>{}, Minus{}, Dit{}); __syncthreads(); }
            if constexpr (n >= 3) { level_ss(integral_constant<int, 1>{}, Minus{}, Dit{}); __syncthreads(); }
            if constexpr (n >= 2) { level_ss(integral_constant<int, 0>{}, Minus{}, Dit{}); __syncthreads(); }
        }
        // untangle the two real lines of each job: needs C[k] and C[sx-k]; two neighbouring k per thread (128-bit stores)
        {
            const unsigned tps = NT / S;
            const unsigned s = tid / tps;
            const long long l0 = line0(s);
            const float4* xb = bufs + s * per;
            float2* q0 = a.out[o] + l0 * a.pitch;
            auto at = [&](unsigned k) -> float4 {   // C[k mod SX] of both jobs
                const unsigned kk = k & (SX - 1);
                return xb[xpad(MODE == X_R2C_ONLY ? pos_of_freq<SX>(kk) : kk)];
            };
            auto split = [&](float4 Ck, float4 Cm, float2& A0, float2& B0, float2& A1, float2& B1) {
                A0 = make_float2(0.5f * (Ck.x + Cm.x), 0.5f * (Ck.y - Cm.y));     // (C + conj Cm) / 2
                B0 = make_float2(0.5f * (Ck.y + Cm.y), -0.5f * (Ck.x - Cm.x));    // (C - conj Cm) / (2i)
                A1 = make_float2(0.5f * (Ck.z + Cm.z), 0.5f * (Ck.w - Cm.w));
                B1 = make_float2(0.5f * (Ck.w + Cm.w), -0.5f * (Ck.z - Cm.z));
            };
            for (unsigned k = 2 * (tid % tps); k <= SX / 2; k += 2 * tps) {
                float2 A0, B0, A1, B1, A0n, B0n, A1n, B1n;
                split(at(k), at(SX - k), A0, B0, A1, B1);
                if (k + 1 <= SX / 2) {
                    split(at(k + 1), at(SX - k - 1), A0n, B0n, A1n, B1n);
                    if (l0 < a.nlines) *reinterpret_cast<float4*>(q0 + k) = make_float4(A0.x, A0.y, A0n.x, A0n.y);
                    if (l0 + 1 < a.nlines) *reinterpret_cast<float4*>(q0 + a.pitch + k) = make_float4(B0.x, B0.y, B0n.x, B0n.y);
                    if (l0 + 2 < a.nlines) *reinterpret_cast<float4*>(q0 + 2 * a.pitch + k) = make_float4(A1.x, A1.y, A1n.x, A1n.y);
                    if (l0 + 3 < a.nlines) *reinterpret_cast<float4*>(q0 + 3 * a.pitch + k) = make_float4(B1.x, B1.y, B1n.x, B1n.y);
                } else {
                    if (l0 < a.nlines) q0[k] = A0;
                    if (l0 + 1 < a.nlines) q0[a.pitch + k] = B0;
                    if (l0 + 2 < a.nlines) q0[2 * a.pitch + k] = A1;
                    if (l0 + 3 < a.nlines) q0[3 * a.pitch + k] = B1;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- dispatch
template <int SX> struct XSlots {   // one virtual thread per real thread on the widest level
    static constexpr int RAW = X_THREADS / XCfg<SX>::NVMAX;
    static constexpr int V = RAW < 1 ? 1 : (RAW > 64 ? 64 : RAW);
};

template <int SX, int MODE, bool FAST, bool SUMPOW = false>
static cudaError_t launch_x(XArgs& a, cudaStream_t st) {
    constexpr int XB = XCfg<SX>::XB;
    constexpr bool STASH = (MODE == X_HOT && !FAST);
    constexpr int CT = STASH ? 0 : XSlots<SX>::V;
    const int stashLines = STASH ? a.nIn : 0;
    const int perSlot = XB * (1 + stashLines);                      // float4 elements
    const size_t budget = 100 * 1024;   // two CTAs per SM
    int slots = XSlots<SX>::V;
    int threads = X_THREADS;
    if (CT == 0) {   // run-time slot count (stash path): shrink to the shared-memory budget and size the CTA to it
        while (slots > 1 && (size_t)slots * perSlot * sizeof(float4) > budget) slots >>= 1;
        threads = slots * XCfg<SX>::NVMAX;
        if (threads > X_THREADS) threads = X_THREADS;
        if (threads < 32) threads = 32;
        while (threads % slots) ++threads;   // the untangle splits the CTA evenly over the slots
    }
    const size_t smem = ((size_t)XCfg<SX>::TWF4 + (size_t)slots * perSlot) * sizeof(float4);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(xpass_kernel<SX, MODE, FAST, CT, SUMPOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    a.jobsPerCta = slots;
    a.perJobFloat2 = perSlot;
    const long long nslots = (a.nlines + 3) / 4;
    const unsigned grid = (unsigned)((nslots + slots - 1) / slots);
    xpass_kernel<SX, MODE, FAST, CT, SUMPOW><<<grid, threads, smem, st>>>(a);
    return cudaGetLastError();
}

template <int SX>
static cudaError_t launch_x_mode(int mode, XArgs& a, cudaStream_t st) {
    if (mode == X_C2R_ONLY) return launch_x<SX, X_C2R_ONLY, true>(a, st);
    if (mode == X_R2C_ONLY) return launch_x<SX, X_R2C_ONLY, true>(a, st);
    // FAST: one output whose monomials are powers (<= 4) of a single input each -- one input with up to two monomials, or
    // several inputs ("sum of powers", e.g. the three squared gradients of KPZ) accumulated on registers
    bool fast = a.nOut == 1 && a.nMono >= 1 && (a.nIn > 1 || a.nMono <= 2);
    for (int m = 0; m < a.nMono && fast; ++m) {
        fast = a.mono[m].nfac >= 1 && a.mono[m].nfac <= 4;
        for (int f = 1; f < a.mono[m].nfac && fast; ++f) fast = a.mono[m].fac[f] == a.mono[m].fac[0];
    }
    if (fast && a.tw3 && xpass3_supported(SX) && a.nIn == 1 && a.nMono <= 2) return launch_xpass3(SX, a, st);   // powers 1..4 of the one input
    if (fast && a.nIn > 1 && a.tw3 && xpass3_sumpow_supported(SX) && a.nMono == a.nIn) {
        bool oneEach = true;   // monomial g is a power of input g: the accumulation order of the generic kernels is the input order
        for (int m = 0; m < a.nMono; ++m) oneEach = oneEach && a.mono[m].fac[0] == m;
        if (oneEach) return launch_xpass3_sumpow(SX, a, st);
    }
    if (fast && a.nIn > 1) return launch_x<SX, X_HOT, true, true>(a, st);
    if (fast) return launch_x<SX, X_HOT, true>(a, st);
    if (xstash1_supported(SX)) return launch_xstash1(SX, a, st);   // long lines: one job per CTA (kernels_xs.cu)
    return launch_x<SX, X_HOT, false>(a, st);
}

// Most inputs one x-pass launch can take for this line length: the real fields of all inputs of a slot are stashed in
// shared memory next to the line buffer (one slot must fit in 200 KB).
int xpass_max_inputs(int sx) {
    if (xstash1_supported(sx)) return xstash1_max_inputs(sx);
    const long long xb = (long long)sx + sx / 8 + 1;   // XCfg<SX>::XB
    const long long fit = (200ll * 1024 - ((long long)sx * 8)) / (xb * (long long)sizeof(float4)) - 1;
    return (int)(fit < 1 ? 1 : (fit > XP_MAX_IN ? XP_MAX_IN : fit));
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
