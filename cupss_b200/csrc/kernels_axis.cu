// kernels_axis.cu -- strided-axis FFT passes (y and z) of the half-spectrum, plain and fused.
//
// A CTA owns a tile [L rows] x [C columns] of one k-space array: the L points of C neighbouring
// kx columns along the transformed axis.  Lanes map to columns, so every global access of a
// half-warp is one contiguous, 128-byte-aligned row segment (C = 16 float2) and every
// shared-memory access is conflict-free.  Thread (t, c) owns points t + T*e of column c.
//
//   axis_plain_kernel  : load -> FFT -> store                      (y passes of a 3-D transform;
//                        optional dealias mask on load for extra inverse transforms)
//   axis_kstage_kernel : [load -> forward FFT] -> per-mode update of every field of the sweep
//                        (kstage_point) -> [dealias -> inverse FFT -> store]
// The second is the heart of the step: the last pass of the forward transform of the nonlinear
// term, the semi-implicit Euler update, the dealiasing mask and the first pass of the next inverse
// transform happen on data that never leaves the SM.
#include "kernels.h"

namespace cupss {

template <int C, int PADR>
struct TileEx {
    float2* buf;
    int c;
    __device__ __forceinline__ static int prow(int idx) { return PADR > 0 ? idx + idx / PADR : idx; }
    __device__ __forceinline__ void st(int idx, float2 v) { buf[prow(idx) * C + c] = v; }
    __device__ __forceinline__ float2 ld(int idx) const { return buf[prow(idx) * C + c]; }
    __device__ __forceinline__ void sync() { __syncthreads(); }
};

template <int L> struct AxisCfg {
    static constexpr int C = L <= 512 ? 16 : (L <= 2048 ? 8 : (L <= 4096 ? 4 : 2));
    static constexpr int PADR = C < 16 ? FftPlan<L>::R0 : 0;
    static constexpr int ROWS = PADR > 0 ? L + L / PADR + 1 : L;
    static constexpr int THREADS = FftPlan<L>::T * C;
    static constexpr int MINB = THREADS >= 512 ? 1 : 512 / THREADS;   // cap registers at 128/thread: >= 16 warps per SM
    static constexpr size_t SMEM = FftPlan<L>::R1 > 1 ? (size_t)ROWS * C * sizeof(float2) : 0;
};

__device__ __forceinline__ long long axis_off(const AxisAddr& a, int b, int row, int col) {
    return (long long)b * a.bs + (long long)(row >> a.rpcShift) * a.cs + (long long)(row & a.rpcMask) * a.rs + col;
}

// Destination of row `row` of batch b: local array, or the receive buffer of the peer that owns the row.
__device__ __forceinline__ float2* axis_dst(const AxisArgs& a, int b, int row, int col) {
    if (a.pushOn)
        return a.push[row >> a.pushShift] + a.pushBase + (long long)b * a.pushBs + (long long)(row & a.pushMask) * a.pushRs + col;
    return a.out + axis_off(a.aout, b, row, col);
}

template <int L, int DIR>
__global__ void __launch_bounds__(AxisCfg<L>::THREADS, AxisCfg<L>::MINB) axis_plain_kernel(const __grid_constant__ AxisArgs a) {
    using P = FftPlan<L>;
    constexpr int E = P::E, T = P::T, C = AxisCfg<L>::C;
    extern __shared__ float2 smem[];
    const int c = threadIdx.x % C, t = threadIdx.x / C;
    const int ct = blockIdx.x % a.ncolTiles, b = blockIdx.x / a.ncolTiles;
    const int col = ct * C + c;
    const bool valid = col < a.ncol;
    TileEx<C, AxisCfg<L>::PADR> ex{smem, c};

    if (a.pruneOn) {   // CTA-uniform: the whole tile is outside the dealias cut-off -> output stays zero
        const int iyT = a.kyBase + b;
        const int nyT = iyT > a.sy / 2 ? a.sy - iyT : iyT;
        if (ct * C > a.pruneCutX || (a.axis == 2 && nyT > a.pruneCutY)) return;
    }

    float2 v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int row = t + T * e;
        bool keep = valid;
        if (a.rowCut >= 0) {
            const int nr = row > L / 2 ? L - row : row;
            keep = keep && nr <= a.rowCut;
        }
        if (a.maskOn) {
            const int iy = a.axis == 2 ? a.kyBase + b : (a.axis == 1 ? row : 0);
            const int iz = a.axis == 2 ? row : 0;
            keep = keep && dealias_keep(col, iy, iz, a.sx, a.sy, a.sz, a.cutx, a.cuty, a.cutz);
        }
        v[e] = keep ? __ldg(a.in + axis_off(a.ain, b, row, col)) : make_float2(0.0f, 0.0f);
    }
    fft_line<L, DIR>(v, t, a.tw, ex);
    if (valid) {
        if (a.pushOn) {
#pragma unroll
            for (int e = 0; e < E; ++e) *axis_dst(a, b, t + T * e, col) = v[e];
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) a.out[axis_off(a.aout, b, t + T * e, col)] = v[e];
        }
    }
}

template <int L, int KIND, int SIG>
__global__ void __launch_bounds__(AxisCfg<L>::THREADS, AxisCfg<L>::MINB)
axis_kstage_kernel(const __grid_constant__ AxisArgs a, const __grid_constant__ KStageD ks) {
    using P = FftPlan<L>;
    constexpr int E = P::E, T = P::T, C = AxisCfg<L>::C;
    extern __shared__ float2 smem[];
    const int c = threadIdx.x % C, t = threadIdx.x / C;
    const int ct = blockIdx.x % a.ncolTiles, b = blockIdx.x / a.ncolTiles;
    const int col = ct * C + c;
    const bool valid = col < a.ncol;
    TileEx<C, AxisCfg<L>::PADR> ex{smem, c};

    float2 v[E];
    if (ks.hasFwd) {
#pragma unroll
        for (int e = 0; e < E; ++e)
            v[e] = valid ? __ldg(a.in + axis_off(a.ain, b, t + T * e, col)) : make_float2(0.0f, 0.0f);
        if (KIND == KS_SCALAR_Q2 && c == 0) {
            // pull this tile's rows of the state spectrum towards L2 while the forward transform runs
#pragma unroll
            for (int e = 0; e < E; ++e)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ks.src[0] + axis_off(a.aout, b, t + T * e, col)));
        }
        fft_line<L, -1>(v, t, a.tw, ex);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = make_float2(0.0f, 0.0f);
    }

    // fixed (per thread) part of the mode index: column = kx, and ky for a z pass
    const int iyFix = a.axis == 2 ? a.kyBase + b : 0;
    if (KIND == KS_SCALAR_Q2) {
        // Lean path.  The point loop is ROLLED (CH points per trip) to keep the kernel inside the instruction
        // cache; the register array is rotated by CH after every trip so that all indices stay static.
        const OutD& od = ks.out[0];
        const int sRow = a.axis == 2 ? ks.sz : (a.axis == 1 ? ks.sy : 1);
        const float stepRow = a.axis == 2 ? ks.stepqz : ks.stepqy;
        const int cutRow = a.axis == 2 ? od.cutz : od.cuty;
        const float qx = wavenumber(col, ks.sx, ks.stepqx);
        const float qyFix = wavenumber(iyFix, ks.sy, ks.stepqy);
        const float qx2 = CUPSS_FMUL(qx, qx);
        const float qyFix2 = CUPSS_FMUL(qyFix, qyFix);
        const bool fixSelf = ((col == 0) || (2 * col == ks.sx)) && ((iyFix == 0) || (2 * iyFix == ks.sy));
        const int nyFix = iyFix > ks.sy / 2 ? ks.sy - iyFix : iyFix;
        const bool keepFix = od.inv && col <= od.cutx && (a.axis != 2 || nyFix <= od.cuty);
        const long long rowStride = (long long)T * a.aout.rs;          // natural layout: consecutive e are T rows apart
        const long long off0 = axis_off(a.aout, b, t, col);
        const float2* sp = ks.src[0] + off0;
        float2* dp = ks.dst[0] + off0;
        if constexpr (SIG >= 0) {
            // exponent pattern known at compile time: short straight-line body, fully unrolled, loads batched by 8
            const double tp[3] = {ks.sq2.tpre[0], ks.sq2.tpre[1], ks.sq2.tpre[2]};
            const double ip[4] = {ks.sq2.ipre[0], ks.sq2.ipre[1], ks.sq2.ipre[2], ks.sq2.ipre[3]};
            const bool termFused = ks.sq2.termFused != 0;
            const float dt = ks.dt;
            constexpr int CH = E < 8 ? E : 8;
#pragma unroll
            for (int e0 = 0; e0 < E; e0 += CH) {
                float2 self[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) self[j] = valid ? __ldcg(sp + (e0 + j) * rowStride) : make_float2(0.0f, 0.0f);
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int e = e0 + j;
                    const int row = t + T * e;
                    const float qr = wavenumber(row, sRow, stepRow);
                    const float qr2 = CUPSS_FMUL(qr, qr);
                    const float q2 = CUPSS_FADD(CUPSS_FADD(qx2, a.axis == 2 ? qyFix2 : qr2), a.axis == 2 ? qr2 : 0.0f);
                    float2 val = kstage_point_scalar_q2_sig<SIG>(tp, ip, termFused, dt, q2, v[e], self[j]);
                    if (fixSelf && ((row == 0) || (2 * row == sRow))) val.y = 0.0f;
                    if (valid) dp[e * rowStride] = val;
                    const int nr = row > sRow / 2 ? sRow - row : row;
                    v[e] = (keepFix && nr <= cutRow) ? val : make_float2(0.0f, 0.0f);
                }
            }
        } else {
        constexpr int CH = E < 4 ? E : 4;
        int row0 = t;
#pragma unroll 1
        for (int ch = 0; ch < E / CH; ++ch) {
            float2 self[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) self[j] = valid ? __ldcg(sp + j * rowStride) : make_float2(0.0f, 0.0f);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int row = row0 + T * j;
                const float qr = wavenumber(row, sRow, stepRow);
                const float qr2 = CUPSS_FMUL(qr, qr);
                const float q2 = CUPSS_FADD(CUPSS_FADD(qx2, a.axis == 2 ? qyFix2 : qr2), a.axis == 2 ? qr2 : 0.0f);
                float2 val = kstage_point_scalar_q2(ks.sq2, ks.dt, q2, v[j], self[j]);
                if (fixSelf && ((row == 0) || (2 * row == sRow))) val.y = 0.0f;
                if (valid) dp[j * rowStride] = val;
                const int nr = row > sRow / 2 ? sRow - row : row;
                v[j] = (keepFix && nr <= cutRow) ? val : make_float2(0.0f, 0.0f);
            }
            if (E > CH) {   // rotate: v[i] <- v[i + CH]
                float2 tmp[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) tmp[j] = v[j];
#pragma unroll
                for (int i2 = 0; i2 < E - CH; ++i2) v[i2] = v[i2 + CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) v[E - CH + j] = tmp[j];
            }
            row0 += T * CH;
            sp += CH * rowStride;
            dp += CH * rowStride;
        }
        }
    } else {
        const unsigned int step = ks.stepCounter ? *ks.stepCounter : 0u;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int row = t + T * e;
            const int iy = a.axis == 2 ? iyFix : (a.axis == 1 ? row : 0);
            const int iz = a.axis == 2 ? row : 0;
            if (valid) {
                const KPoint k = make_kpoint(ks, col, iy, iz);
                // k-space arrays share the natural addressing of the pass output
                v[e] = kstage_point(ks, k, v[e], axis_off(a.aout, b, row, col), step);
            } else {
                v[e] = make_float2(0.0f, 0.0f);
            }
        }
    }

    if (ks.hasInv) {
        if (a.pruneOn) {   // CTA-uniform: every mode of this tile is masked out -> nothing to transform or store
            const int nyT = iyFix > ks.sy / 2 ? ks.sy - iyFix : iyFix;
            if (ct * C > a.pruneCutX || (a.axis == 2 && nyT > a.pruneCutY)) return;
        }
        if (ks.hasFwd && P::R1 > 1) __syncthreads();   // exchange buffer still being read by the forward transform
        fft_line<L, +1>(v, t, a.tw, ex);
        if (valid) {
            if (a.pushOn) {
#pragma unroll
                for (int e = 0; e < E; ++e) *axis_dst(a, b, t + T * e, col) = v[e];
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) a.out[axis_off(a.aout, b, t + T * e, col)] = v[e];
            }
        }
    }
}

__global__ void bump_counter_kernel(unsigned int* c) { *c += 1u; }

__global__ void xgpu_barrier_kernel(const XBarrier b) {
    __shared__ unsigned int target;
    if (threadIdx.x == 0) {
        target = b.epoch[b.pt] + 1u;
        b.epoch[b.pt] = target;
    }
    __syncthreads();
    const int d = threadIdx.x;
    if (d < b.nranks) {
        // everything this GPU pushed in earlier kernels of the stream is ordered before the flag (cumulative fence)
        __threadfence_system();
        unsigned int* theirs = b.flags[d] + b.pt * CUPSS_MAX_PEERS + b.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(target) : "memory");
        const unsigned int* mine = b.flags[b.rank] + b.pt * CUPSS_MAX_PEERS + d;
        const long long t0 = clock64();
        unsigned int seen;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
            if (clock64() - t0 > 6000000000LL) { *b.error = 1; break; }   // ~3 s: a peer died; do not hang the GPU
        } while ((int)(seen - target) < 0);
        __threadfence_system();
    }
}

// ---------------------------------------------------------------- dispatch
template <int L>
static cudaError_t launch_plain_L(int dir, const AxisArgs& a, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (AxisCfg<L>::SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(axis_plain_kernel<L, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AxisCfg<L>::SMEM);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(axis_plain_kernel<L, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AxisCfg<L>::SMEM);
            if (e != cudaSuccess) return e;
        }
        attr = true;
    }
    const unsigned grid = (unsigned)a.ncolTiles * (unsigned)a.nbatch;
    if (dir < 0) axis_plain_kernel<L, -1><<<grid, AxisCfg<L>::THREADS, AxisCfg<L>::SMEM, st>>>(a);
    else axis_plain_kernel<L, 1><<<grid, AxisCfg<L>::THREADS, AxisCfg<L>::SMEM, st>>>(a);
    return cudaGetLastError();
}

template <int L>
static cudaError_t launch_kstage_L(const AxisArgs& a, const KStageD& ks, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (AxisCfg<L>::SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(axis_kstage_kernel<L, KS_GENERIC, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AxisCfg<L>::SMEM);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(axis_kstage_kernel<L, KS_SCALAR_Q2, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AxisCfg<L>::SMEM);
            if (e != cudaSuccess) return e;
            if constexpr (L >= 64) {
                e = cudaFuncSetAttribute(axis_kstage_kernel<L, KS_SCALAR_Q2, SQ2_SIG_CAHN_HILLIARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AxisCfg<L>::SMEM);
                if (e != cudaSuccess) return e;
                e = cudaFuncSetAttribute(axis_kstage_kernel<L, KS_SCALAR_Q2, SQ2_SIG_DIFFUSION>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AxisCfg<L>::SMEM);
                if (e != cudaSuccess) return e;
            }
        }
        attr = true;
    }
    const unsigned grid = (unsigned)a.ncolTiles * (unsigned)a.nbatch;
    if (ks.fastKind == KS_SCALAR_Q2) {
        const int sig = sq2_signature(ks.sq2);
        if (L >= 64 && sig == SQ2_SIG_CAHN_HILLIARD) {
            if constexpr (L >= 64) axis_kstage_kernel<L, KS_SCALAR_Q2, SQ2_SIG_CAHN_HILLIARD><<<grid, AxisCfg<L>::THREADS, AxisCfg<L>::SMEM, st>>>(a, ks);
        } else if (L >= 64 && sig == SQ2_SIG_DIFFUSION) {
            if constexpr (L >= 64) axis_kstage_kernel<L, KS_SCALAR_Q2, SQ2_SIG_DIFFUSION><<<grid, AxisCfg<L>::THREADS, AxisCfg<L>::SMEM, st>>>(a, ks);
        } else {
            axis_kstage_kernel<L, KS_SCALAR_Q2, -1><<<grid, AxisCfg<L>::THREADS, AxisCfg<L>::SMEM, st>>>(a, ks);
        }
    } else {
        axis_kstage_kernel<L, KS_GENERIC, -1><<<grid, AxisCfg<L>::THREADS, AxisCfg<L>::SMEM, st>>>(a, ks);
    }
    return cudaGetLastError();
}

#define CUPSS_FOR_SIZES(X) X(1) X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192)

cudaError_t launch_axis_plain(int L, int dir, const AxisArgs& a, cudaStream_t st) {
    switch (L) {
#define X(N) case N: return launch_plain_L<N>(dir, a, st);
        CUPSS_FOR_SIZES(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_axis_kstage(int L, const AxisArgs& a, const KStageD& ks, cudaStream_t st) {
    switch (L) {
#define X(N) case N: return launch_kstage_L<N>(a, ks, st);
        CUPSS_FOR_SIZES(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

int axis_tile_cols(int L) {
    switch (L) {
#define X(N) case N: return AxisCfg<N>::C;
        CUPSS_FOR_SIZES(X)
#undef X
    }
    return 0;
}

bool fft_size_supported(int n) { return axis_tile_cols(n) != 0; }

cudaError_t launch_xgpu_barrier(const XBarrier& b, cudaStream_t st) {
    xgpu_barrier_kernel<<<1, 32, 0, st>>>(b);
    return cudaGetLastError();
}

cudaError_t launch_bump_counter(unsigned int* counter, cudaStream_t st) {
    bump_counter_kernel<<<1, 1, 0, st>>>(counter);
    return cudaGetLastError();
}

}  // namespace cupss
