// kernels_axis.cu -- strided-axis FFT passes (y and z) of the half-spectrum, plain and fused.
//
// A CTA owns a tile [L rows] x [C columns] of one k-space array: the L points of C neighbouring kx columns
// along the transformed axis, held in shared memory as float4 = two neighbouring columns, so every global
// and shared access is a 128-bit access and a quarter-warp moves one contiguous, 128-byte-aligned row
// segment (C = 16 float2).  The transform runs in place on the tile, level by level (fft_core.cuh); rows are
// permuted for free on the way in / out because a row is a whole coalesced segment.
//
//   axis_plain_kernel  : load -> FFT -> store                      (y passes of a 3-D transform;
//                        optional dealias mask on load for extra inverse transforms)
//   axis_kstage_kernel : [load -> forward FFT] -> per-mode update of every field of the sweep
//                        (kstage_point) -> [dealias -> inverse FFT -> store]
// The second is the heart of the step: the last level of the forward transform of the nonlinear term, the
// semi-implicit Euler update, the dealiasing mask and the first level of the next inverse transform happen
// on the same registers; the data never leaves the SM between the two transforms.
#include "kernels_axis.cuh"

namespace cupss {

__global__ void bump_counter_kernel(unsigned int* c) { *c += 1u; }

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// A peer that never arrives (dead process, a rank stuck for longer than the time-out in host code) must not let the step
// run on with stale or half-written receive slots: the error word is raised in device AND mapped host memory and the
// kernel traps, which fails every later call on this context -- the engine reports the time-out (engine.cu: comm_guard).
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
        const unsigned long long t0 = global_timer_ns();
        unsigned int seen, spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
            if ((int)(seen - target) >= 0) break;
            if ((++spins & 1023u) == 0u && b.timeoutNs && global_timer_ns() - t0 > b.timeoutNs) {
                *b.error = 1;
                if (b.hostError) *reinterpret_cast<volatile int*>(b.hostError) = 1;
                __threadfence_system();
                __trap();
            }
        } while (true);
        __threadfence_system();
    }
}

// ---------------------------------------------------------------- dispatch
template <class K>
static cudaError_t set_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return cudaSuccess;
}

template <int L>
static cudaError_t launch_plain_L(int dir, const AxisArgs& a, cudaStream_t st) {
    using Cfg = AxisCfg<L>;
    constexpr bool CLUSTER = Cfg::CL > 1;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaSuccess;
        if constexpr (CLUSTER) {
            e = set_smem(axis_plain_cluster_kernel<L, -1, false>, Cfg::SMEM);
            if (e == cudaSuccess) e = set_smem(axis_plain_cluster_kernel<L, 1, false>, Cfg::SMEM);
            if (e == cudaSuccess) e = set_smem(axis_plain_cluster_kernel<L, 1, true>, Cfg::SMEM);
        } else {
            e = set_smem(axis_plain_kernel<L, -1, false>, Cfg::SMEM);
            if (e == cudaSuccess) e = set_smem(axis_plain_kernel<L, 1, false>, Cfg::SMEM);
            if (e == cudaSuccess) e = set_smem(axis_plain_kernel<L, 1, true>, Cfg::SMEM);
        }
        if (e != cudaSuccess) return e;
        attr = true;
    }
    // pruned inverse passes: column tiles beyond the dealias cut-off produce nothing -- they are not launched at all
    AxisArgs la = a;
    if (a.pruneOn && dir > 0) {
        const int live = a.pruneCutX / Cfg::C + 1 - a.ctBase;   // live tiles of this launch's column chunk
        if (live <= 0) return cudaSuccess;
        if (live < la.ncolTiles) la.ncolTiles = live;
    }
    const unsigned grid = (unsigned)la.ncolTiles * (unsigned)la.nbatch * (unsigned)Cfg::CL;   // CL CTAs (one cluster) per tile
    if (grid == 0) return cudaSuccess;
    if (dir < 0 && a.maskOn) return cudaErrorInvalidValue;   // the mask only exists on inverse transforms
    if constexpr (CLUSTER) {
        if (dir < 0) axis_plain_cluster_kernel<L, -1, false><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(la);
        else if (a.maskOn) axis_plain_cluster_kernel<L, 1, true><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(la);
        else axis_plain_cluster_kernel<L, 1, false><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(la);
    } else {
        if (dir < 0) axis_plain_kernel<L, -1, false><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(la);
        else if (a.maskOn) axis_plain_kernel<L, 1, true><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(la);
        else axis_plain_kernel<L, 1, false><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(la);
    }
    return cudaGetLastError();
}

template <int L, int KIND, int SIG>
static cudaError_t launch_kstage_variant(unsigned grid, const AxisArgs& a, const KStageD& ks, cudaStream_t st) {
    using Cfg = AxisCfg<L>;
    static bool attr = false;
    if constexpr (Cfg::CL > 1) {
        if (!attr) {
            cudaError_t e = set_smem(axis_kstage_cluster_kernel<L, KIND, SIG>, Cfg::SMEM);
            if (e != cudaSuccess) return e;
            attr = true;
        }
        axis_kstage_cluster_kernel<L, KIND, SIG><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a, ks);
    } else {
        if (!attr) {
            cudaError_t e = set_smem(axis_kstage_kernel<L, KIND, SIG>, Cfg::SMEM);
            if (e != cudaSuccess) return e;
            attr = true;
        }
        axis_kstage_kernel<L, KIND, SIG><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a, ks);
    }
    return cudaGetLastError();
}

template <int L>
static cudaError_t launch_kstage_L(const AxisArgs& a, const KStageD& ks, cudaStream_t st) {
    const unsigned grid = (unsigned)a.ncolTiles * (unsigned)a.nbatch * (unsigned)AxisCfg<L>::CL;
    if (ks.fastKind == KS_SCALAR_Q2) {
        const int sig = sq2_signature(ks.sq2);
        if constexpr (L >= 64) {
            if (sig == SQ2_SIG_CAHN_HILLIARD) return launch_kstage_variant<L, KS_SCALAR_Q2, SQ2_SIG_CAHN_HILLIARD>(grid, a, ks, st);
            if (sig == SQ2_SIG_DIFFUSION) return launch_kstage_variant<L, KS_SCALAR_Q2, SQ2_SIG_DIFFUSION>(grid, a, ks, st);
        }
        return launch_kstage_variant<L, KS_SCALAR_Q2, -1>(grid, a, ks, st);
    }
    return launch_kstage_variant<L, KS_GENERIC, -1>(grid, a, ks, st);
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

bool axis_kstage_geometry(int L, int* threads, size_t* smem, int* minBlocks) {
    switch (L) {
#define X(N) case N: *threads = AxisCfg<N>::THREADS; *smem = AxisCfg<N>::SMEM; *minBlocks = AxisCfg<N>::MINB; return true;
        CUPSS_FOR_SIZES(X)
#undef X
    }
    return false;
}

int axis_cluster_size(int L) {
    switch (L) {
#define X(N) case N: return AxisCfg<N>::CL;
        CUPSS_FOR_SIZES(X)
#undef X
    }
    return 1;
}
int axis_cta_rows(int L) {
    switch (L) {
#define X(N) case N: return AxisCfg<N>::LS;
        CUPSS_FOR_SIZES(X)
#undef X
    }
    return L;
}
// (CL-1) x S table of the level that couples the blocks of a cluster: entry (c-1) S + j = exp(-2 pi i j c / L)
int host_cross_twiddles(int L, float2* out) {
    const int CL = axis_cluster_size(L), S = axis_cta_rows(L);
    if (CL <= 1) return 0;
    for (int c = 1; c < CL; ++c)
        for (int j = 0; j < S; ++j) {
            const double ang = -2.0 * kPi * (double)(((long long)j * c) % L) / (double)L;
            out[(c - 1) * S + j] = make_float2((float)__builtin_cos(ang), (float)__builtin_sin(ang));
        }
    return (CL - 1) * S;
}

int host_level_twiddles(int L, float2* out) {
    switch (L) {
#define X(N) case N: fill_level_twiddles<N, 0>(out); return TwTable<N>::LEN;
        CUPSS_FOR_SIZES(X)
#undef X
    }
    return 0;
}

cudaError_t launch_xgpu_barrier(const XBarrier& b, cudaStream_t st) {
    xgpu_barrier_kernel<<<1, 32, 0, st>>>(b);
    return cudaGetLastError();
}

__global__ void real_expand_kernel(const float* __restrict__ in, float2* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = make_float2(in[i], 0.0f);
}
__global__ void real_compress_kernel(const float2* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i].x;
}
__global__ void spectrum_expand_kernel(const float2* __restrict__ half, float2* __restrict__ full, int sx, int sy, int sz, int pitch, int kyl, int z0, int zl, int cyclicP) {
    const size_t n = (size_t)sx * sy * zl;
    auto at = [&](int k, int j, int i) -> float2 {   // [src][k][jl][i]: block distribution src = j / kyl, cyclic src = j % P
        const int src = cyclicP > 1 ? j % cyclicP : j / kyl, jl = cyclicP > 1 ? j / cyclicP : j - src * kyl;
        return half[(((size_t)src * sz + k) * kyl + jl) * pitch + i];
    };
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(o % sx);
        const size_t r = o / sx;
        const int j = (int)(r % sy), k = z0 + (int)(r / sy);
        float2 v;
        if (i <= sx / 2) {
            v = at(k, j, i);
        } else {   // X(k) = conj X(-k)
            v = at((sz - k) % sz, (sy - j) % sy, sx - i);
            v.y = -v.y;
        }
        full[o] = v;
    }
}
cudaError_t launch_spectrum_expand(const float2* half, float2* full, int sx, int sy, int sz, int pitch, int kyl, int z0, int zl, int cyclicP, cudaStream_t st) {
    spectrum_expand_kernel<<<148 * 8, 256, 0, st>>>(half, full, sx, sy, sz, pitch, kyl, z0, zl, cyclicP);
    return cudaGetLastError();
}
// Hermitian part of a full spectrum, stored as the half spectrum: H(k) = (F(k) + conj F(-k)) / 2 for kx <= sx/2.
// This is what the reference keeps of a user-modified comp_array: toReal -> normalize (real part) -> toComp
// (src/field.cpp:59-66, 88-89).  Pad columns of the pitch stay untouched.
__global__ void spectrum_compress_kernel(const float2* __restrict__ full, float2* __restrict__ half, int sx, int sy, int sz, int pitch) {
    const int nc = sx / 2 + 1;
    const size_t n = (size_t)nc * sy * sz;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(o % nc);
        const size_t r = o / nc;
        const int j = (int)(r % sy), k = (int)(r / sy);
        const int mi = (sx - i) % sx, mj = (sy - j) % sy, mk = (sz - k) % sz;
        const float2 a = full[((size_t)k * sy + j) * sx + i];
        const float2 b = full[((size_t)mk * sy + mj) * sx + mi];
        half[((size_t)k * sy + j) * pitch + i] = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
    }
}
cudaError_t launch_spectrum_compress(const float2* full, float2* half, int sx, int sy, int sz, int pitch, cudaStream_t st) {
    spectrum_compress_kernel<<<148 * 8, 256, 0, st>>>(full, half, sx, sy, sz, pitch);
    return cudaGetLastError();
}
cudaError_t launch_real_expand(const float* in, float2* out, size_t n, cudaStream_t st) {
    real_expand_kernel<<<148 * 8, 256, 0, st>>>(in, out, n);
    return cudaGetLastError();
}
cudaError_t launch_real_compress(const float2* in, float* out, size_t n, cudaStream_t st) {
    real_compress_kernel<<<148 * 8, 256, 0, st>>>(in, out, n);
    return cudaGetLastError();
}

cudaError_t launch_bump_counter(unsigned int* counter, cudaStream_t st) {
    bump_counter_kernel<<<1, 1, 0, st>>>(counter);
    return cudaGetLastError();
}

}  // namespace cupss
