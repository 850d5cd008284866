// kernels_axis.cuh -- device code of the strided-axis passes (see kernels_axis.cu for the overview and the launchers).
// A header so that the plan-specialised k stage can be compiled at run time (NVRTC, engine.cu) from the same source.
#pragma once
#include "kernels.h"

namespace cupss {

// Long axes (1024, 2048, 4096 rows) are shared by a thread-block cluster of CL = L / 512 CTAs: every CTA keeps 512 rows x 16
// columns of the tile in its own shared memory (the geometry of the 512-point kernels: 128-byte row segments, 64 KB per CTA,
// 2-3 CTAs per SM), and the one level of the transform that couples the blocks runs over distributed shared memory
// (cross_scatter / cross_gather below).  Without a cluster a 4096-row tile fits one SM's shared memory only 4 columns wide
// (32-byte segments, one CTA per SM): 29 % of the HBM roofline on Cahn-Hilliard 2-D 4096^2 (profiles/r2a_bench.json).
template <int L> struct AxisCfg {
    static constexpr int CL = (L >= 1024 && L <= 4096) ? L / 512 : 1;   // CTAs per cluster
    static constexpr int LS = L / CL;                                   // rows per CTA
    static constexpr int C = LS <= 512 ? 16 : (LS <= 2048 ? 8 : (LS <= 4096 ? 4 : 2));
    static constexpr int CP = C / 2;                                   // float4 column pairs
    static constexpr int NVMAX = LS / FftLevels<LS>::min_rad();        // most virtual threads any level has
    static constexpr size_t TILE = (size_t)LS * CP * sizeof(float4);
    static constexpr size_t TWB = ((size_t)TwTable<LS>::LEN * sizeof(float2) + 15) / 16 * 16;
    static constexpr size_t SMEM = (FftLevels<LS>::n > 1 ? TILE : 0) + TWB + 16;   // tile, twiddle table, mbarrier of the TMA prologue
    static constexpr int WANT = CP * NVMAX;
    static constexpr int TMAX = SMEM > 100 * 1024 ? 512 : 256;
    static constexpr int THREADS = WANT < 32 ? 32 : (WANT > TMAX ? TMAX : WANT);
    static constexpr int TV = THREADS / CP;
    static constexpr int MINB = THREADS >= 512 ? 1 : (SMEM > 100 * 1024 ? 1 : (SMEM > 70 * 1024 ? 2 : 3));
};

// Element offset of (batch b, row, column col).  32-bit arithmetic: the engine refuses arrays of 2^31 elements or more.
__device__ __forceinline__ unsigned axis_off(const AxisAddr& a, unsigned b, unsigned row, unsigned col) {
    return b * (unsigned)a.bs + ((row >> a.rpcShift) & (unsigned)a.chunkMask) * (unsigned)a.cs + ((row >> a.locShift) & (unsigned)a.rpcMask) * (unsigned)a.rs + col;
}
// Same with the row-independent part (b * bs + col) hoisted by the caller.
__device__ __forceinline__ unsigned row_off(const AxisAddr& a, unsigned row) {
    return ((row >> a.rpcShift) & (unsigned)a.chunkMask) * (unsigned)a.cs + ((row >> a.locShift) & (unsigned)a.rpcMask) * (unsigned)a.rs;
}

// Destination of row `row`: local array, or the receive buffer of the peer that owns the row (fused slab exchange).
// `lbase` = out + b*bs + col (local), `pbase` = pushBase + b*pushBs + col (element offset inside every peer's arena).
__device__ __forceinline__ float2* axis_dst(const AxisArgs& a, float2* lbase, unsigned long long pbase, unsigned row) {
    if (a.pushOn)
        return a.push[(row >> a.pushShift) & (unsigned)a.pushPeerMask] + pbase + (unsigned long long)(((row >> a.pushLocShift) & (unsigned)a.pushMask)) * (unsigned long long)a.pushRs;
    return lbase + row_off(a.aout, row);
}

__device__ __forceinline__ float4 ld4(const float2* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float2* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// 16-byte asynchronous global -> shared copy (LDGSTS, L1 bypassed): the whole input tile of a CTA is put in flight by
// its first few instructions, with no register staging, so the HBM latency is paid once per tile and overlaps the
// arithmetic of the other CTAs resident on the SM.
__device__ __forceinline__ void cp_async16(float4* smemDst, const float2* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------- TMA (cp.async.bulk.tensor) tile prologue
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CUPSS_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CUPSS_MBAR_DONE_%=;\n"
        "bra CUPSS_MBAR_WAIT_%=;\n"
        "CUPSS_MBAR_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// box of the 4-D tensor map at (c0, c1, c2, c3) -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, int c0, int c1, int c2, int c3, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
// plain bulk copy global -> shared on the same barrier (twiddle table)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Tile prologue by TMA, rows in NATURAL order: one elected thread issues a few box loads (tmaBoxRows rows x 128 bytes each)
// plus the twiddle table; nobody computes an address.  Rows beyond the dealias cut-off (keepLo < row < keepHi) are zero-filled
// by the threads, row keepLo itself (one 128-byte segment that no box of a power-of-two row count covers) by cp.async.
// tile_wait_tma returns with the tile complete and visible to every thread of the CTA.
template <int L, int CP, int THREADS>
__device__ __forceinline__ void tile_fetch_tma(float4* tile, float2* twS, unsigned long long* bar, const AxisArgs& a, unsigned ct, unsigned b,
                                               unsigned keepLo, unsigned keepHi, const float2* ibase, bool valid) {
    constexpr int C = 2 * CP;
    constexpr unsigned TWBYTES = (unsigned)TwTable<L>::LEN * sizeof(float2);
    const bool pruned = keepLo < (unsigned)L;
    const unsigned B = (unsigned)a.tmaBoxRows;
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned nbox = pruned ? 2u * (keepLo / B) : (unsigned)L / B;
        mbar_expect_tx(bar, nbox * B * (unsigned)(CP * sizeof(float4)) + TWBYTES);
        bulk_load(twS, a.tw, TWBYTES, bar);
        for (unsigned i = 0; i < nbox; ++i) {
            // pruned: boxes over [0, keepLo) and [keepHi, L)
            const unsigned r0 = pruned ? (i < nbox / 2 ? i * B : keepHi + (i - nbox / 2) * B) : i * B;
            const int v0 = (int)((r0 >> a.ain.locShift) & (unsigned)a.ain.rpcMask), v1 = (int)((r0 >> a.ain.rpcShift) & (unsigned)a.ain.chunkMask), v2 = (int)b;
            auto pick = [&](int role) { return role == 0 ? v0 : (role == 1 ? v1 : v2); };
            tma_load_4d(tile + (size_t)r0 * CP, a.tmap, (int)(ct * C * 2), pick(a.tmaSlot[0]), pick(a.tmaSlot[1]), pick(a.tmaSlot[2]), bar);
        }
    }
    if (pruned) {
        const unsigned cp = threadIdx.x % CP, tv = threadIdx.x / CP;
        for (unsigned p = keepLo + 1 + tv; p < keepHi; p += THREADS / CP) tile[p * CP + cp] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (tv == 0) {
            if (valid) cp_async16(tile + keepLo * CP + cp, ibase + row_off(a.ain, keepLo));
            else tile[keepLo * CP + cp] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
}
__device__ __forceinline__ void tile_wait_tma(unsigned long long* bar) {
    cp_async_wait_all();
    mbar_wait(bar, 0);
    __syncthreads();
}

// ---------------------------------------------------------------- thread-block clusters: the level that couples the CTAs' blocks
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned dsmem_addr(unsigned localAddr, unsigned ctaRank) {   // same offset in CTA `ctaRank` of the cluster
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(localAddr), "r"(ctaRank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_dsmem4(unsigned addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// First level of a decimation-in-frequency transform of L = CL * S rows held S rows per CTA (natural row order): a radix-CL
// butterfly over rows j, j + S, ..., j + (CL-1) S, the twiddle w_L^(j c), and output c goes to row j of CTA c's tile.  CTA
// `crank` does this for its slice of S / CL values of j, reading the CL rows straight from global memory (`gld`, whole
// 128-byte segments) and scattering over distributed shared memory: per CTA one tile in, one tile out, whatever CL is.
// Afterwards CTA c holds the S-point sequence whose transform is the frequencies c, c + CL, c + 2 CL, ...
// twX: (CL-1) x S table, entry (c-1) S + j = exp(-2 pi i j c / L).
template <int L, int CL, int SIGN, int CP, int TV, class LdRow>
__device__ __forceinline__ void cross_scatter(float4* tile, unsigned crank, unsigned tv, unsigned cp, const float2* __restrict__ twX, LdRow gld) {
    constexpr unsigned S = L / CL, JS = S / CL;
    const unsigned base = smem_u32(tile);
#pragma unroll 1
    for (unsigned jj = tv; jj < JS; jj += TV) {
        const unsigned j = crank * JS + jj;
        float2 x0[CL], x1[CL];
#pragma unroll
        for (unsigned q = 0; q < (unsigned)CL; ++q) {
            const float4 t = gld(j + S * q);
            x0[q] = make_float2(t.x, t.y); x1[q] = make_float2(t.z, t.w);
        }
        Dft<CL, SIGN>::run(x0);
        Dft<CL, SIGN>::run(x1);
#pragma unroll
        for (unsigned c = 1; c < (unsigned)CL; ++c) {
            const float2 w = __ldg(twX + (c - 1) * S + j);
            x0[c] = SIGN > 0 ? cmul_conj(x0[c], w) : cmul(x0[c], w);
            x1[c] = SIGN > 0 ? cmul_conj(x1[c], w) : cmul(x1[c], w);
        }
        const unsigned off = base + (j * CP + cp) * (unsigned)sizeof(float4);
#pragma unroll
        for (unsigned c = 0; c < (unsigned)CL; ++c) st_dsmem4(dsmem_addr(off, c), make_float4(x0[c].x, x0[c].y, x1[c].x, x1[c].y));
    }
}
// Last level of the matching decimation-in-time inverse: CTA c holds (natural order) the inverse S-point transform of the
// frequencies congruent to c; row j of every CTA is gathered, twiddled by conj w_L^(j c) and butterflied into the real-space
// rows j, j + S, ..., which `gst` stores to global memory (or pushes to the peer that owns them).
template <int L, int CL, int CP, int TV, class StRow>
__device__ __forceinline__ void cross_gather(const float4* tile, unsigned crank, unsigned tv, unsigned cp, const float2* __restrict__ twX, StRow gst) {
    constexpr unsigned S = L / CL, JS = S / CL;
    const unsigned base = smem_u32(tile);
#pragma unroll 1
    for (unsigned jj = tv; jj < JS; jj += TV) {
        const unsigned j = crank * JS + jj;
        const unsigned off = base + (j * CP + cp) * (unsigned)sizeof(float4);
        float2 x0[CL], x1[CL];
#pragma unroll
        for (unsigned c = 0; c < (unsigned)CL; ++c) {
            const float4 t = ld_dsmem4(dsmem_addr(off, c));
            x0[c] = make_float2(t.x, t.y); x1[c] = make_float2(t.z, t.w);
        }
#pragma unroll
        for (unsigned c = 1; c < (unsigned)CL; ++c) {
            const float2 w = __ldg(twX + (c - 1) * S + j);
            x0[c] = cmul_conj(x0[c], w);
            x1[c] = cmul_conj(x1[c], w);
        }
        Dft<CL, +1>::run(x0);
        Dft<CL, +1>::run(x1);
#pragma unroll
        for (unsigned q = 0; q < (unsigned)CL; ++q) gst(j + S * q, make_float4(x0[q].x, x0[q].y, x1[q].x, x1[q].y));
    }
}

// Level twiddle table -> shared memory with the same asynchronous copies as the tile (no LDG -> STS round trip through
// registers in front of the first barrier; profiles/r02z: that chain drew 11-14 % of the stall samples of these kernels).
template <int L, int THREADS>
__device__ __forceinline__ void tw_fetch(float2* twS, const float2* __restrict__ tw) {
    constexpr int LEN = TwTable<L>::LEN;
    if constexpr (LEN >= 2 && LEN % 2 == 0) {
        for (int i = threadIdx.x; i < LEN / 2; i += THREADS) cp_async16(reinterpret_cast<float4*>(twS) + i, tw + 2 * i);
    } else {
        for (int i = threadIdx.x; i < LEN; i += THREADS) twS[i] = __ldg(tw + i);
    }
}

// Tile prologue: position p of the tile <- row rowOf(p) of the input (natural order for a forward transform, frequency
// order for an inverse one).  Rows that are known zeros (keep(row) false) and invalid column pairs are zero-filled.
template <int L, int CP, int TV, class RowOf, class Keep>
__device__ __forceinline__ void tile_fetch(float4* tile, unsigned tv, unsigned cp, bool valid, const float2* ibase, const AxisAddr& ain,
                                           RowOf rowOf, Keep keep) {
#pragma unroll 4
    for (unsigned p = tv; p < (unsigned)L; p += TV) {
        const unsigned row = rowOf(p);
        float4* dst = tile + p * CP + cp;
        if (valid && keep(row)) cp_async16(dst, ibase + row_off(ain, row));
        else *dst = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// One level over the tile: every real thread walks its virtual threads (v = tv, tv + TV, ...).
//   ld(pos, frow) -> float4, st(pos, frow, value);  pos = position inside the tile, frow = frequency row that
//   position holds in the digit-reversed order (only meaningful on the innermost level, M == 1).
template <int L, int LV, int DIR, int TV, bool DIF = (DIR < 0), class LdF, class StF>
__device__ __forceinline__ void tile_level(unsigned tv, const float2* __restrict__ twS, LdF ld, StF st) {
    using G = LevelGeom<L, LV>;
    constexpr unsigned R = G::R, M = G::M, N = G::N;
#pragma unroll 1
    for (unsigned v = tv; v < (unsigned)G::NV; v += TV) {
        const unsigned blk = v / M, j = v % M;
        const unsigned row0 = blk * N + j;
        const unsigned f0 = M == 1 ? freq_of_pos<L>(v * R) : 0u;
        float2 x0[R], x1[R];
#pragma unroll
        for (unsigned q = 0; q < R; ++q) {
            const float4 t = ld(row0 + M * q, f0 + (L / R) * q);
            x0[q] = make_float2(t.x, t.y); x1[q] = make_float2(t.z, t.w);
        }
        level_butterfly2<L, LV, DIR, DIF>(x0, x1, j, twS);
#pragma unroll
        for (unsigned q = 0; q < R; ++q) st(row0 + M * q, f0 + (L / R) * q, make_float4(x0[q].x, x0[q].y, x1[q].x, x1[q].y));
    }
}

// MASK: apply the full dealias mask on load (extra inverse transforms of a sweep; AxisArgs::maskOn).
template <int L, int DIR, bool MASK>
__global__ void __launch_bounds__(AxisCfg<L>::THREADS, AxisCfg<L>::MINB) axis_plain_kernel(const __grid_constant__ AxisArgs a) {
    using Cfg = AxisCfg<L>;
    constexpr int C = Cfg::C, CP = Cfg::CP, TV = Cfg::TV, n = FftLevels<L>::n;
    extern __shared__ float4 smem4[];
    float4* tile = smem4;
    float2* twS = reinterpret_cast<float2*>(smem4 + (n > 1 ? (size_t)L * CP : 0));
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(twS) + Cfg::TWB);
    const unsigned cp = threadIdx.x % CP, tv = threadIdx.x / CP;
    const unsigned ct = blockIdx.x % (unsigned)a.ncolTiles + (unsigned)a.ctBase, b = blockIdx.x / (unsigned)a.ncolTiles;
    const unsigned col = ct * C + 2 * cp;
    const bool valid = col < (unsigned)a.ncol;

    if (a.pruneOn) {   // CTA-uniform: the whole tile is outside the dealias cut-off -> output stays zero
        const int iyT = a.kyBase + (int)b * a.kyStride;
        const int nyT = iyT > a.sy / 2 ? a.sy - iyT : iyT;
        if ((int)(ct * C) > a.pruneCutX || (a.axis == 2 && nyT > a.pruneCutY)) return;
    }
    auto load_twiddles = [&]() { tw_fetch<L, Cfg::THREADS>(twS, a.tw); };

    const float2* ibase = a.in + (b * (unsigned)a.ain.bs + col);
    float2* lbase = a.out + (b * (unsigned)a.aout.bs + col);
    const unsigned long long pbase = (unsigned long long)a.pushBase + (unsigned long long)b * (unsigned long long)a.pushBs + col;
    // rows beyond rowCut (|n| > rowCut) are known zeros: not loaded.  rowCut < 0: off.
    const unsigned keepLo = a.rowCut >= 0 ? (unsigned)a.rowCut : (unsigned)L, keepHi = a.rowCut >= 0 ? (unsigned)(L - a.rowCut) : 0u;

    auto keepRow = [&](unsigned row) -> bool { return row <= keepLo || row >= keepHi; };
    auto gload = [&](unsigned row) -> float4 {   // direct path (single-level transforms, no tile)
        if (!(valid && keepRow(row))) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return ld4(ibase + row_off(a.ain, row));
    };
    auto masked = [&](float4 t, unsigned row) -> float4 {
        if constexpr (MASK) {
            const int iy = a.axis == 2 ? a.kyBase + (int)b * a.kyStride : (a.axis == 1 ? (int)row : 0);
            const int iz = a.axis == 2 ? (int)row : 0;
            if (!dealias_keep((int)col, iy, iz, a.sx, a.sy, a.sz, a.cutx, a.cuty, a.cutz)) { t.x = 0.0f; t.y = 0.0f; }
            if (!dealias_keep((int)col + 1, iy, iz, a.sx, a.sy, a.sz, a.cutx, a.cuty, a.cutz)) { t.z = 0.0f; t.w = 0.0f; }
        }
        return t;
    };
    auto gstore = [&](unsigned row, float4 v) {
        if (valid) st4(axis_dst(a, lbase, pbase, row), v);
    };
    auto sld = [&](unsigned pos, unsigned) -> float4 { return tile[pos * CP + cp]; };
    auto sst = [&](unsigned pos, unsigned, float4 v) { tile[pos * CP + cp] = v; };
    auto fetch_natural = [&]() {   // whole input tile + twiddle table in flight, then wait
        if (a.tmaOn) {
            tile_fetch_tma<L, CP, Cfg::THREADS>(tile, twS, bar, a, ct, b, keepLo, keepHi, ibase, valid);
            tile_wait_tma(bar);
        } else {
            tile_fetch<L, CP, TV>(tile, tv, cp, valid, ibase, a.ain, [](unsigned p) { return p; }, keepRow);
            load_twiddles();
            cp_async_wait_all();
            __syncthreads();
        }
    };

    if constexpr (DIR < 0) {   // forward: natural rows in, frequency rows out
        auto gst = [&](unsigned, unsigned frow, float4 v) { gstore(frow, v); };
        if constexpr (n == 1) {
            auto gld = [&](unsigned pos, unsigned) -> float4 { return gload(pos); };
            tile_level<L, 0, DIR, TV>(tv, twS, gld, gst);
        } else {
            fetch_natural();
            tile_level<L, 0, DIR, TV>(tv, twS, sld, sst);
            __syncthreads();
            if constexpr (n >= 3) { tile_level<L, 1, DIR, TV>(tv, twS, sld, sst); __syncthreads(); }
            if constexpr (n >= 4) { tile_level<L, 2, DIR, TV>(tv, twS, sld, sst); __syncthreads(); }
            tile_level<L, n - 1, DIR, TV>(tv, twS, sld, gst);
        }
    } else {
        // inverse: frequency rows in, in NATURAL order as well (decimation in frequency with the conjugate twiddles), so the
        // live rows of a pruned transform are two contiguous ranges -- box loads -- and the real-space rows leave permuted
        if constexpr (n == 1) {
            auto gst1 = [&](unsigned pos, unsigned, float4 v) { gstore(pos, v); };
            auto gld = [&](unsigned, unsigned frow) -> float4 { return masked(gload(frow), frow); };
            tile_level<L, 0, DIR, TV>(tv, twS, gld, gst1);
        } else {
            auto gst = [&](unsigned, unsigned yrow, float4 v) { gstore(yrow, v); };
            fetch_natural();
            auto mld = [&](unsigned pos, unsigned) -> float4 { return masked(tile[pos * CP + cp], pos); };
            tile_level<L, 0, DIR, TV, true>(tv, twS, mld, sst);
            __syncthreads();
            if constexpr (n >= 3) { tile_level<L, 1, DIR, TV, true>(tv, twS, sld, sst); __syncthreads(); }
            if constexpr (n >= 4) { tile_level<L, 2, DIR, TV, true>(tv, twS, sld, sst); __syncthreads(); }
            tile_level<L, n - 1, DIR, TV, true>(tv, twS, sld, gst);
        }
    }
}

// Same pass for an axis shared by a cluster (AxisCfg<L>::CL > 1); both directions run decimation in frequency from natural
// row order: cross level (global loads -> distributed shared memory), then the 512-point levels on the CTA's own block.  The
// block of CTA c ends with position p holding output row c + CL * freq_of_pos<512>(p).
template <int L, int DIR, bool MASK>
__global__ void __cluster_dims__(AxisCfg<L>::CL, 1, 1) __launch_bounds__(AxisCfg<L>::THREADS, AxisCfg<L>::MINB)
axis_plain_cluster_kernel(const __grid_constant__ AxisArgs a) {
    using Cfg = AxisCfg<L>;
    constexpr int C = Cfg::C, CP = Cfg::CP, TV = Cfg::TV, CL = Cfg::CL, S = Cfg::LS, n = FftLevels<S>::n;
    static_assert(CL > 1 && n >= 2, "cluster kernel");
    extern __shared__ float4 smem4[];
    float4* tile = smem4;
    float2* twS = reinterpret_cast<float2*>(smem4 + (size_t)S * CP);
    const unsigned cp = threadIdx.x % CP, tv = threadIdx.x / CP;
    const unsigned crank = cluster_ctarank(), tileId = blockIdx.x / CL;
    const unsigned ct = tileId % (unsigned)a.ncolTiles + (unsigned)a.ctBase, b = tileId / (unsigned)a.ncolTiles;
    const unsigned col = ct * C + 2 * cp;
    const bool valid = col < (unsigned)a.ncol;

    if (a.pruneOn) {   // cluster-uniform: the whole tile is outside the dealias cut-off -> output stays zero
        const int iyT = a.kyBase + (int)b * a.kyStride;
        const int nyT = iyT > a.sy / 2 ? a.sy - iyT : iyT;
        if ((int)(ct * C) > a.pruneCutX || (a.axis == 2 && nyT > a.pruneCutY)) return;
    }
    const float2* ibase = a.in + (b * (unsigned)a.ain.bs + col);
    float2* lbase = a.out + (b * (unsigned)a.aout.bs + col);
    const unsigned long long pbase = (unsigned long long)a.pushBase + (unsigned long long)b * (unsigned long long)a.pushBs + col;
    const unsigned keepLo = a.rowCut >= 0 ? (unsigned)a.rowCut : (unsigned)L, keepHi = a.rowCut >= 0 ? (unsigned)(L - a.rowCut) : 0u;
    auto gld = [&](unsigned row) -> float4 {
        if (!(valid && (row <= keepLo || row >= keepHi))) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float4 t = ld4(ibase + row_off(a.ain, row));
        if constexpr (MASK) {
            const int iy = a.axis == 2 ? a.kyBase + (int)b * a.kyStride : (a.axis == 1 ? (int)row : 0);
            const int iz = a.axis == 2 ? (int)row : 0;
            if (!dealias_keep((int)col, iy, iz, a.sx, a.sy, a.sz, a.cutx, a.cuty, a.cutz)) { t.x = 0.0f; t.y = 0.0f; }
            if (!dealias_keep((int)col + 1, iy, iz, a.sx, a.sy, a.sz, a.cutx, a.cuty, a.cutz)) { t.z = 0.0f; t.w = 0.0f; }
        }
        return t;
    };
    auto gst = [&](unsigned, unsigned frow, float4 v) {   // frow: local frequency of the 512-point transform
        if (valid) st4(axis_dst(a, lbase, pbase, crank + (unsigned)CL * frow), v);
    };
    auto sld = [&](unsigned pos, unsigned) -> float4 { return tile[pos * CP + cp]; };
    auto sst = [&](unsigned pos, unsigned, float4 v) { tile[pos * CP + cp] = v; };

    tw_fetch<S, Cfg::THREADS>(twS, a.tw);
    cluster_sync();   // every CTA of the cluster is running: its shared memory may be written
    cross_scatter<L, CL, DIR, CP, TV>(tile, crank, tv, cp, a.twX, gld);
    cp_async_wait_all();
    cluster_sync();   // all blocks complete (also a CTA barrier)
    tile_level<S, 0, DIR, TV, true>(tv, twS, sld, sst);
    __syncthreads();
    if constexpr (n >= 3) { tile_level<S, 1, DIR, TV, true>(tv, twS, sld, sst); __syncthreads(); }
    tile_level<S, n - 1, DIR, TV, true>(tv, twS, sld, gst);
}

template <class PLAN> struct PlanNsrc { static constexpr int value = PLAN::nsrc; };
template <> struct PlanNsrc<void> { static constexpr int value = 0; };
template <class PLAN> struct PlanNout { static constexpr int value = PLAN::nout; };
template <> struct PlanNout<void> { static constexpr int value = 0; };

// PLAN: compile-time structure of the sweep for KIND == KS_JIT (kstage.cuh), void otherwise.
// A cluster-shared axis (AxisCfg<L>::CL > 1): the CTA runs the S = 512-point levels on its own block between the two cross
// levels; position p of CTA c's block holds frequency row c + CL * freq_of_pos<S>(p), so with q unrolled the row of the
// fused level is still "f0 + (L / R) q" with f0 < L / R -- the per-mode code below is the same for both geometries.
template <int L, int KIND, int SIG, class PLAN>
__device__ __forceinline__ void axis_kstage_body(const AxisArgs& a, const KStageD& ks) {
    using Cfg = AxisCfg<L>;
    constexpr int C = Cfg::C, CP = Cfg::CP, TV = Cfg::TV, CL = Cfg::CL, S = Cfg::LS, n = FftLevels<S>::n;
    constexpr int LAST = n - 1;
    using G = LevelGeom<S, LAST>;
    constexpr unsigned R = G::R;
    extern __shared__ float4 smem4[];
    float4* tile = smem4;
    float2* twS = reinterpret_cast<float2*>(smem4 + (n > 1 ? (size_t)S * CP : 0));
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(twS) + Cfg::TWB);
    const unsigned cp = threadIdx.x % CP, tv = threadIdx.x / CP;
    const unsigned crank = CL > 1 ? cluster_ctarank() : 0u, tileId = blockIdx.x / (unsigned)CL;
    const unsigned ct = tileId % (unsigned)a.ncolTiles + (unsigned)a.ctBase, b = tileId / (unsigned)a.ncolTiles;
    const unsigned col = ct * C + 2 * cp;
    const bool valid = col < (unsigned)a.ncol, valid1 = col + 1 < (unsigned)a.ncol;

    // fixed (per thread) part of the mode index: columns = kx, and ky for a z pass
    const int iyFix = a.axis == 2 ? a.kyBase + (int)b * a.kyStride : 0;
    bool doInv = ks.hasInv != 0;
    if (doInv && a.pruneOn) {   // CTA-uniform: every mode of this tile is masked out -> nothing to transform or store
        const int nyT = iyFix > ks.sy / 2 ? ks.sy - iyFix : iyFix;
        if ((int)(ct * C) > a.pruneCutX || (a.axis == 2 && nyT > a.pruneCutY)) doInv = false;
    }

    // k-space arrays (sources, destinations) share the natural addressing of the pass output: b*bs + row*rs + col
    const unsigned kbase = b * (unsigned)a.aout.bs + col;
    const unsigned krs = (unsigned)a.aout.rs;
    const float2* ibase = a.in + (b * (unsigned)a.ain.bs + col);
    float2* lbase = a.out + kbase;
    const unsigned long long pbase = (unsigned long long)a.pushBase + (unsigned long long)b * (unsigned long long)a.pushBs + col;

    const bool tma = CL == 1 && n > 1 && a.tmaOn && ks.hasFwd;   // CTA-uniform
    if constexpr (CL > 1) {
        // cross level of the forward transform: rows straight from global memory, butterflied and scattered over the cluster
        tw_fetch<S, Cfg::THREADS>(twS, a.tw);
        cluster_sync();
        if (ks.hasFwd)
            cross_scatter<L, CL, -1, CP, TV>(tile, crank, tv, cp, a.twX, [&](unsigned row) -> float4 {
                return valid ? ld4(ibase + row_off(a.ain, row)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            });
    } else if constexpr (n > 1) {
        if (tma) tile_fetch_tma<L, CP, Cfg::THREADS>(tile, twS, bar, a, ct, b, (unsigned)L, 0u, ibase, valid);
        else if (ks.hasFwd) tile_fetch<L, CP, TV>(tile, tv, cp, valid, ibase, a.ain, [](unsigned p) { return p; }, [](unsigned) { return true; });
    }
    if (CL == 1 && !tma) tw_fetch<S, Cfg::THREADS>(twS, a.tw);
    if (ks.hasFwd) {
        // pull this tile's rows of the state spectrum (generic sweeps: of every pointwise source) towards L2 while the forward
        // transform runs: the generic evaluator reads its sources mode by mode inside a rolled loop, one exposed latency per row
        // (noisy KPZ-3D 512^3 k stage 0.807 -> 0.769 ms).  Without a forward transform there is nothing to hide the prefetch
        // behind and issuing it for the whole tile at once costs more than it saves (KPZ constraint sweep 0.775 -> 0.927 ms).
        const int nsrc = KIND == KS_SCALAR_Q2 ? 1 : ks.nsrc;
        for (int sI = 0; sI < nsrc; ++sI) {
            const float2* p0 = ks.src[sI] + (b * (unsigned)a.aout.bs + ct * C);
            for (unsigned r = threadIdx.x; r < (unsigned)S; r += Cfg::THREADS)   // the rows this CTA's block will hold: crank + CL r
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + (crank + (unsigned)CL * r) * krs));
        }
    }
    if (tma) {
        tile_wait_tma(bar);
    } else {
        cp_async_wait_all();
        if constexpr (CL > 1) cluster_sync(); else __syncthreads();
    }

    auto sld = [&](unsigned pos, unsigned) -> float4 { return tile[pos * CP + cp]; };
    auto sst = [&](unsigned pos, unsigned, float4 v) { tile[pos * CP + cp] = v; };

    // ---- forward levels 0 .. n-2 (the last level is fused with the k stage below)
    if (ks.hasFwd) {
        if constexpr (n > 1) {
            tile_level<S, 0, -1, TV>(tv, twS, sld, sst);
            __syncthreads();
            if constexpr (n >= 3) { tile_level<S, 1, -1, TV>(tv, twS, sld, sst); __syncthreads(); }
            if constexpr (n >= 4) { tile_level<S, 2, -1, TV>(tv, twS, sld, sst); __syncthreads(); }
        }
    }

    // ---- per-thread constants of the lean evaluator
    const OutD& od0 = ks.out[0];
    const int sRow = a.axis == 2 ? ks.sz : (a.axis == 1 ? ks.sy : 1);
    const float stepRow = a.axis == 2 ? ks.stepqz : ks.stepqy;
    const int cutRow = a.axis == 2 ? od0.cutz : od0.cuty;
    const float qxa = wavenumber((int)col, ks.sx, ks.stepqx), qxb = wavenumber((int)col + 1, ks.sx, ks.stepqx);
    const float qyFix = wavenumber(iyFix, ks.sy, ks.stepqy);
    const float qxa2 = CUPSS_FMUL(qxa, qxa), qxb2 = CUPSS_FMUL(qxb, qxb);
    const float qyFix2 = CUPSS_FMUL(qyFix, qyFix);
    // q^2 = (qx^2 + qy^2) + qz^2 in the reference's order; along y (2-D) the fixed part is qx^2 alone and the "+ 0" of the
    // missing axis is the identity on a sum of squares
    const float baseA = a.axis == 2 ? CUPSS_FADD(qxa2, qyFix2) : qxa2, baseB = a.axis == 2 ? CUPSS_FADD(qxb2, qyFix2) : qxb2;
    const bool fixY = (iyFix == 0) || (2 * iyFix == ks.sy);
    const bool fixSelfA = ((col == 0) || (2 * (int)col == ks.sx)) && fixY;
    const bool fixSelfB = (2 * ((int)col + 1) == ks.sx) && fixY;
    const int nyFix = iyFix > ks.sy / 2 ? ks.sy - iyFix : iyFix;
    const bool keepY = od0.inv && (a.axis != 2 || nyFix <= od0.cuty);
    const bool keepA = keepY && (int)col <= od0.cutx, keepB = keepY && (int)col + 1 <= od0.cutx;
    // lean evaluator of a noisy field (run-time compiled signatures only): the noise of the thread's modes is generated in a
    // rolled loop into the tile slots the thread has just emptied, and picked up by the unrolled evaluator below
    constexpr bool NOISE = KIND == KS_SCALAR_Q2 && SIG >= 0 && (SIG & SQ2_SIG_NOISE) != 0;
    static_assert(!NOISE || n > 1, "the noisy lean evaluator parks its noise in the tile");
    const unsigned int step = ((KIND != KS_SCALAR_Q2 || NOISE) && ks.stepCounter) ? *ks.stepCounter : 0u;

    // ---- fused level: last forward butterfly -> k stage -> first inverse butterfly
#pragma unroll 1
    for (unsigned v = tv; v < (unsigned)G::NV; v += TV) {
        const unsigned f0 = crank + (unsigned)CL * freq_of_pos<S>(v * R);
        float2 x0[R], x1[R];
        // state rows of this virtual thread: issued before the butterfly so that their (L2) latency overlaps it
        const unsigned rowStride = (L / R) * krs;
        const unsigned off0 = kbase + f0 * krs;
        float4 self[KIND == KS_SCALAR_Q2 ? R : 1];
        auto load_self = [&]() {
            const float2* sp = ks.src[0] + off0;
#pragma unroll
            for (unsigned q = 0; q < R; ++q)
                self[q] = __ldcg(reinterpret_cast<const float4*>(sp + q * rowStride));   // columns up to the pitch exist: no predicate, stores are guarded
        };
        if constexpr (KIND == KS_SCALAR_Q2 && !NOISE) load_self();   // (a noisy field: after the noise loop, which needs the registers)
        if (ks.hasFwd) {
#pragma unroll
            for (unsigned q = 0; q < R; ++q) {
                float4 t;
                if constexpr (n > 1) t = tile[(v * R + q) * CP + cp];
                else t = valid ? ld4(ibase + row_off(a.ain, v * R + q)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                x0[q] = make_float2(t.x, t.y); x1[q] = make_float2(t.z, t.w);
            }
            level_butterfly2<S, LAST, -1, true>(x0, x1, 0, twS);
        } else {
#pragma unroll
            for (unsigned q = 0; q < R; ++q) { x0[q] = make_float2(0.0f, 0.0f); x1[q] = make_float2(0.0f, 0.0f); }
        }

        if constexpr (NOISE) {
            // noise increments amp * xi of the thread's modes (row f0 + (L/R) q; columns col, col + 1): one generator call per
            // column pair -- the same calls, in the same arithmetic, as the generic evaluator makes.  Columns inside the
            // self-conjugate planes kx = 0, sx/2 (one lane of the first and of the last column tile) take white_noise_mode with
            // its mirror-row logic; every other pair is two Box-Muller transforms of the four words.
            const bool ampQ = od0.noise.q2n != 0;   // conserved noise: the amplitude follows q^2
            const bool planes = col == 0u || 2 * (int)col == ks.sx || 2 * ((int)col + 1) == ks.sx;
#pragma unroll 1
            for (unsigned q = 0; q < R; ++q) {
                const int row = (int)(f0 + (L / R) * q);
                const int iy = a.axis == 2 ? iyFix : (a.axis == 1 ? row : 0), iz = a.axis == 2 ? row : 0;
                float4 g = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (valid) {
                    unsigned int c[4];
                    philox_pair(ks, (int)(col >> 1), iy, iz, (unsigned int)od0.fieldId, step, c);
                    float ampA = od0.noiseAmp0, ampB = od0.noiseAmp0;
                    if (ampQ) {
                        const float qr = wavenumber(row, sRow, stepRow);
                        const float qr2 = CUPSS_FMUL(qr, qr);
                        ampA = noise_amplitude_q2(od0.noise, od0.noiseAmp0, CUPSS_FADD(baseA, qr2));
                        ampB = noise_amplitude_q2(od0.noise, od0.noiseAmp0, CUPSS_FADD(baseB, qr2));
                    }
                    float2 ga, gb = make_float2(0.0f, 0.0f);
                    if (!planes) {
                        ga = normal2_from_words(c[0], c[1]);
                        gb = normal2_from_words(c[2], c[3]);
                        ga.x *= ks.whitePair; ga.y *= ks.whitePair; gb.x *= ks.whitePair; gb.y *= ks.whitePair;
                    } else {
                        KPoint ka{}, kb{};
                        ka.ix = (int)col; kb.ix = (int)col + 1;
                        ka.iy = kb.iy = iy; ka.iz = kb.iz = iz;
                        ka.rnd0 = c[0]; ka.rnd1 = c[1]; kb.rnd0 = c[2]; kb.rnd1 = c[3];
                        ka.rndField = kb.rndField = od0.fieldId;
                        ga = white_noise_mode(ks, ka, od0.fieldId, step);
                        if (valid1) gb = white_noise_mode(ks, kb, od0.fieldId, step);
                    }
                    g = make_float4(CUPSS_FMUL(ampA, ga.x), CUPSS_FMUL(ampA, ga.y), CUPSS_FMUL(ampB, gb.x), CUPSS_FMUL(ampB, gb.y));
                }
                tile[(v * R + q) * CP + cp] = g;
            }
            load_self();
        }
        if constexpr (KIND == KS_SCALAR_Q2) {
            float2* dp = ks.dst[0] + off0;
            const double tp[3] = {ks.sq2.tpre[0], ks.sq2.tpre[1], ks.sq2.tpre[2]};
            const double ip[4] = {ks.sq2.ipre[0], ks.sq2.ipre[1], ks.sq2.ipre[2], ks.sq2.ipre[3]};
            const bool termFused = ks.sq2.termFused != 0;
            const float dt = ks.dt;
            // The pass runs along the whole last axis, so sRow == L and, with q unrolled, the sign of the mode number of
            // row f0 + (L/R) q, its self-conjugacy and its distance from zero are known per q at compile time up to f0.
            // (float)(f0 + c) == (float)f0 + (float)c exactly (integers below 2^24): one I2F per virtual thread.
            const float f0f = (float)(int)f0;
            const bool f00 = f0 == 0u;
#pragma unroll
            for (unsigned q = 0; q < R; ++q) {
                constexpr int STEP = L / R;
                const bool neg = L > 1 && (int)q * STEP >= (L + 1) / 2;            // wavenumber(): i < (s+1)/2 ? i : i - s
                const int row = (int)f0 + STEP * (int)q;
                const float qr = CUPSS_FMUL(CUPSS_FADD(f0f, (float)(STEP * (int)q - (neg ? L : 0))), stepRow);
                const float qr2 = CUPSS_FMUL(qr, qr);
                const float q2a = CUPSS_FADD(baseA, qr2);
                const float q2b = CUPSS_FADD(baseB, qr2);
                float2 va, vb;
                if constexpr (NOISE) {
                    const float4 g = tile[(v * R + q) * CP + cp];
                    va = kstage_point_scalar_q2_sig<SIG>(tp, ip, termFused, dt, q2a, x0[q], make_float2(self[q].x, self[q].y), make_float2(g.x, g.y));
                    vb = kstage_point_scalar_q2_sig<SIG>(tp, ip, termFused, dt, q2b, x1[q], make_float2(self[q].z, self[q].w), make_float2(g.z, g.w));
                } else if constexpr (SIG >= 0) {
                    va = kstage_point_scalar_q2_sig<SIG>(tp, ip, termFused, dt, q2a, x0[q], make_float2(self[q].x, self[q].y));
                    vb = kstage_point_scalar_q2_sig<SIG>(tp, ip, termFused, dt, q2b, x1[q], make_float2(self[q].z, self[q].w));
                } else {
                    va = kstage_point_scalar_q2(ks.sq2, dt, q2a, x0[q], make_float2(self[q].x, self[q].y));
                    vb = kstage_point_scalar_q2(ks.sq2, dt, q2b, x1[q], make_float2(self[q].z, self[q].w));
                }
                if (q == 0 || 2 * q * STEP == (unsigned)L) {   // rows 0 and L/2 (f0 == 0 only) are self-conjugate
                    if (fixSelfA && f00) va.y = 0.0f;
                    if (fixSelfB && f00) vb.y = 0.0f;
                }
                if (!valid1) vb = make_float2(0.0f, 0.0f);   // padding column of the pitch
                if (valid) st4(dp + q * rowStride, make_float4(va.x, va.y, vb.x, vb.y));
                const int nr = neg ? L - row : row;
                const bool keepR = nr <= cutRow;
                x0[q] = (keepA && keepR) ? va : make_float2(0.0f, 0.0f);
                x1[q] = (keepB && keepR) ? vb : make_float2(0.0f, 0.0f);
            }
        } else {
            // run-time compiled plan: the pointwise sources of the NEXT row are requested (one 128-bit load per source for the
            // column pair) before this row is evaluated -- the rolled loop otherwise pays one exposed load latency per row
            // (KPZ-3D 512^3 constraint sweep: 59 % of the stall samples on the first use of the state, profiles/README.md)
            constexpr int NSRC = PlanNsrc<PLAN>::value;
            float4 pre[NSRC > 0 ? NSRC : 1];
            auto fetch_sources = [&](unsigned q) {
                const unsigned o = kbase + (f0 + (unsigned)(L / R) * q) * krs;
#pragma unroll
                for (int i = 0; i < NSRC; ++i) pre[i] = valid ? ld4(ks.src[i] + o) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            };
            if constexpr (KIND == KS_JIT) fetch_sources(0);
#pragma unroll 1
            for (unsigned q = 0; q < R; ++q) {
                const int row = (int)(f0 + (L / R) * q);
                const int iy = a.axis == 2 ? iyFix : (a.axis == 1 ? row : 0);
                const int iz = a.axis == 2 ? row : 0;
                const long long off = (long long)(kbase + (unsigned)row * krs);
                float2 ra = make_float2(0.0f, 0.0f), rb = make_float2(0.0f, 0.0f);
                float2 sa[NSRC > 0 ? NSRC : 1], sb[NSRC > 0 ? NSRC : 1];
                if constexpr (KIND == KS_JIT) {
#pragma unroll
                    for (int i = 0; i < (NSRC > 0 ? NSRC : 1); ++i) { sa[i] = make_float2(pre[i].x, pre[i].y); sb[i] = make_float2(pre[i].z, pre[i].w); }
                    if (q + 1 < R) fetch_sources(q + 1);
                }
                KPoint ka = make_kpoint(ks, (int)col, iy, iz), kb = make_kpoint(ks, (int)col + 1, iy, iz);
                if (ks.noiseField >= 0 && valid) {   // one generator call for both columns of the pair
                    unsigned int c[4];
                    philox_pair(ks, (int)(col >> 1), iy, iz, (unsigned int)ks.noiseField, step, c);
                    ka.rnd0 = c[0]; ka.rnd1 = c[1]; kb.rnd0 = c[2]; kb.rnd1 = c[3];
                    ka.rndField = kb.rndField = ks.noiseField;
                }
                if constexpr (KIND == KS_JIT) {
                    // up to four outputs: the new values of both columns go out with one 128-bit store per output (KPZ-3D 512^3
                    // constraint sweep 0.70 -> 0.60 ms); more outputs cost more in registers than the stores save (Model H 2048^2,
                    // eight outputs: 0.102 -> 0.119 ms) and are stored mode by mode
                    constexpr int NOUT = PlanNout<PLAN>::value;
                    if constexpr (NOUT <= 4) {
                        float2 oa[NOUT > 0 ? NOUT : 1], ob[NOUT > 0 ? NOUT : 1];
#pragma unroll
                        for (int o = 0; o < NOUT; ++o) ob[o] = make_float2(0.0f, 0.0f);   // padding column of the pitch
                        if (valid) ra = kstage_point_plan_src<PLAN>(ks, ka, x0[0], off, step, sa, oa);
                        if (valid1) rb = kstage_point_plan_src<PLAN>(ks, kb, x1[0], off + 1, step, sb, ob);
                        if (valid) {
#pragma unroll
                            for (int o = 0; o < NOUT; ++o) st4(ks.dst[o] + off, make_float4(oa[o].x, oa[o].y, ob[o].x, ob[o].y));
                        }
                    } else {
                        if (valid) ra = kstage_point_plan_src<PLAN>(ks, ka, x0[0], off, step, sa);
                        if (valid1) rb = kstage_point_plan_src<PLAN>(ks, kb, x1[0], off + 1, step, sb);
                    }
                } else {
                    if (valid) ra = kstage_point(ks, ka, x0[0], off, step);
                    if (valid1) rb = kstage_point(ks, kb, x1[0], off + 1, step);
                }
                // rotate so that the loop body only ever touches register 0 and R-1 (rolled loop, static indices)
#pragma unroll
                for (unsigned i = 0; i + 1 < R; ++i) { x0[i] = x0[i + 1]; x1[i] = x1[i + 1]; }
                x0[R - 1] = ra; x1[R - 1] = rb;
            }
        }

        if (doInv) {
            level_butterfly2<S, LAST, +1, false>(x0, x1, 0, twS);
#pragma unroll
            for (unsigned q = 0; q < R; ++q) {
                const float4 t = make_float4(x0[q].x, x0[q].y, x1[q].x, x1[q].y);
                if constexpr (n > 1) tile[(v * R + q) * CP + cp] = t;
                else if (valid) st4(axis_dst(a, lbase, pbase, v * R + q), t);
            }
        }
    }

    // ---- inverse levels n-2 .. 0
    if constexpr (n > 1) {
        if (doInv) {
            auto gst = [&](unsigned pos, unsigned, float4 v) {
                if (valid) st4(axis_dst(a, lbase, pbase, pos), v);
            };
            __syncthreads();
            if constexpr (n >= 4) { tile_level<S, 2, +1, TV>(tv, twS, sld, sst); __syncthreads(); }
            if constexpr (n >= 3) { tile_level<S, 1, +1, TV>(tv, twS, sld, sst); __syncthreads(); }
            if constexpr (CL == 1) {
                tile_level<S, 0, +1, TV>(tv, twS, sld, gst);
            } else {
                tile_level<S, 0, +1, TV>(tv, twS, sld, sst);
                cluster_sync();   // every block holds its inverse 512-point transform
                cross_gather<L, CL, CP, TV>(tile, crank, tv, cp, a.twX, [&](unsigned row, float4 v) {
                    if (valid) st4(axis_dst(a, lbase, pbase, row), v);
                });
            }
        }
        if constexpr (CL > 1) cluster_sync();   // no CTA leaves while a peer may still read its block (doInv is cluster-uniform)
    }
}

template <int L, int KIND, int SIG>
__global__ void __launch_bounds__(AxisCfg<L>::THREADS, (AxisCfg<L>::MINB > 2 ? 2 : AxisCfg<L>::MINB))
axis_kstage_kernel(const __grid_constant__ AxisArgs a, const __grid_constant__ KStageD ks) {
    axis_kstage_body<L, KIND, SIG, void>(a, ks);
}
template <int L, int KIND, int SIG>
__global__ void __cluster_dims__(AxisCfg<L>::CL, 1, 1) __launch_bounds__(AxisCfg<L>::THREADS, (AxisCfg<L>::MINB > 2 ? 2 : AxisCfg<L>::MINB))
axis_kstage_cluster_kernel(const __grid_constant__ AxisArgs a, const __grid_constant__ KStageD ks) {
    axis_kstage_body<L, KIND, SIG, void>(a, ks);
}

}  // namespace cupss
