// kernels.h -- host-visible launch interface of the sm_100a step kernels (internal to the engine).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#include "kstage.cuh"

namespace cupss {

constexpr int CUPSS_MAX_PEERS = 8;

// Addressing of one side (input or output) of a strided-axis pass.  Row r of batch b, column c:
//   b*bs + chunk(r)*cs + local(r)*rs + c          (float2 elements)
//   chunk(r) = (r >> rpcShift) & chunkMask,  local(r) = (r >> locShift) & rpcMask
// Natural layouts: one chunk (rpcShift = log2 L, locShift = 0, rpcMask = L-1).  The multi-GPU exchange layout
// [peer][z_loc][ky_loc][kx] of the y pass: chunk = the peer that owns row ky, local = its index there --
//   block distribution  (ky = peer*kyl + ky_loc):  rpcShift = log2 kyl, locShift = 0
//   cyclic distribution (ky = ky_loc*P + peer):    rpcShift = 0, chunkMask = P-1, locShift = log2 P
// (cyclic is the default: the rows a dealiased inverse transform keeps, |ky| <= cut, are then spread evenly over the ranks).
struct AxisAddr {
    long long bs, cs, rs;
    int rpcShift, rpcMask;
    int chunkMask, locShift;
};

struct AxisArgs {
    const float2* in;
    float2* out;
    AxisAddr ain, aout;
    int ncol;        // valid columns (sx/2 + 1)
    int ncolTiles;   // column tiles of THIS launch (ceil(ncol / C) unless the pass is split into column chunks)
    int ctBase;      // first column tile of this launch (column-chunked exchange pipeline, engine.cu)
    int nbatch;
    const float2* tw;    // level twiddle table of the transform a CTA runs (fft_core.cuh; 512 points when a cluster shares the axis)
    const float2* twX;   // cluster kernels: (CL-1) x 512 twiddles of the level that couples the CTAs' blocks, else null
    int axis;        // k index carried by the rows: 1 = ky, 2 = kz, 0 = none (L = 1)
    int kyBase, kyStride;   // axis == 2: iky = kyBase + batch * kyStride (block: rank*kyl, 1; cyclic: rank, P)
    int maskOn, cutx, cuty, cutz;   // plain inverse: dealias mask applied on load
    int sx, sy, sz;
    // Dealias-aware pruning of inverse transforms: a dealiased spectrum is zero outside |n_a| <= cut_a, so
    //  * a tile whose columns all lie beyond pruneCutX (or whose fixed ky lies beyond pruneCutY) is skipped
    //    entirely -- its output is never written and stays at its initial zero;
    //  * rows of the transformed axis beyond rowCut are known zeros and are not loaded (rowCut < 0: off).
    int pruneOn, pruneCutX, pruneCutY;
    int rowCut;
    // Fused slab exchange (multi-GPU): instead of `out`, row r of batch b is stored straight into the receive
    // buffer of the peer that owns it over NVLink:  push[peer(r)] + pushBase + b*pushBs + loc(r)*pushRs + col,
    // peer(r) = (r >> pushShift) & pushPeerMask, loc(r) = (r >> pushLocShift) & pushMask (same two distributions as AxisAddr).
    int pushOn, pushShift, pushMask;
    int pushPeerMask, pushLocShift;
    long long pushRs, pushBs, pushBase;
    float2* push[CUPSS_MAX_PEERS];   // base of every rank's exchange arena (header: flags / epochs / error word)
    int rank, nranks;
    // TMA tile prologue (cp.async.bulk.tensor, kernels_axis.cuh: tile_fetch_tma).  `tmap` is a CUtensorMap of the input array
    // seen as float32 [2*ncol][.][.][.] -- columns, then (row inside a chunk, chunk, batch) sorted by stride; tmaSlot[k] says
    // which of those three the k-th outer coordinate is.  One instruction moves a box of tmaBoxRows rows x 128 bytes straight
    // into the tile; columns beyond ncol are zero-filled by the hardware.  tmaOn = 0: per-thread cp.async path.
    int tmaOn, tmaBoxRows;
    int tmaSlot[3];
    alignas(64) unsigned long long tmap[16];
};
// Arena header layout (32-bit words from the arena base): flag table [pt][CUPSS_MAX_PEERS], epochs, error word.
constexpr int XH_FLAGS = 0, XH_EPOCH = 1024, XH_ERROR = 2048;

// Cross-GPU barrier after a pushed exchange: every rank bumps its epoch for exchange point `pt`, publishes it in
// flag[pt][rank] of every peer (release.sys) and waits until all peers have published theirs (acquire.sys).
struct XBarrier {
    unsigned int* flags[CUPSS_MAX_PEERS];   // base of each peer's flag table [pt][CUPSS_MAX_PEERS]
    unsigned int* epoch;                    // local epoch counters [pt]
    int* error;                             // local: set to 1 on timeout
    int* hostError;                         // the same word in mapped page-locked host memory: readable after the trap
    unsigned long long timeoutNs;           // 0: wait for ever
    int rank, nranks, pt;
};

// x pass: C2R of up to XP_MAX_IN half-spectrum lines, real-space products, R2C of the results
// (computeProduct_k + normalize_k + the x part of every cuFFT call; /root/reference/src/term_kernels.cu:48-70,
//  src/field_kernels.cu:115-128).
constexpr int XP_MAX_IN = 8;
constexpr int XP_MAX_OUT = 6;
constexpr int XP_MAX_MONO = 16;
constexpr int XP_MAX_FAC = 6;
struct XMono {
    float coef;
    signed char out, nfac;
    signed char fac[XP_MAX_FAC];
};
struct XArgs {
    const float2* in[XP_MAX_IN];
    float2* out[XP_MAX_OUT];
    XMono mono[XP_MAX_MONO];
    int nIn, nOut, nMono;
    int pitch;
    long long nlines;
    float norm;      // 1 / (sx*sy*sz)
    const float2* tw;
    const float2* tw3;     // strided-level twiddles of the two-level kernel (kernels_x3.cu); null: not available
    float* realOut;        // mode C2R_ONLY: [nlines][sx]
    const float* realIn;   // mode R2C_ONLY
    int jobsPerCta;
    int perJobFloat2;      // shared-memory float2 per job
    int kmax[XP_MAX_IN];   // input g is zero for kx > kmax[g] (dealias cut-off): those points are not loaded
};

enum XMode { X_HOT = 0, X_C2R_ONLY = 1, X_R2C_ONLY = 2 };

#ifndef __CUDACC_RTC__
// Launchers return cudaError_t of the launch; unsupported sizes return cudaErrorInvalidValue.
cudaError_t launch_axis_plain(int L, int dir, const AxisArgs& a, cudaStream_t st);
cudaError_t launch_axis_kstage(int L, const AxisArgs& a, const KStageD& ks, cudaStream_t st);
cudaError_t launch_xpass(int sx, int mode, XArgs& a, cudaStream_t st);
int xpass_max_inputs(int sx);   // inputs one launch can take (shared-memory stash)
cudaError_t launch_bump_counter(unsigned int* counter, cudaStream_t st);
// float <-> float2 views of a real field (user callbacks see float2[N] with the value in .x, like the reference)
cudaError_t launch_spectrum_compress(const float2* full, float2* half, int sx, int sy, int sz, int pitch, cudaStream_t st);
cudaError_t launch_real_expand(const float* in, float2* out, size_t n, cudaStream_t st);
cudaError_t launch_real_compress(const float2* in, float* out, size_t n, cudaStream_t st);
// Hermitian half spectrum -> planes [z0, z0 + zl) of the full spectrum float2[sz][sy][sx] (comp_array layout of the reference).
// `half` is [src][sz][kyl][pitch] with src = ky / kyl: one rank's array when kyl == sy, the all-gathered ky-slabs otherwise.
// cyclicP > 1: the ky rows are dealt out cyclically over cyclicP ranks (src = ky % cyclicP, row ky / cyclicP of its slab).
cudaError_t launch_spectrum_expand(const float2* half, float2* full, int sx, int sy, int sz, int pitch, int kyl, int z0, int zl, int cyclicP, cudaStream_t st);
cudaError_t launch_xgpu_barrier(const XBarrier& b, cudaStream_t st);
int axis_tile_cols(int L);          // C used for length L
// Launch geometry of the k-stage kernel for length L (for kernels compiled at run time)
bool axis_kstage_geometry(int L, int* threads, size_t* smem, int* minBlocks);
// Level twiddle table of an L-point transform (fft_core.cuh); returns the number of entries written (<= L), 0 if unsupported.
int host_level_twiddles(int L, float2* out);
// Long axes shared by a thread-block cluster (kernels_axis.cuh: AxisCfg): CTAs per cluster (1: none), rows per CTA, and the
// (CL-1) x rows table of the level that couples the CTAs' blocks (returns the number of entries, 0 without a cluster).
int axis_cluster_size(int L);
int axis_cta_rows(int L);
int host_cross_twiddles(int L, float2* out);
// Two-level x pass (kernels_x3.cu): sizes it covers, its twiddle table (<= sx entries) and its launcher.
bool xpass3_supported(int sx);
int host_x3_twiddles(int sx, float2* out);
cudaError_t launch_xpass3(int sx, XArgs& a, cudaStream_t st);
bool xpass3_sumpow_supported(int sx);
cudaError_t launch_xpass3_sumpow(int sx, XArgs& a, cudaStream_t st);
// three-level variant for sx = 1024, 2048, 4096 (kernels_x4.cu); reached through the xpass3 entry points
bool xpass4_supported(int sx);
int host_x4_twiddles(int sx, float2* out);
cudaError_t launch_xpass4(int sx, XArgs& a, cudaStream_t st);
cudaError_t launch_xpass4_sumpow(int sx, XArgs& a, cudaStream_t st);
// one-job stash x pass for long lines (kernels_xs.cu; sx = 1024, 2048, 4096): reached through launch_xpass
bool xstash1_supported(int sx);
int xstash1_max_inputs(int sx);
cudaError_t launch_xstash1(int sx, XArgs& a, cudaStream_t st);
bool fft_size_supported(int n);

#endif

}  // namespace cupss
