// p2p_push.cu -- which store pattern moves a [rows x 128 B] tile fastest into a PEER GPU's memory over NVLink?
// (design input for the pushed slab exchange of kernels_axis.cuh; one process, 2 GPUs with peer access)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o p2p_push p2p_push.cu && ./p2p_push
// Modes (every CTA owns tiles [512 rows][C cols] of a [nb][512][272] float2 array, source = shared memory):
//   0  st.global.v4 from registers, 16-column tiles (128-byte row segments)      -- what the engine does today
//   1  same, 32-column tiles (256-byte row segments)
//   2  cp.async.bulk shared -> global, one 128-byte row segment per instruction
//   3  cp.async.bulk, 256-byte row segments (32-column tiles)
//   4  contiguous st.global.v4 (every warp 512 contiguous bytes): upper bound of SM stores
//   5  cp.async.bulk of contiguous 8 KB pieces: upper bound of the bulk path
// Each mode is timed one-directional (GPU0 -> GPU1) and bidirectional (both at once, per-direction rate reported).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int L = 512, PITCH = 272;

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256) push_kernel(float2* __restrict__ dst, int nb, int reps) {
    extern __shared__ float4 tile[];
    constexpr int C = (MODE == 1 || MODE == 3) ? 32 : 16;
    constexpr int CP = C / 2;                     // float4 per row
    constexpr int NT = PITCH / C + (PITCH % C ? 1 : 0);
    for (int i = threadIdx.x; i < L * CP; i += 256) tile[i] = make_float4(i, 1.0f, 2.0f, 3.0f);
    __syncthreads();
    if (MODE >= 2 && MODE != 4) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int ntiles = nb * NT;
    for (int rep = 0; rep < reps; ++rep)
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int b = t / NT, ct = t % NT;
        float2* base = dst + (size_t)b * L * PITCH + ct * C;
        if (MODE == 0 || MODE == 1) {
            const int cp = threadIdx.x % CP, tv = threadIdx.x / CP;
            if (ct * C + 2 * cp < PITCH)
                for (int r = tv; r < L; r += 256 / CP) {
                    // digit-reversed-like row order: consecutive row groups of a warp go to different slabs
                    const int row = ((r & 7) << 6) | (r >> 3);
                    *reinterpret_cast<float4*>(base + (size_t)row * PITCH + 2 * cp) = tile[r * CP + cp];
                }
        } else if (MODE == 2 || MODE == 3) {
            const int valid = PITCH - ct * C < C ? PITCH - ct * C : C;
            for (int r = threadIdx.x; r < L; r += 256) {
                const int row = ((r & 7) << 6) | (r >> 3);
                bulk_store(base + (size_t)row * PITCH, tile + r * CP, valid * 8);
            }
            bulk_commit_wait();
        } else if (MODE == 4) {
            float4* d4 = reinterpret_cast<float4*>(dst) + (size_t)t * (L * 8);
            for (int i = threadIdx.x; i < L * 8; i += 256) d4[i] = tile[i];
        } else {
            float4* d4 = reinterpret_cast<float4*>(dst) + (size_t)t * (L * 8);
            if (threadIdx.x < 8) bulk_store(d4 + threadIdx.x * 512, tile + threadIdx.x * 512, 8192);
            bulk_commit_wait();
        }
    }
}

template <int MODE>
static void run(float2* d01, float2* d10, int nb, cudaStream_t s0, cudaStream_t s1, int ctasPerSm) {
    constexpr int C = (MODE == 1 || MODE == 3) ? 32 : 16;
    const size_t smem = (size_t)L * (C / 2) * 16;
    const int grid = 148 * ctasPerSm;
    cudaEvent_t e0, e1, f0, f1;
    for (int bidir = 0; bidir < 2; ++bidir) {
        CK(cudaSetDevice(0));
        CK(cudaFuncSetAttribute(push_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaSetDevice(1));
        CK(cudaFuncSetAttribute(push_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
        const int reps = 4;
        for (int it = 0; it < 2; ++it) {   // warm-up, then timed
            CK(cudaSetDevice(0)); CK(cudaEventRecord(e0, s0));
            push_kernel<MODE><<<grid, 256, smem, s0>>>(d01, nb, reps);
            CK(cudaEventRecord(e1, s0));
            if (bidir) {
                CK(cudaSetDevice(1)); CK(cudaEventRecord(f0, s1));
                push_kernel<MODE><<<grid, 256, smem, s1>>>(d10, nb, reps);
                CK(cudaEventRecord(f1, s1));
            }
            CK(cudaSetDevice(0)); CK(cudaStreamSynchronize(s0));
            CK(cudaSetDevice(1)); CK(cudaStreamSynchronize(s1));
        }
        float ms0 = 0, ms1 = 0;
        CK(cudaEventElapsedTime(&ms0, e0, e1));
        if (bidir) CK(cudaEventElapsedTime(&ms1, f0, f1));
        const double bytes = (double)nb * L * PITCH * 8.0 * reps;
        printf("mode %d ctas/SM %d %s: %.1f GB/s", MODE, ctasPerSm, bidir ? "bidir" : "unidir", bytes / ms0 * 1e-6);
        if (bidir) printf("  | reverse %.1f GB/s", bytes / ms1 * 1e-6);
        printf("\n");
    }
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    const int nb = 256;
    const size_t bytes = (size_t)nb * L * PITCH * 8 + (1 << 20);
    float2 *b0, *b1;
    cudaStream_t s0, s1;
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&b0, bytes)); CK(cudaStreamCreate(&s0));
    CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&b1, bytes)); CK(cudaStreamCreate(&s1));
    for (int c : {1, 3}) {
        run<0>(b1, b0, nb, s0, s1, c);
        run<2>(b1, b0, nb, s0, s1, c);
        run<4>(b1, b0, nb, s0, s1, c);
        run<5>(b1, b0, nb, s0, s1, c);
    }
    run<1>(b1, b0, nb, s0, s1, 1);
    run<3>(b1, b0, nb, s0, s1, 1);
    // local reference: the same kernels writing into the GPU's own memory
    CK(cudaSetDevice(0));
    printf("local (GPU0 -> GPU0):\n");
    run<0>(b0, b1, nb, s0, s1, 3);
    return 0;
}
