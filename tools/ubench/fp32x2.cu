// Microbenchmark: scalar vs packed (f32x2) FP32 issue rate on sm_100a.  Build: nvcc -arch=sm_100a -O3 fp32x2.cu -o fp32x2
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096, NACC = 8;
template <int MODE>
__global__ void k(float2* out, float2 s, float2 m) {
    float2 a[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) { a[i].x = __fadd_rn(a[i].x, s.x); a[i].y = __fadd_rn(a[i].y, s.y); }                  // 2 FADD
            if (MODE == 1) a[i] = __fadd2_rn(a[i], s);                                                                // 1 FADD2
            if (MODE == 2) { a[i].x = __fmaf_rn(a[i].x, m.x, s.x); a[i].y = __fmaf_rn(a[i].y, m.y, s.y); }       // 2 FFMA
            if (MODE == 3) a[i] = __ffma2_rn(a[i], m, s);                                                             // 1 FFMA2
            if (MODE == 4) { a[i].x = __fmul_rn(a[i].x, m.x); a[i].y = __fmul_rn(a[i].y, m.y); }                   // 2 FMUL
            if (MODE == 5) a[i] = __fmul2_rn(a[i], m);                                                                // 1 FMUL2
            if (MODE == 6) { float2 t = make_float2(a[i].x, a[i].x); a[i] = __ffma2_rn(t, m, a[(i + 1) % NACC]); }  // broadcast + FFMA2
        }
    }
    float2 r = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < NACC; ++i) { r.x += a[i].x; r.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, float2* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(out, make_float2(1e-3f, 2e-3f), make_float2(1.0001f, 0.9999f));
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, make_float2(1e-3f, 2e-3f), make_float2(1.0001f, 0.9999f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flopsLanes = (double)blocks * threads * ITER * NACC * 2;   // scalar-equivalent ops
    printf("%-28s %8.3f ms  %7.1f G scalar-op/s  (%.1f per clk per SM at 1.9 GHz)\n", name, ms, flopsLanes / ms / 1e6, flopsLanes / ms / 1e6 / 148 / 1.9);
}
int main() {
    float2* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    run<0>("FADD x2 (scalar)", out); run<1>("FADD2 (packed)", out);
    run<2>("FFMA x2 (scalar)", out); run<3>("FFMA2 (packed)", out);
    run<4>("FMUL x2 (scalar)", out); run<5>("FMUL2 (packed)", out);
    run<6>("bcast + FFMA2", out);
    cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
