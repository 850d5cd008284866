/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product path.
 *
 * The reference allocates device memory and builds cuFFT/cuRAND objects in
 * every field/term constructor even when RUN_CPU is selected
 * (/root/reference/src/field_init.cpp:20-43,70-137, src/term_init.cpp:15-16,33-60),
 * so its CPU path cannot start on a GPU-less host.  This translation unit
 * defines the handful of CUDA runtime / cuFFT / cuRAND entry points and the
 * extern "C" kernel launchers those sources reference, backed by plain host
 * memory (allocation, copies) or as loud failures (anything that would mean
 * the GPU branch of the reference was taken).  With it the UNMODIFIED
 * reference sources build with g++ alone and RUN_CPU runs anywhere.
 */
#include <cuda_runtime.h>
#include <cufft.h>
#include <curand.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cupss/defines.h" /* -I<reference>/inc supplied by oracle/Makefile */

static void gpu_branch_taken(const char *what)
{
    std::fprintf(stderr, "cupss oracle stub: %s called -- the oracle build only supports RUN_CPU\n", what);
    std::abort();
}

extern "C" {

cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = std::calloc(n ? n : 1, 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind)
{
    std::memcpy(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "cupss oracle stub"; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 0; return cudaSuccess; }
cudaError_t cudaRuntimeGetVersion(int *v) { *v = 0; return cudaSuccess; }
cudaError_t cudaDriverGetVersion(int *v) { *v = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp *p, int) { std::memset(p, 0, sizeof(*p)); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int) { *s = nullptr; return cudaSuccess; }

cufftResult cufftPlan1d(cufftHandle *h, int, cufftType, int) { *h = 0; return CUFFT_SUCCESS; }
cufftResult cufftPlan2d(cufftHandle *h, int, int, cufftType) { *h = 0; return CUFFT_SUCCESS; }
cufftResult cufftPlan3d(cufftHandle *h, int, int, int, cufftType) { *h = 0; return CUFFT_SUCCESS; }
cufftResult cufftDestroy(cufftHandle) { return CUFFT_SUCCESS; }
cufftResult cufftExecC2C(cufftHandle, cufftComplex *, cufftComplex *, int) { gpu_branch_taken("cufftExecC2C"); return CUFFT_EXEC_FAILED; }

curandStatus_t curandCreateGenerator(curandGenerator_t *g, curandRngType_t) { *g = nullptr; return CURAND_STATUS_SUCCESS; }
curandStatus_t curandSetStream(curandGenerator_t, cudaStream_t) { return CURAND_STATUS_SUCCESS; }
curandStatus_t curandSetGeneratorOffset(curandGenerator_t, unsigned long long) { return CURAND_STATUS_SUCCESS; }
curandStatus_t curandSetGeneratorOrdering(curandGenerator_t, curandOrdering_t) { return CURAND_STATUS_SUCCESS; }
curandStatus_t curandSetPseudoRandomGeneratorSeed(curandGenerator_t, unsigned long long) { return CURAND_STATUS_SUCCESS; }
curandStatus_t curandGenerateNormal(curandGenerator_t, float *, size_t, float, float) { gpu_branch_taken("curandGenerateNormal"); return CURAND_STATUS_LAUNCH_FAILURE; }

/* launchers declared in /root/reference/inc/cupss/field_kernels.cuh:7-19 and term_kernels.cuh:7-15 */
void setNotDynamic_gpu(float2 **, int, pres *, int, float2 *, int, int, int, float, float, float, float *, bool, float2 *, float *, float, dim3, dim3) { gpu_branch_taken("setNotDynamic_gpu"); }
void setDynamic_gpu(float2 **, int, pres *, int, float2 *, int, int, int, float, float, float, float, float *, bool, float2 *, float *, dim3, dim3) { gpu_branch_taken("setDynamic_gpu"); }
void normalize_gpu(float2 *, int, int, int, dim3, dim3) { gpu_branch_taken("normalize_gpu"); }
void createNoise_gpu(float *, float *, int, int, int, float *, dim3, dim3) { gpu_branch_taken("createNoise_gpu"); }
void dealias_gpu(float2 *, float2 *, int, int, int, int, dim3, dim3) { gpu_branch_taken("dealias_gpu"); }
void copyToFloat2_gpu(float *, float2 *, int, int, int, dim3, dim3) { gpu_branch_taken("copyToFloat2_gpu"); }
void correctNoiseAmplitude_gpu(float2 *, float *, int, int, int, dim3, dim3) { gpu_branch_taken("correctNoiseAmplitude_gpu"); }
void computeProduct_gpu(float2 **, float2 *, int, int, int, int, dim3, dim3) { gpu_branch_taken("computeProduct_gpu"); }
void applyPrefactor_gpu(float2 *, float, int, int, int, int, int, int, int, int, float, float, float, dim3, dim3) { gpu_branch_taken("applyPrefactor_gpu"); }
void copyComp_gpu(float2 *, float2 *, int, int, int, dim3, dim3) { gpu_branch_taken("copyComp_gpu"); }
void applyPres_vector_gpu(float2 *, pres *, int, int, int, int, float, float, float, dim3, dim3) { gpu_branch_taken("applyPres_vector_gpu"); }
void applyPres_vector_pre_gpu(float2 *, float *, int, int, int, int, dim3, dim3) { gpu_branch_taken("applyPres_vector_pre_gpu"); }

} /* extern "C" */
