// host_expand.cu -- is it worth sending real fields over PCIe as packed floats and expanding them to the reference's float2
// host arrays (value in .x, zero in .y) with host threads?  Measures, for one 512^3 field:
//   (a) D2H of the float2 array as it is done now (1 GiB, one cudaMemcpyAsync from device to pinned host memory);
//   (b) D2H of the packed floats (0.5 GiB) in chunks into a pinned staging buffer, each chunk expanded into the float2 array by
//       T host threads while the next chunk is in flight;
//   (c) the mirror image for uploads: host threads pack chunks of the float2 array into the staging buffer, H2D of 0.5 GiB.
// Build: nvcc -O3 -std=c++17 -o host_expand host_expand.cu -lpthread      Run: ./host_expand [threads] [chunk MiB]
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <algorithm>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); return 1; } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void parallel(int T, size_t n, const std::function<void(size_t, size_t)>& f) {
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([=, &f] { f(n * t / T, n * (t + 1) / T); });
    for (auto& x : th) x.join();
}

int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 14;
    const size_t chunkMiB = argc > 2 ? atoi(argv[2]) : 32;
    const size_t N = 512ull * 512 * 512;
    float2* h2; float* hs; float2* d2; float* d1;
    CK(cudaMallocHost(&h2, N * sizeof(float2)));
    CK(cudaMallocHost(&hs, N * sizeof(float)));
    CK(cudaMalloc(&d2, N * sizeof(float2)));
    CK(cudaMalloc(&d1, N * sizeof(float)));
    CK(cudaMemset(d2, 0, N * sizeof(float2)));
    CK(cudaMemset(d1, 0, N * sizeof(float)));
    for (size_t i = 0; i < N; ++i) { h2[i].x = 1.0f; h2[i].y = 0.0f; }
    for (size_t i = 0; i < N; ++i) hs[i] = 1.0f;
    cudaStream_t st; CK(cudaStreamCreate(&st));
    for (int rep = 0; rep < 3; ++rep) {
        double t0 = now();
        CK(cudaMemcpyAsync(h2, d2, N * sizeof(float2), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const double ta = now() - t0;
        // (b) chunked packed D2H + threaded expansion
        const size_t ce = chunkMiB * 1024 * 1024 / sizeof(float);
        const size_t nch = (N + ce - 1) / ce;
        std::vector<cudaEvent_t> ev(nch);
        for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        t0 = now();
        for (size_t c = 0; c < nch; ++c) {
            const size_t o = c * ce, n = std::min(ce, N - o);
            CK(cudaMemcpyAsync(hs + o, d1 + o, n * sizeof(float), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(ev[c], st));
        }
        for (size_t c = 0; c < nch; ++c) {
            CK(cudaEventSynchronize(ev[c]));
            const size_t o = c * ce, n = std::min(ce, N - o);
            parallel(T, n, [&](size_t b, size_t e) { for (size_t i = o + b; i < o + e; ++i) { h2[i].x = hs[i]; h2[i].y = 0.0f; } });
        }
        const double tb = now() - t0;
        // (c) threaded packing + chunked H2D
        t0 = now();
        for (size_t c = 0; c < nch; ++c) {
            const size_t o = c * ce, n = std::min(ce, N - o);
            parallel(T, n, [&](size_t b, size_t e) { for (size_t i = o + b; i < o + e; ++i) hs[i] = h2[i].x; });
            CK(cudaMemcpyAsync(d1 + o, hs + o, n * sizeof(float), cudaMemcpyHostToDevice, st));
        }
        CK(cudaStreamSynchronize(st));
        const double tc = now() - t0;
        t0 = now();
        CK(cudaMemcpyAsync(d2, h2, N * sizeof(float2), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        const double td = now() - t0;
        printf("rep %d threads %d chunk %zu MiB: D2H float2 %.1f ms | packed D2H + expand %.1f ms | H2D float2 %.1f ms | pack + packed H2D %.1f ms\n", rep, T, chunkMiB,
               1e3 * ta, 1e3 * tb, 1e3 * td, 1e3 * tc);
        for (auto& e : ev) cudaEventDestroy(e);
    }
    return 0;
}
