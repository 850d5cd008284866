// Host-side check of cupss_b200/csrc/fft_core.cuh: runs the SAME butterflies and Stockham index
// arithmetic the kernels use, thread by thread on the CPU, against a double-precision naive DFT.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../cupss_b200/csrc/fft_core.cuh"
using namespace cupss;

template <int L, int DIR>
double check() {
    using P = FftPlan<L>;
    constexpr int E = P::E, T = P::T, R0 = P::R0, R1 = P::R1, R2 = P::R2;
    std::vector<float2> tw(L), x(L), buf(L), y(L);
    for (int k = 0; k < L; ++k) tw[k] = make_float2((float)std::cos(-2.0 * kPi * k / L), (float)std::sin(-2.0 * kPi * k / L));
    for (int i = 0; i < L; ++i) x[i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    std::vector<std::vector<float2>> regs(T, std::vector<float2>(E));
    auto asarr = [&](int t) -> float2(&)[E] { return *reinterpret_cast<float2(*)[E]>(regs[t].data()); };
    for (int t = 0; t < T; ++t) for (int e = 0; e < E; ++e) regs[t][e] = x[t + T * e];
    for (int t = 0; t < T; ++t) stockham_pass<L, E, R0, 1, DIR>(asarr(t), t, tw.data());
    if (R1 > 1) {
        for (int t = 0; t < T; ++t) for (int e = 0; e < E; ++e) buf[stockham_out_index<L, E, R0, 1>(t, e)] = regs[t][e];
        for (int t = 0; t < T; ++t) for (int e = 0; e < E; ++e) regs[t][e] = buf[t + T * e];
        for (int t = 0; t < T; ++t) stockham_pass<L, E, R1, R0, DIR>(asarr(t), t, tw.data());
        if (R2 > 1) {
            for (int t = 0; t < T; ++t) for (int e = 0; e < E; ++e) buf[stockham_out_index<L, E, R1, R0>(t, e)] = regs[t][e];
            for (int t = 0; t < T; ++t) for (int e = 0; e < E; ++e) regs[t][e] = buf[t + T * e];
            for (int t = 0; t < T; ++t) stockham_pass<L, E, R2, R0 * R1, DIR>(asarr(t), t, tw.data());
        }
    }
    for (int t = 0; t < T; ++t) for (int e = 0; e < E; ++e) y[t + T * e] = regs[t][e];
    double err = 0, nrm = 0;
    for (int k = 0; k < L; ++k) {
        double re = 0, im = 0;
        for (int n = 0; n < L; ++n) {
            double a = DIR * 2.0 * kPi * (double)((long)k * n % L) / L;
            re += x[n].x * std::cos(a) - x[n].y * std::sin(a);
            im += x[n].x * std::sin(a) + x[n].y * std::cos(a);
        }
        err += (re - y[k].x) * (re - y[k].x) + (im - y[k].y) * (im - y[k].y);
        nrm += re * re + im * im;
    }
    return std::sqrt(err / nrm);
}

int main() {
    int bad = 0;
#define CHK(L) { double a = check<L, -1>(), b = check<L, 1>(); printf("L=%5d  fwd %.3e  inv %.3e\n", L, a, b); if (!(a < 5e-7 && b < 5e-7)) bad++; }
    CHK(1) CHK(2) CHK(4) CHK(8) CHK(16) CHK(32) CHK(64) CHK(128) CHK(256) CHK(512) CHK(1024) CHK(2048) CHK(4096) CHK(8192)
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
