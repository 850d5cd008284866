// Host-side check of cupss_b200/csrc/fft_core.cuh: runs the SAME level butterflies, twiddles and index
// arithmetic the kernels use, virtual thread by virtual thread on the CPU, against a double-precision naive DFT.
//   forward: natural order in -> position p holds frequency freq_of_pos(p)
//   inverse: that order in    -> natural order out (unnormalised)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../cupss_b200/csrc/fft_core.cuh"
using namespace cupss;

template <int L, int LV, int SIGN, bool DIF>
void run_level(std::vector<float2>& buf, const float2* tw) {
    using G = LevelGeom<L, LV>;
    for (int v = 0; v < G::NV; ++v) {
        const int blk = v / G::M, j = v % G::M, row0 = blk * G::N + j;
        float2 x[G::R];
        for (int q = 0; q < G::R; ++q) x[q] = buf[row0 + G::M * q];
        level_butterfly<L, LV, SIGN, DIF>(x, j, tw);
        for (int q = 0; q < G::R; ++q) buf[row0 + G::M * q] = x[q];
    }
}
template <int L, int LV, int SIGN>
void dif_from(std::vector<float2>& buf, const float2* tw) {   // natural in -> digit-reversed out
    if constexpr (LV < FftLevels<L>::n) { run_level<L, LV, SIGN, true>(buf, tw); dif_from<L, LV + 1, SIGN>(buf, tw); }
}
template <int L, int LV, int SIGN>
void dit_from(std::vector<float2>& buf, const float2* tw) {   // digit-reversed in -> natural out
    if constexpr (LV >= 0) { run_level<L, LV, SIGN, false>(buf, tw); dit_from<L, LV - 1, SIGN>(buf, tw); }
}

template <int L>
void naive(const std::vector<float2>& x, int dir, std::vector<double>& re, std::vector<double>& im) {
    re.assign(L, 0); im.assign(L, 0);
    for (int k = 0; k < L; ++k)
        for (int n = 0; n < L; ++n) {
            const double a = dir * 2.0 * kPi * (double)((long)k * n % L) / L;
            re[k] += x[n].x * std::cos(a) - x[n].y * std::sin(a);
            im[k] += x[n].x * std::sin(a) + x[n].y * std::cos(a);
        }
}

template <int L>
bool check() {
    std::vector<float2> tw(TwTable<L>::LEN), x(L);
    fill_level_twiddles<L, 0>(tw.data());
    for (int i = 0; i < L; ++i) x[i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    std::vector<double> re, im;
    // permutation is a bijection
    std::vector<int> seen(L, 0);
    for (int p = 0; p < L; ++p) seen[freq_of_pos<L>(p)]++;
    for (int k = 0; k < L; ++k) if (seen[k] != 1) { printf("L=%d: freq_of_pos is not a permutation\n", L); return false; }
    double worst = 0;
    for (int sign = -1; sign <= 1; sign += 2) {
        naive<L>(x, sign, re, im);
        // DIF: natural in, position p holds frequency freq_of_pos(p)
        std::vector<float2> buf = x;
        if (sign < 0) dif_from<L, 0, -1>(buf, tw.data()); else dif_from<L, 0, +1>(buf, tw.data());
        double err = 0, nrm = 0;
        for (int p = 0; p < L; ++p) {
            const int k = freq_of_pos<L>(p);
            if ((int)pos_of_freq<L>(k) != p) { printf("L=%d: pos_of_freq is not the inverse of freq_of_pos\n", L); return false; }
            err += (re[k] - buf[p].x) * (re[k] - buf[p].x) + (im[k] - buf[p].y) * (im[k] - buf[p].y);
            nrm += re[k] * re[k] + im[k] * im[k];
        }
        const double e1 = std::sqrt(err / nrm);
        // DIT: input index n at position pos_of_freq(n), natural order out
        for (int nn = 0; nn < L; ++nn) buf[pos_of_freq<L>(nn)] = x[nn];
        if (sign < 0) dit_from<L, FftLevels<L>::n - 1, -1>(buf, tw.data()); else dit_from<L, FftLevels<L>::n - 1, +1>(buf, tw.data());
        err = 0; nrm = 0;
        for (int k = 0; k < L; ++k) {
            err += (re[k] - buf[k].x) * (re[k] - buf[k].x) + (im[k] - buf[k].y) * (im[k] - buf[k].y);
            nrm += re[k] * re[k] + im[k] * im[k];
        }
        const double e2 = std::sqrt(err / nrm);
        printf("L=%5d  sign %+d  DIF %.3e  DIT %.3e\n", L, sign, e1, e2);
        worst = std::max(worst, std::max(e1, e2));
    }
    return worst < 5e-7;
}

// PrunedDft<R0> (strided level of the x pass for input band-limited to sx/4) and the rotation folded into the twiddle
// multiplication: against the full in-register DFT of the zero-padded vector.
template <int R0>
bool check_pruned() {
    constexpr int Q = R0 / 4, M = 2 * R0;
    float2 lo[Q], hi[Q], full[R0], y[R0];
    std::vector<float2> tw((R0 - 1) * M);
    for (int q = 1; q < R0; ++q)
        for (int j = 0; j < M; ++j) {
            const double a = -2.0 * kPi * (double)(j * q) / (double)(R0 * M);
            tw[(q - 1) * M + j] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    double worst = 0;
    for (int trial = 0; trial < 8; ++trial) {
        for (int i = 0; i < R0; ++i) full[i] = make_float2(0.0f, 0.0f);
        for (int i = 0; i < Q; ++i) {
            lo[i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
            hi[i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
            full[i] = lo[i]; full[R0 - Q + i] = hi[i];
        }
        const int j = trial % M;
        PrunedDft<R0>::run(lo, hi, y);
        x3_twiddle_rot(y, tw.data() + j, cupss_std::make_integer_sequence<int, R0>{});
        Dft<R0, +1>::run(full);
        for (int q = 1; q < R0; ++q) full[q] = cmul_conj(full[q], tw[(q - 1) * M + j]);
        double err = 0, nrm = 0;
        for (int p = 0; p < R0; ++p) {
            err += (double)(y[p].x - full[p].x) * (y[p].x - full[p].x) + (double)(y[p].y - full[p].y) * (y[p].y - full[p].y);
            nrm += (double)full[p].x * full[p].x + (double)full[p].y * full[p].y;
        }
        worst = std::max(worst, std::sqrt(err / nrm));
    }
    printf("PrunedDft<%d> + rotated twiddles vs full DFT: %.3e\n", R0, worst);
    return worst < 5e-7;
}

int main() {
    int bad = 0;
    if (!check_pruned<16>()) bad++;
    if (!check_pruned<8>()) bad++;
#define CHK(L) if (!check<L>()) bad++;
    CHK(1) CHK(2) CHK(4) CHK(8) CHK(16) CHK(32) CHK(64) CHK(128) CHK(256) CHK(512) CHK(1024) CHK(2048) CHK(4096) CHK(8192)
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
