// Host-side check of the noise generator in cupss_b200/csrc/kstage.cuh (everything there is __host__ __device__):
//   1. philox4x32_10 against the Random123 known-answer vectors (kat_vectors, "philox4x32 10");
//   2. white_noise_mode on whole grids: inside the self-conjugate planes kx = 0 and kx = sx/2 the value at (ky, kz) is the
//      complex conjugate of the value at (-ky, -kz), self-conjugate bins are real, and the second moments are
//      E|xi|^2 = N everywhere (N = sx*sy*sz; a self-conjugate bin carries it all in its real part);
//   3. Box-Muller: mean 0, variance 1, no correlation between the two outputs.
// Prints "OK" and exits 0 on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../cupss_b200/csrc/kstage.cuh"
using namespace cupss;

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++fails; } } while (0)

static void kat() {
    struct V { unsigned c[4], k[2], want[4]; };
    const V v[3] = {
        {{0u, 0u, 0u, 0u}, {0u, 0u}, {0x6627e8d5u, 0xe169c58du, 0xbc57ac4cu, 0x9b00dbd8u}},
        {{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu}, {0x408f276du, 0x41c83b0eu, 0xa20bc7c6u, 0x6d5451fdu}},
        {{0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u}, {0xa4093822u, 0x299f31d0u}, {0xd16cfe09u, 0x94fdccebu, 0x5001e420u, 0x24126ea1u}},
    };
    for (const V& t : v) {
        unsigned int c[4] = {t.c[0], t.c[1], t.c[2], t.c[3]};
        philox4x32_10(c, t.k[0], t.k[1]);
        for (int i = 0; i < 4; ++i) CHECK(c[i] == t.want[i], "philox4x32_10 word %d: %08x, Random123 says %08x", i, c[i], t.want[i]);
        std::printf("philox4x32_10 -> %08x %08x %08x %08x\n", c[0], c[1], c[2], c[3]);
        // the variant the step kernels call: round keys from a table (KStageD::philoxKey, in the constant bank on the device)
        unsigned int rk[20], d[4] = {t.c[0], t.c[1], t.c[2], t.c[3]};
        philox_round_keys((unsigned long long)t.k[0] | ((unsigned long long)t.k[1] << 32), rk);
        philox4x32_10_keys(d, rk);
        for (int i = 0; i < 4; ++i) CHECK(d[i] == t.want[i], "philox4x32_10_keys word %d: %08x, Random123 says %08x", i, d[i], t.want[i]);
    }
}

static KStageD make_ks(int sx, int sy, int sz, unsigned long long seed) {
    KStageD ks{};
    ks.sx = sx; ks.sy = sy; ks.sz = sz;
    ks.stepqx = ks.stepqy = ks.stepqz = 1.0f;
    ks.seed = seed;
    philox_round_keys(seed, ks.philoxKey);
    const float n = (float)sx * (float)sy * (float)sz;
    ks.whiteSelf = std::sqrt(n);
    ks.whitePair = std::sqrt(0.5f * n);
    ks.noiseField = -1;
    return ks;
}

static void hermitian(int sx, int sy, int sz) {
    const KStageD ks = make_ks(sx, sy, sz, 0x1234abcd5678ull);
    const double N = (double)sx * sy * sz;
    const int steps = 400;
    // accumulators per plane: second moments of ordinary and of self-conjugate bins
    double sumPair = 0, sumSelfRe = 0, sumBulk = 0;
    long nPair = 0, nSelf = 0, nBulk = 0;
    for (unsigned step = 0; step < (unsigned)steps; ++step) {
        for (int ix = 0; ix <= sx / 2; ++ix) {
            const bool plane = ix == 0 || 2 * ix == sx;
            for (int iz = 0; iz < sz; ++iz)
                for (int iy = 0; iy < sy; ++iy) {
                    const KPoint k = make_kpoint(ks, ix, iy, iz);
                    const float2 a = white_noise_mode(ks, k, 3, step);
                    if (!plane) { sumBulk += (double)a.x * a.x + (double)a.y * a.y; ++nBulk; continue; }
                    const int my = (sy - iy) % sy, mz = (sz - iz) % sz;
                    const KPoint km = make_kpoint(ks, ix, my, mz);
                    const float2 b = white_noise_mode(ks, km, 3, step);
                    CHECK(a.x == b.x && a.y == -b.y, "plane kx=%d: mode (%d,%d) = (%g,%g) but its mirror (%d,%d) = (%g,%g)", ix, iy, iz, a.x, a.y, my, mz, b.x, b.y);
                    if (my == iy && mz == iz) {
                        CHECK(a.y == 0.0f, "self-conjugate bin (%d,%d,%d) has imaginary part %g", ix, iy, iz, a.y);
                        sumSelfRe += (double)a.x * a.x; ++nSelf;
                    } else {
                        sumPair += (double)a.x * a.x + (double)a.y * a.y; ++nPair;
                    }
                }
        }
    }
    const double rBulk = sumBulk / nBulk / N, rPair = nPair ? sumPair / nPair / N : 1.0, rSelf = sumSelfRe / nSelf / N;
    std::printf("%dx%dx%d: E|xi|^2/N bulk %.4f (%ld), plane pairs %.4f (%ld), self-conjugate bins %.4f (%ld)\n", sx, sy, sz, rBulk, nBulk, rPair, nPair, rSelf, nSelf);
    CHECK(std::fabs(rBulk - 1) < 5.0 / std::sqrt((double)nBulk), "bulk variance %g", rBulk);
    if (nPair) CHECK(std::fabs(rPair - 1) < 5.0 / std::sqrt((double)nPair), "plane variance %g", rPair);
    CHECK(std::fabs(rSelf - 1) < 5.0 * std::sqrt(2.0 / (double)nSelf), "self-conjugate variance %g", rSelf);   // chi^2_1: relative sd sqrt(2)
    // a mode outside the planes is NOT tied to its (ky,kz) mirror
    if (sx >= 8) {
        const KPoint k = make_kpoint(ks, 1, 1 % sy, 1 % sz), km = make_kpoint(ks, 1, (sy - 1) % sy, (sz - 1) % sz);
        const float2 a = white_noise_mode(ks, k, 3, 0), b = white_noise_mode(ks, km, 3, 0);
        if (sy > 2 || sz > 2) CHECK(!(a.x == b.x && a.y == -b.y), "bulk modes must be independent of their in-plane mirrors");
    }
    // streams: other field id, other step, other seed -> other numbers; same inputs -> same numbers
    const KPoint k = make_kpoint(ks, 1, 0, 0);
    const float2 r0 = white_noise_mode(ks, k, 3, 7), r1 = white_noise_mode(ks, k, 3, 7), r2 = white_noise_mode(ks, k, 4, 7), r3 = white_noise_mode(ks, k, 3, 8);
    KStageD ks2 = ks; ks2.seed += 1; philox_round_keys(ks2.seed, ks2.philoxKey);
    const float2 r4 = white_noise_mode(ks2, k, 3, 7);
    CHECK(r0.x == r1.x && r0.y == r1.y, "not reproducible");
    CHECK(r0.x != r2.x && r0.x != r3.x && r0.x != r4.x, "field / step / seed do not separate the streams");
}

static void box_muller() {
    unsigned int c[4] = {0, 0, 0, 0};
    double s1 = 0, s2 = 0, sxy = 0, s4 = 0;
    const long n = 400000;
    for (long i = 0; i < n; ++i) {
        c[0] = (unsigned)i; c[1] = 0; c[2] = 42; c[3] = 0;
        philox4x32_10(c, 1u, 2u);
        const float2 g = normal2_from_words(c[0], c[1]);
        s1 += g.x + g.y; s2 += (double)g.x * g.x + (double)g.y * g.y; sxy += (double)g.x * g.y; s4 += std::pow((double)g.x, 4) + std::pow((double)g.y, 4);
    }
    const double mean = s1 / (2 * n), var = s2 / (2 * n), cov = sxy / n, kurt = s4 / (2 * n);
    std::printf("Box-Muller: mean %.5f var %.5f cov %.5f fourth moment %.4f\n", mean, var, cov, kurt);
    CHECK(std::fabs(mean) < 5.0 / std::sqrt(2.0 * n), "mean %g", mean);
    CHECK(std::fabs(var - 1) < 5.0 * std::sqrt(2.0 / (2.0 * n)), "variance %g", var);
    CHECK(std::fabs(cov) < 5.0 / std::sqrt((double)n), "covariance %g", cov);
    CHECK(std::fabs(kurt - 3) < 0.05, "fourth moment %g", kurt);
}

int main() {
    kat();
    hermitian(8, 8, 8);
    hermitian(16, 4, 2);
    hermitian(8, 16, 1);
    hermitian(32, 1, 1);
    box_muller();
    if (fails) { std::printf("%d check(s) failed\n", fails); return 1; }
    std::printf("OK\n");
    return 0;
}
