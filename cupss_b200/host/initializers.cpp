// initializers.cpp -- host-side initial-condition helpers (write real_array[].x before prepareProblem).
// Formulas follow /root/reference/src/initializers.cpp; the droplet/half-system profiles are pinned by
// tests/base_truths/phi_{1,2,3}d.
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <fstream>

#include "../../inc/cupss.h"

namespace {
template <class F>
void for_each_site(int sx, int sy, int sz, F f) {
    for (int k = 0; k < sz; k++)
        for (int j = 0; j < sy; j++)
            for (int i = 0; i < sx; i++) f(i, j, k, ((size_t)k * sy + j) * sx + i);
}
}  // namespace

void evolver::initializeUniform(std::string name, float value) {
    float2 *a = findField(name, "initialize uniform")->real_array;
    for_each_site(sx, sy, sz, [&](int, int, int, size_t idx) { a[idx].x = value; });
}

void evolver::initializeUniformNoise(std::string name, float amplitude) {
    float2 *a = findField(name, "initialize uniform")->real_array;
    std::srand(time(0));
    for_each_site(sx, sy, sz, [&](int, int, int, size_t idx) { a[idx].x = amplitude * 0.01f * (float)(std::rand() % 200 - 100); });
}

void evolver::initializeNormalNoise(std::string name, float mean, float sigma) {
    float2 *a = findField(name, "initialize uniform")->real_array;
    std::srand(time(0));
    const size_t n = (size_t)sx * sy * sz;
    // Box-Muller on pairs of sites; the reference scales by sigma*sigma (src/initializers.cpp:89-90) -- kept.
    for (size_t idx = 0; idx < n; idx += 2) {
        const float u1 = 0.01 * (float)(std::rand() % 100 + 1), u2 = 0.01 * (float)(std::rand() % 100 + 1);
        const double r = sigma * sigma * std::sqrt(-2.0 * std::log(u1));
        a[idx].x = r * std::cos(2.0 * PI * u2) + mean;
        if (idx + 1 < n) a[idx + 1].x = r * std::sin(2.0 * PI * u2) + mean;
    }
}

void evolver::initializeHalfSystem(std::string name, float v1, float v2, float xi, int direction) {
    float2 *a = findField(name, "initialize half system")->real_array;
    if (xi <= 0.0) {
        std::cout << "ERROR in initialize, interface width cannot be 0 or negative" << std::endl;
        std::exit(1);
    }
    if (direction < 1 || direction > 3) {
        std::cout << "ERROR in initialize, direction can be 1, 2 or 3 for x, y, z, respectively" << std::endl;
        std::exit(1);
    }
    const int extent = direction == 1 ? sx : (direction == 2 ? sy : sz);
    for_each_site(sx, sy, sz, [&](int i, int j, int k, size_t idx) {
        const int ref = direction == 1 ? i : (direction == 2 ? j : k);
        a[idx].x = v1 + (v2 - v1) * 0.5 * (1.0 + std::tanh((ref - extent / 2) / (std::sqrt(2) * xi)));
    });
}

void evolver::initializeDroplet(std::string name, float v_out, float v_in, float radius, float xi, int cx, int cy, int cz) {
    float2 *a = findField(name, "initialize droplet")->real_array;
    if (xi <= 0.0 || radius <= 0.0) {
        std::cout << "ERROR in initialize, droplet radius and interface width cannot be 0 or negative" << std::endl;
        std::exit(1);
    }
    const int x0 = cx % sx, y0 = cy % sy, z0 = cz % sz;
    for_each_site(sx, sy, sz, [&](int i, int j, int k, size_t idx) {
        const float ddx = i - x0, ddy = j - y0, ddz = k - z0;
        const float r = std::sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        a[idx].x = v_out + (v_in - v_out) * 0.5 * (1.0 + std::tanh((r - radius) / (std::sqrt(2) * xi)));
    });
}

void evolver::addDroplet(std::string name, float value, float radius, float xi, int cx, int cy, int cz) {
    float2 *a = findField(name, "initialize droplet")->real_array;
    if (xi <= 0.0 || radius <= 0.0) {
        std::cout << "ERROR in initialize, droplet radius and interface width cannot be 0 or negative" << std::endl;
        std::exit(1);
    }
    const int x0 = cx % sx, y0 = cy % sy, z0 = cz % sz;
    for_each_site(sx, sy, sz, [&](int i, int j, int k, size_t idx) {
        const float ddx = i - x0, ddy = j - y0, ddz = k - z0;
        const float r = std::sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        a[idx].x += value * 0.5 * (1.0 + std::tanh((radius - r) / (std::sqrt(2) * xi)));
    });
}

// Rows "x[, y[, z]], value" as written by field::writeToFile; `skiprows` header lines are dropped.
void evolver::initializeFromFile(std::string name, std::string file, int skiprows, char delimiter) {
    float2 *a = findField(name, "initialize droplet")->real_array;
    std::ifstream in(file.c_str());
    if (!in.good()) {
        std::cout << "Initial file: " << file << " not found. Doing nothing" << std::endl;
        std::exit(1);
    }
    std::string tok;
    for (int i = 0; i < skiprows; i++) std::getline(in, tok);
    const size_t n = (size_t)sx * sy * sz;
    for (size_t row = 0; row < n; row++) {
        std::string px, py = "0", pz = "0", val;
        if (!std::getline(in, px, delimiter)) {
            std::cout << "Incompatible number of lines on file " << file << ", only " << row + 1 << " found for a system size of " << n << std::endl;
            std::exit(1);
        }
        if (dimension > 1) std::getline(in, py, delimiter);
        if (dimension > 2) std::getline(in, pz, delimiter);
        std::getline(in, val);
        const size_t idx = ((size_t)std::stoi(pz) * sy + std::stoi(py)) * sx + std::stoi(px);
        a[idx].x = std::stof(val);
    }
}
