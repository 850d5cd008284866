// user_entry_check.cpp -- user code that calls the reference's field::toReal / normalize / toComp / dealias between steps
// (INTEGRATION.md).  Diffusion of a droplet: after toReal() the host real array holds the field; adding a constant to it and
// calling toComp() must show up in the zero mode of the Fourier mirror and survive further steps (the mean is conserved).
// Build: g++ -std=c++17 -I inc -I /usr/local/cuda/include tools/ubench/user_entry_check.cpp -L lib -lcupss -Wl,-rpath,$PWD/lib ...
#include <cupss.h>

#include <cmath>
#include <cstdio>

int main() {
    const int n = 64;
    evolver system(RUN_GPU, n, n, 1.0f, 1.0f, 0.1f, 10);
    system.createField("phi", true);
    system.addEquation("dt phi + q^2*phi = 0");
    system.initializeDroplet("phi", 0.0f, 1.0f, 10.0f, 2.0f, n / 2, n / 2, 0);
    system.prepareProblem();
    for (int i = 0; i < 5; ++i) system.advanceTime();
    field *f = system.fields[0];
    f->toReal();
    f->normalize();
    double mean0 = 0;
    for (int i = 0; i < n * n; ++i) mean0 += f->real_array[i].x;
    mean0 /= n * n;
    for (int i = 0; i < n * n; ++i) f->real_array[i].x += 0.25f;
    f->toComp();
    f->dealias();
    const double zero_mode = f->comp_array[0].x / (n * n);
    for (int i = 0; i < 5; ++i) system.advanceTime();
    f->toReal();
    double mean1 = 0;
    for (int i = 0; i < n * n; ++i) mean1 += f->real_array[i].x;
    mean1 /= n * n;
    std::printf("mean before %.6f, zero mode after toComp %.6f, mean after 5 more steps %.6f\n", mean0, zero_mode, mean1);
    const bool ok = std::fabs(zero_mode - (mean0 + 0.25)) < 1e-5 && std::fabs(mean1 - (mean0 + 0.25)) < 1e-5 && mean0 > 0.0;
    std::printf(ok ? "USER_ENTRY_OK\n" : "USER_ENTRY_FAILED\n");
    return ok ? 0 : 1;
}
