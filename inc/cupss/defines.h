// defines.h -- plain data types shared by the public API.
// Layout-compatible with /root/reference/inc/cupss/defines.h (struct pres :31-39 is passed by value
// through evolver::createTerm, so its members and their order are part of the API).
#ifndef CUPSS_B200_DEFINES_H
#define CUPSS_B200_DEFINES_H

#include <cuda_runtime.h>   // float2, dim3 (the reference's users rely on this transitive include)

#include <iostream>
#include <map>
#include <string>
#include <vector>

// One monomial of a Fourier-space prefactor:
//   preFactor * q^(2*q2n) * (i qx)^iqx * (i qy)^iqy * (i qz)^iqz * |q|^(-invq)
struct pres {
    float preFactor = 0.0f;
    int q2n = 0;
    int iqx = 0;
    int iqy = 0;
    int iqz = 0;
    int invq = 0;
};

struct system_constants {
    int sx, sy, sz;
    float dx, dy, dz;
    float dt;
    int writeEveryNSteps;
};

struct full_term {
    std::vector<pres> prefactors;
    std::vector<std::string> fields;
};

// The reference's single-precision pi (defines.h:54); wavenumbers are built from it, so parity needs the same digits.
#define PI 3.1415926535f

// RUN_GPU: reference-GPU semantics (dealias_k mask, host mirrors refreshed on copyAllDataToHost/writeOut).
// RUN_CPU: reference-CPU semantics (the CPU loop's dealias mask, host mirrors live after every step) --
//          still executed by the B200 engine; this library has no CPU compute path.
enum DeviceToRunOn { RUN_CPU, RUN_GPU };
enum integrators { EULER, RK2, RK4 };
enum NoiseType { GaussianWhite };

#endif
