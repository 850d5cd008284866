// field.h -- one scalar field of the system.
// Public members follow /root/reference/inc/cupss/field.h:47-151 (examples and tests poke them directly);
// device state lives inside the engine plan (include/cupss_b200.h), not here.
#ifndef CUPSS_B200_FIELD_H
#define CUPSS_B200_FIELD_H

#include <map>
#include <string>
#include <vector>
#include "defines.h"

class term;
class evolver;

class field {
   public:
    field(int sx, float dx);
    field(int sx, int sy, float dx, float dy);
    field(int sx, int sy, int sz, float dx, float dy, float dz);
    ~field();

    std::string name;
    bool dynamic = false;
    bool isCUDA = true;
    bool outputToFile = false;
    int integrator = EULER;
    evolver *system_p = nullptr;
    int engine_id = -1;   // id inside the engine plan
    bool mirror_in_sync = false;   // real_array was uploaded and no step ran since: it IS the device state, exactly (what the
                                   // reference's real_array_d holds at that point), so a download must not round-trip it

    // host mirrors, float2[sz*sy*sx] row-major [z][y][x], value in .x
    float2 *real_array;
    float2 *comp_array;
    // device views (borrowed from the engine; valid after prepareProblem)
    float2 *real_array_d = nullptr;
    float2 *comp_array_d = nullptr;

    bool needsaliasing = false;
    int aliasing_order = 1;

    std::vector<term *> terms;
    std::vector<pres> implicit;
    std::map<std::string, int> usedParameters;   // parameters appearing in the implicit part

    bool isNoisy = false;
    NoiseType noiseType = GaussianWhite;
    pres noise_amplitude;

    // user callbacks (boundary conditions)
    bool hasCB = false;
    void (*callback)(evolver *, float2 *, int, int, int) = nullptr;
    bool hasCBFourier = false;
    void (*callbackFourier)(evolver *, float2 *, int, int, int) = nullptr;

    dim3 threads_per_block, blocks;

    // The reference's own transform entry points (/root/reference/inc/cupss/field.h:121-124), for user code that calls them between
    // steps.  The engine keeps spectra on the device and no real-space copy, so: toComp() = the spectrum becomes the forward
    // transform of the host real array (and the Fourier mirror follows); toReal() = the host real array becomes the inverse
    // transform of the device spectrum, already normalised -- normalize() has nothing left to do, nor has dealias() (the
    // dealiased copy is produced inside the fused k stage).
    void toReal();
    void toComp();
    void normalize();
    void dealias();

    void copyHostToDevice();
    void copyDeviceToHost();
    void copyRealHostToDevice();
    void copyRealDeviceToHost();
    void writeToFile(int currentTimeStep, int dim, int writePrecision);
    void prepareDevice();
    void precalculateImplicit(float dt);

    float getStepqx();
    float getStepqy();
    float getStepqz();

    int addImplicitString(const std::string &s);
    void printImplicitString();
    int updateParameter(const std::string &name, float value);

   private:
    const int sx, sy, sz;
    const float dx, dy, dz;
    const float stepqx, stepqy, stepqz;
    std::vector<std::string> implicit_prefactor_strings;
    bool real_pinned = false, comp_pinned = false;   // host mirrors are page-locked when a device is present
};

#endif
