// field.cpp -- host side of a field: mirrors, CSV output, parameter updates.
// Device work of the reference's field.cpp/field_init.cpp (updateTerms, setRHS, toReal, toComp, dealias,
// normalize, createNoise, precalculateImplicit) is fused inside the engine; see cupss_b200/csrc/kstage.cuh.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../inc/cupss.h"

static float q_step(float d, int n) { return 2.0f * PI / (d * (float)n); }

// Host mirrors (float2[N], value in .x, as in the reference).  Page-locked when a CUDA device is present and the array is
// at most 2 GiB, so that the upload / download copies run at full PCIe speed; larger arrays (1024^3: 8 GiB each, and every
// rank of a partitioned run holds full-size mirrors but touches only its slab) come from calloc, i.e. untouched pages cost
// nothing.  A GPU-less host can still declare a system.
static float2 *alloc_mirror(size_t n, bool *pinned) {
    void *p = nullptr;
    *pinned = false;
    const size_t bytes = n * sizeof(float2);
    if (bytes >= (1u << 20) && bytes <= (2ull << 30) && cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) {
        *pinned = true;
        memset(p, 0, bytes);
        return static_cast<float2 *>(p);
    }
    (void)cudaGetLastError();
    p = calloc(n, sizeof(float2));
    if (!p) {
        std::cout << "ERROR: cannot allocate a host mirror of " << bytes << " bytes" << std::endl;
        std::exit(1);
    }
    return static_cast<float2 *>(p);
}
static void free_mirror(float2 *p, bool pinned) {
    if (pinned) cudaFreeHost(p); else free(p);
}

field::field(int nx, float hx)
    : sx(nx), sy(1), sz(1), dx(hx), dy(1.0f), dz(1.0f), stepqx(q_step(hx, nx)), stepqy(2.0f * PI), stepqz(2.0f * PI) {
    real_array = alloc_mirror((size_t)sx * sy * sz, &real_pinned);
    comp_array = alloc_mirror((size_t)sx * sy * sz, &comp_pinned);
}
field::field(int nx, int ny, float hx, float hy)
    : sx(nx), sy(ny), sz(1), dx(hx), dy(hy), dz(1.0f), stepqx(q_step(hx, nx)), stepqy(q_step(hy, ny)), stepqz(2.0f * PI) {
    real_array = alloc_mirror((size_t)sx * sy * sz, &real_pinned);
    comp_array = alloc_mirror((size_t)sx * sy * sz, &comp_pinned);
}
field::field(int nx, int ny, int nz, float hx, float hy, float hz)
    : sx(nx), sy(ny), sz(nz), dx(hx), dy(hy), dz(hz), stepqx(q_step(hx, nx)), stepqy(q_step(hy, ny)), stepqz(q_step(hz, nz)) {
    real_array = alloc_mirror((size_t)sx * sy * sz, &real_pinned);
    comp_array = alloc_mirror((size_t)sx * sy * sz, &comp_pinned);
}

field::~field() {
    free_mirror(real_array, real_pinned);
    free_mirror(comp_array, comp_pinned);
    for (term *t : terms) delete t;
}

float field::getStepqx() { return stepqx; }
float field::getStepqy() { return stepqy; }
float field::getStepqz() { return stepqz; }

// The engine owns device state; these keep the reference's names for user code that calls them.
// Mid-run edits of the host real array reach the device through these, as in the reference; the engine keeps the
// spectrum, so both spell "forward-transform my real array" (the reference's comp_array copy is implied by it).
void field::copyHostToDevice() { if (system_p) system_p->uploadHostMirror(this); }
void field::copyRealHostToDevice() { if (system_p) system_p->uploadHostMirror(this); }
void field::copyDeviceToHost() { if (system_p) system_p->refreshHostMirror(this, true, true); }
void field::copyRealDeviceToHost() { if (system_p) system_p->refreshHostMirror(this, true, false); }
// field::toComp / toReal / normalize / dealias of the reference (src/field.cpp:247-298, 203-232) as user-callable entry points;
// inside a step all four are fused into the engine's passes.
void field::toComp() {
    if (!system_p) return;
    system_p->uploadHostMirror(this);
    system_p->refreshHostMirror(this, false, true);
}
void field::toReal() { if (system_p) system_p->refreshHostMirror(this, true, false); }
void field::normalize() {}
void field::dealias() {}
void field::prepareDevice() { for (term *t : terms) t->prepareDevice(); }
void field::precalculateImplicit(float) { /* implicit and noise factors are evaluated in-kernel from the mode index */ }

// data/<name>.csv.<step>: header "x, [y, [z, ]]<name>", rows "%i, [%i, [%i, ]]%.<precision>f"
// (format of /root/reference/src/field.cpp:350-402; NaN aborts with exit(1)).
void field::writeToFile(int step, int dim, int precision) {
    if (!outputToFile) return;
    if (system_p) system_p->refreshHostMirror(this, true, false, /*keep_exact=*/true);
    const std::string path = "data/" + name + ".csv." + std::to_string(step);
    FILE *fp = std::fopen(path.c_str(), "w+");
    if (!fp) {
        std::cout << "Error creating output file at timestep" << step << std::endl;
        std::exit(1);
    }
    std::fprintf(fp, dim == 1 ? "x, %s\n" : (dim == 2 ? "x, y, %s\n" : "x, y, z, %s\n"), name.c_str());
    const std::string vf = "%." + std::to_string(precision) + "f\n";
    for (int k = 0; k < sz; k++)
        for (int j = 0; j < sy; j++)
            for (int i = 0; i < sx; i++) {
                const float v = real_array[((size_t)k * sy + j) * sx + i].x;
                if (std::isnan(v)) {
                    std::cout << "NaN found in field " << name << std::endl;
                    std::exit(1);
                }
                int w = std::fprintf(fp, "%i, ", i);
                if (dim >= 2 && w >= 0) w = std::fprintf(fp, "%i, ", j);
                if (dim >= 3 && w >= 0) w = std::fprintf(fp, "%i, ", k);
                if (w >= 0) w = std::fprintf(fp, vf.c_str(), v);
                if (w < 0) {
                    std::cout << "Error writing data to output file " << path << std::endl;
                    std::exit(1);
                }
            }
    std::fclose(fp);
}

int field::addImplicitString(const std::string &s) {
    implicit_prefactor_strings.push_back(s);
    return 0;
}

void field::printImplicitString() {
    for (const std::string &s : implicit_prefactor_strings) std::cout << s << std::endl;
    for (const auto &kv : usedParameters) std::cout << kv.first << " " << kv.second << std::endl;
}

// Re-parse the stored prefactor strings of everything that mentions `name`
// (/root/reference/src/field.cpp:422-441); the evolver re-bakes the plan constants afterwards.
int field::updateParameter(const std::string &name, float) {
    if (usedParameters[name]) system_p->_parser->recalculateImplicits(implicit_prefactor_strings, implicit, dynamic ? 1 : 0);
    for (term *t : terms)
        if (t->usedParameters[name]) system_p->_parser->recalculateImplicits(t->prefactor_strings, t->prefactors_h, 0);
    return 0;
}
