/*
 * Flat C facade over the public `evolver` API -- see cupss_capi.h.
 * Compiled against EITHER this repo's inc/cupss.h (product) OR the reference's
 * (oracle build); only public members/methods are used
 * (/root/reference/inc/cupss/evolver.h:31-85, inc/cupss/field.h:47-106,
 *  inc/cupss/term.h:38-60).
 */
#include "cupss_capi.h"

#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "cupss.h"

static evolver *EV(void *p) { return static_cast<evolver *>(p); }

extern "C" {

void *cupss_capi_create(int with_cuda, int sx, int sy, int sz, float dx, float dy, float dz, float dt, int write_every)
{
    bool dev = with_cuda ? RUN_GPU : RUN_CPU;
    /* the three public constructors fix the dimension (evolver.h:31-33) */
    if (sz == 1 && sy == 1) return new evolver(dev, sx, dx, dt, write_every);
    if (sz == 1) return new evolver(dev, sx, sy, dx, dy, dt, write_every);
    return new evolver(dev, sx, sy, sz, dx, dy, dz, dt, write_every);
}

void cupss_capi_destroy(void *ev) { delete EV(ev); }

int cupss_capi_create_field(void *ev, const char *name, int dynamic)
{
    int r = EV(ev)->createField(name, dynamic != 0);
    /* field::integrator is never initialised by the reference (inc/cupss/field.h:53, SURVEY.md 8c hazard 7);
     * a user can and must set the public member, otherwise the CPU path may print "RK2 not implemented"
     * instead of stepping. */
    if (r == 0) EV(ev)->fieldsMap[name]->integrator = EULER;
    return r;
}
int cupss_capi_add_parameter(void *ev, const char *name, float value) { return EV(ev)->addParameter(name, value); }
int cupss_capi_add_equation(void *ev, const char *equation) { return EV(ev)->addEquation(equation); }
int cupss_capi_add_noise(void *ev, const char *field, const char *expr) { return EV(ev)->addNoise(field, expr); }

int cupss_capi_create_term(void *ev, const char *field, const cupss_capi_pres *p, int npres, const char *const *product, int nproduct)
{
    std::vector<pres> pv;
    for (int i = 0; i < npres; i++) {
        pres q;
        q.preFactor = p[i].preFactor; q.q2n = p[i].q2n; q.iqx = p[i].iqx;
        q.iqy = p[i].iqy; q.iqz = p[i].iqz; q.invq = p[i].invq;
        pv.push_back(q);
    }
    std::vector<std::string> prod;
    for (int i = 0; i < nproduct; i++) prod.push_back(product[i]);
    return EV(ev)->createTerm(field, pv, prod);
}

int cupss_capi_create_from_file(void *ev, const char *path) { return EV(ev)->createFromFile(path); }
void cupss_capi_prepare_problem(void *ev) { EV(ev)->prepareProblem(); }

int cupss_capi_advance_time(void *ev, int nsteps)
{
    evolver *e = EV(ev);
    for (int i = 0; i < nsteps; i++) e->advanceTime();
    return 0;
}

void cupss_capi_copy_all_data_to_host(void *ev) { EV(ev)->copyAllDataToHost(); }
void cupss_capi_write_out(void *ev) { EV(ev)->writeOut(); }
void cupss_capi_set_output_field(void *ev, const char *name, int on) { EV(ev)->setOutputField(name, on); }
int cupss_capi_update_parameter(void *ev, const char *name, float value) { return EV(ev)->updateParameter(name, value); }
float cupss_capi_get_parameter(void *ev, const char *name) { return EV(ev)->getParameter(name); }
int cupss_capi_get_timestep(void *ev) { return EV(ev)->getCurrentTimestep(); }
float cupss_capi_get_time(void *ev) { return EV(ev)->getCurrentTime(); }
void cupss_capi_set_write_precision(void *ev, int digits) { EV(ev)->writePrecision = digits; }

float *cupss_capi_field_real(void *ev, const char *name)
{
    evolver *e = EV(ev);
    if (e->existsField(name) < 0) return nullptr;
    return reinterpret_cast<float *>(e->fieldsMap[name]->real_array);
}

float *cupss_capi_field_comp(void *ev, const char *name)
{
    evolver *e = EV(ev);
    if (e->existsField(name) < 0) return nullptr;
    return reinterpret_cast<float *>(e->fieldsMap[name]->comp_array);
}

void cupss_capi_initialize_uniform(void *ev, const char *name, float value) { EV(ev)->initializeUniform(name, value); }
void cupss_capi_initialize_droplet(void *ev, const char *name, float v_out, float v_in, float radius, float width, int cx, int cy, int cz)
{
    EV(ev)->initializeDroplet(name, v_out, v_in, radius, width, cx, cy, cz);
}
void cupss_capi_add_droplet(void *ev, const char *name, float value, float radius, float width, int cx, int cy, int cz)
{
    EV(ev)->addDroplet(name, value, radius, width, cx, cy, cz);
}
void cupss_capi_initialize_half_system(void *ev, const char *name, float v1, float v2, float width, int direction)
{
    EV(ev)->initializeHalfSystem(name, v1, v2, width, direction);
}
void cupss_capi_initialize_from_file(void *ev, const char *name, const char *path, int skiprows, char delimiter)
{
    EV(ev)->initializeFromFile(name, path, skiprows, delimiter);
}

static void put_pres(std::ostringstream &os, const pres &p)
{
    char num[64];
    std::snprintf(num, sizeof num, "%.9g", (double)p.preFactor);
    os << "{" << num << "," << p.q2n << "," << p.iqx << "," << p.iqy << "," << p.iqz << "," << p.invq << "}";
}

/* Built-in HOST callbacks for tests: mirror boundary conditions on a doubled 2-D domain (the construction of
 * examples/07 and 08 of the reference, written from its description): the field on the quadrant i <= sx/2, j <= sy/2 is
 * reflected into the other three with sign `parity` per reflection.  They are installed through the public members
 * field::hasCB / field::callback exactly like a user's main() would (examples/08_neumann_dirichlet_bc). */
static void mirror_apply(float2 *a, int sx, int sy, int sz, float parity)
{
    /* every z plane on its own: the x/y mirror is local to a plane, so the same function serves 2-D grids (sz = 1), 3-D
     * grids and the z-slab a rank of a partitioned run is handed (sz = local planes) */
    for (int k = 0; k < sz; ++k) {
        float2 *p = a + (size_t)k * sx * sy;
        for (int j = 0; j < sy; ++j)
            for (int i = 0; i < sx; ++i) {
                const bool ri = i > sx / 2, rj = j > sy / 2;
                if (!ri && !rj) continue;
                const int si = ri ? sx - i : i, sj = rj ? sy - j : j;
                float sign = 1.0f;
                if (ri) sign *= parity;
                if (rj) sign *= parity;
                p[j * sx + i].x = sign * p[sj * sx + si].x;
            }
    }
}
static void mirror_even_cb(evolver *, float2 *a, int sx, int sy, int sz) { mirror_apply(a, sx, sy, sz, 1.0f); }
static void mirror_odd_cb(evolver *, float2 *a, int sx, int sy, int sz) { mirror_apply(a, sx, sy, sz, -1.0f); }
/* kind 2: a NON-symmetric, nonlinear callback -- the strip i < sx/8 is clamped to [-0.05, 0.05] and the strip
 * sx/2 <= i < sx/2 + sx/16 is scaled by 0.9.  What it leaves is not band-limited in x, which is what tells a product that
 * reads the dealiased copy unchanged (the reference) from one that low-passes it again. */
static void clamp_strip_cb(evolver *, float2 *a, int sx, int sy, int sz)
{
    for (size_t r = 0; r < (size_t)sy * sz; ++r) {
        float2 *row = a + r * sx;
        for (int i = 0; i < sx / 8; ++i) row[i].x = row[i].x > 0.05f ? 0.05f : (row[i].x < -0.05f ? -0.05f : row[i].x);
        for (int i = sx / 2; i < sx / 2 + sx / 16; ++i) row[i].x *= 0.9f;
    }
}

int cupss_capi_set_mirror_callback(void *ev, const char *name, int odd)
{
    if (EV(ev)->fieldsMap.find(name) == EV(ev)->fieldsMap.end()) return 1;
    field *f = EV(ev)->fieldsMap[name];
    f->hasCB = true;
    f->callback = odd == 2 ? clamp_strip_cb : (odd ? mirror_odd_cb : mirror_even_cb);
    return 0;
}

/* Built-in FOURIER-space callbacks for tests (field::hasCBFourier / callbackFourier, installed like a user's main() would).
 * They see comp_array as the reference lays it out: the full float2[sz][sy][sx] spectrum.
 *   kind 0  smooth low-pass g(n) = 1 / (1 + 0.0005 |n|^2) on every mode (keeps the Hermitian symmetry)
 *   kind 1  kind 0, then the modes 1 <= n_x <= 3 with n_y >= 0 are multiplied by 0.98 + 0.05 i on the POSITIVE-n_x side
 *           only: the spectrum is no longer Hermitian, and what survives is decided by the real-part projection that
 *           follows in field::setRHS
 *   kind 2  the columns sx/8 <= i < sx/4 are zeroed (a contiguous band: the device flavour does it with cudaMemset2D) */
static inline int signed_mode(int i, int n) { return i <= n / 2 ? i : i - n; }
static void fourier_apply(float2 *a, int sx, int sy, int sz, int kind)
{
    for (int k = 0; k < sz; ++k)
        for (int j = 0; j < sy; ++j)
            for (int i = 0; i < sx; ++i) {
                float2 &v = a[((size_t)k * sy + j) * sx + i];
                const int ni = signed_mode(i, sx), nj = signed_mode(j, sy), nk = signed_mode(k, sz);
                if (kind == 2) {
                    if (i >= sx / 8 && i < sx / 4) { v.x = 0.0f; v.y = 0.0f; }
                    continue;
                }
                const float g = 1.0f / (1.0f + 0.0005f * (float)(ni * ni + nj * nj + nk * nk));
                v.x *= g; v.y *= g;
                if (kind == 1 && ni >= 1 && ni <= 3 && nj >= 0) {
                    const float re = 0.98f * v.x - 0.05f * v.y, im = 0.98f * v.y + 0.05f * v.x;
                    v.x = re; v.y = im;
                }
            }
}
static void fourier_cb0(evolver *, float2 *a, int sx, int sy, int sz) { fourier_apply(a, sx, sy, sz, 0); }
static void fourier_cb1(evolver *, float2 *a, int sx, int sy, int sz) { fourier_apply(a, sx, sy, sz, 1); }
static void fourier_cb2(evolver *, float2 *a, int sx, int sy, int sz) { fourier_apply(a, sx, sy, sz, 2); }
#ifdef CUPSS_B200_PRODUCT
static void fourier_cb2_device(evolver *, float2 *a, int sx, int sy, int sz)   /* a is a DEVICE pointer (RUN_GPU flavour) */
{
    cudaMemset2D(a + sx / 8, (size_t)sx * sizeof(float2), 0, (size_t)(sx / 4 - sx / 8) * sizeof(float2), (size_t)sy * sz);
}
#endif

int cupss_capi_set_fourier_callback(void *ev, const char *name, int kind, int device_flavour)
{
    if (EV(ev)->fieldsMap.find(name) == EV(ev)->fieldsMap.end()) return 1;
    field *f = EV(ev)->fieldsMap[name];
    if (device_flavour) {
#ifdef CUPSS_B200_PRODUCT
        if (kind != 2) return 2;
        f->hasCBFourier = true;
        f->callbackFourier = fourier_cb2_device;
        return 0;
#else
        return 2;
#endif
    }
    if (kind < 0 || kind > 2) return 2;
    f->hasCBFourier = true;
    f->callbackFourier = kind == 0 ? fourier_cb0 : (kind == 1 ? fourier_cb1 : fourier_cb2);
    return 0;
}

int cupss_capi_dump_plan(void *ev, char *buf, int buflen)
{
    evolver *e = EV(ev);
    std::ostringstream os;
    for (size_t f = 0; f < e->fields.size(); f++) {
        field *F = e->fields[f];
        os << "field " << F->name << " dynamic=" << (F->dynamic ? 1 : 0)
           << " alias=" << (F->needsaliasing ? 1 : 0) << " order=" << F->aliasing_order
           << " noisy=" << (F->isNoisy ? 1 : 0);
        if (F->isNoisy) { os << " noise="; put_pres(os, F->noise_amplitude); }
        os << "\n  implicit";
        for (size_t i = 0; i < F->implicit.size(); i++) { os << " "; put_pres(os, F->implicit[i]); }
        os << "\n";
        for (size_t t = 0; t < F->terms.size(); t++) {
            term *T = F->terms[t];
            os << "  term";
            for (size_t i = 0; i < T->prefactors_h.size(); i++) { os << " "; put_pres(os, T->prefactors_h[i]); }
            os << " (";
            for (size_t k = 0; k < T->product.size(); k++) os << " " << T->product[k]->name;
            os << " )\n";
        }
    }
    std::string s = os.str();
    if ((int)s.size() + 1 > buflen) return -(int)(s.size() + 1);
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

void cupss_capi_print_information(void *ev) { EV(ev)->printInformation(); }
void cupss_capi_copy_host_to_device(void *ev, const char *name)
{
    if (EV(ev)->fieldsMap.find(name) != EV(ev)->fieldsMap.end()) EV(ev)->fieldsMap[name]->copyHostToDevice();
}

#ifdef CUPSS_B200_PRODUCT
void *cupss_capi_engine_plan(void *ev) { return EV(ev)->enginePlan(); }
void cupss_capi_set_noise_seed(void *ev, unsigned long long seed) { EV(ev)->setNoiseSeed(seed); }
unsigned long long cupss_capi_get_noise_seed(void *ev) { return EV(ev)->getNoiseSeed(); }
void cupss_capi_set_partition(void *ev, int rank, int nranks, const void *id) { EV(ev)->setPartition(rank, nranks, id); }
#endif

} /* extern "C" */
