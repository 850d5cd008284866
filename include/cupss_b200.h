/*
 * cupss_b200.h -- C ABI of the B200-native (sm_100a) advanceTime engine.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The C++ host
 * layer in inc/cupss/ (same classes as the reference: evolver / field / term / parser) forwards
 * its RUN_GPU path to these entry points; INTEGRATION.md shows the equivalent patch against the
 * reference's own sources.  Every function returns 0 on success and a non-zero code on failure
 * (cupss_b200_last_error() gives the message); the C++ layer turns failures into the reference's
 * behaviour (message + std::exit(1), /root/reference/src/cu_utils.cpp:4-10).
 *
 * Reference interface replaced by each entry point (paths under /root/reference):
 *   cupss_b200_create            evolver ctor + field/term ctors' cudaMalloc + cufftPlan*    src/evolver.cpp:47-74, src/field_init.cpp:15-137, src/term_init.cpp:11-75
 *   cupss_b200_add_field         evolver::createField                                        src/evolver.cpp:177-196
 *   cupss_b200_set_implicit      field::implicit + field::precalculateImplicit               src/parser.cpp:728-739, src/field_init.cpp:237-284
 *   cupss_b200_add_term          evolver::createTerm + term::prepareDevice/precomputePrefactors  src/evolver.cpp:326-362, src/term_init.cpp:108-200
 *   cupss_b200_set_noise         evolver::addNoise + curand generator set-up                 src/evolver.cpp:166-175, src/field_init.cpp:130-137
 *   cupss_b200_upload_real       field::copyHostToDevice + field::toComp                     src/field.cpp:332-335, 261-273 (called from evolver::prepareProblem, src/evolver.cpp:114-117)
 *   cupss_b200_finalize          field::prepareDevice / term::prepareDevice                  src/evolver.cpp:121-125, src/field_init.cpp:215-235
 *   cupss_b200_step              evolver::advanceTime minus the output trigger               src/evolver.cpp:205-224
 *                                (= field::updateTerms/setRHS, term::update and the 12 extern "C" *_gpu
 *                                   launchers of inc/cupss/field_kernels.cuh:7-19, inc/cupss/term_kernels.cuh:7-15,
 *                                   plus every cufftExecC2C / curandGenerateNormal on that path)
 *   cupss_b200_step_stage +      field::setRHS callback hook: callback(system, real_array_d, ...) and, for a field that
 *   cupss_b200_real_view_*         products read, callback(system, real_dealiased_d, ...)             src/field.cpp:68-86
 *   cupss_b200_comp_view_*       field::setRHS Fourier hook: callbackFourier(system, comp_array_d, ...) between the update
 *                                and the dealias / toReal that follow it                              src/field.cpp:48-57
 *   cupss_b200_download_real     field::copyRealDeviceToHost (real_array)                    src/field.cpp:345-348
 *   cupss_b200_download_comp     field::copyDeviceToHost (comp_array)                        src/field.cpp:337-340
 *   cupss_b200_destroy           evolver/field/term dtors                                    src/evolver.cpp:40-45, src/field_init.cpp:156-189
 */
#ifndef CUPSS_B200_H
#define CUPSS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cupss_b200_plan cupss_b200_plan;

/* struct pres of the reference, field for field (inc/cupss/defines.h:31-39). */
typedef struct cupss_b200_pres {
    float preFactor;
    int q2n, iqx, iqy, iqz, invq;
} cupss_b200_pres;

enum {
    CUPSS_B200_OK = 0,
    CUPSS_B200_ERR_ARG = 1,      /* bad argument / unsupported configuration */
    CUPSS_B200_ERR_CUDA = 2,     /* a CUDA runtime call failed */
    CUPSS_B200_ERR_STATE = 3,    /* call out of order (e.g. step before finalize) */
    CUPSS_B200_ERR_COMM = 4      /* NCCL failure */
};

/* Dealias mask convention. */
enum {
    CUPSS_B200_DEALIAS_GPU_RULE = 0,  /* dealias_k, src/field_kernels.cu:229-256: |n_a| <= s_a/(order+1) on every axis */
    CUPSS_B200_DEALIAS_CPU_RULE = 1   /* field::dealias CPU loop with its `nj` typo, src/field.cpp:220 */
};

int cupss_b200_create(cupss_b200_plan **out, int sx, int sy, int sz, float dx, float dy, float dz, float dt);
void cupss_b200_destroy(cupss_b200_plan *p);

/* Multi-GPU slab partition (3-D only; one process per GPU).  Must precede finalize/upload.
 * Real space is split on z, Fourier space on ky; sz and sy must be divisible by nranks.
 * `nccl_id` is the 128-byte ncclUniqueId produced by cupss_b200_nccl_unique_id on rank 0. */
int cupss_b200_nccl_unique_id(void *id128);
int cupss_b200_set_partition(cupss_b200_plan *p, int rank, int nranks, const void *nccl_id128);

int cupss_b200_add_field(cupss_b200_plan *p, const char *name, int dynamic);   /* returns field id >= 0, or -code */
int cupss_b200_set_implicit(cupss_b200_plan *p, int field, const cupss_b200_pres *pres, int n);
int cupss_b200_clear_terms(cupss_b200_plan *p, int field);
int cupss_b200_add_term(cupss_b200_plan *p, int field, const cupss_b200_pres *pres, int n, const int *product, int m);
int cupss_b200_set_noise(cupss_b200_plan *p, int field, const cupss_b200_pres *amplitude /* NULL: off */, unsigned long long seed);
int cupss_b200_set_dealias_rule(cupss_b200_plan *p, int rule);

/* Builds (or rebuilds, after parameter changes) the fused per-equation plan.  Re-callable; keeps field data. */
int cupss_b200_finalize(cupss_b200_plan *p);

/* Host <-> device, reference layout: float2[sz_local][sy][sx], value in .x (inc/cupss/field.h:67-70). */
int cupss_b200_upload_real(cupss_b200_plan *p, int field, const float *host_float2);
int cupss_b200_download_real(cupss_b200_plan *p, int field, float *host_float2);
int cupss_b200_download_comp(cupss_b200_plan *p, int field, float *host_float2);   /* full spectrum; nranks == 1 */

int cupss_b200_step(cupss_b200_plan *p, int nsteps);   /* asynchronous on the plan's stream */

/* User callbacks (boundary conditions).  A step is then driven as: step_stage(0) [constraint fields], callbacks of the
 * constraint fields, step_stage(1) [dynamic fields + step counter], callbacks of the dynamic fields.  A callback sees
 * float2[sz][sy][sx] on the device, value in .x (the reference's real_array_d): which = 0 the field, 1 its dealiased
 * copy.  begin materialises the view and synchronises; commit waits for the device and transforms the view back. */
int cupss_b200_step_stage(cupss_b200_plan *p, int stage);
int cupss_b200_real_view_begin(cupss_b200_plan *p, int field, int which, void **dev_float2);
int cupss_b200_real_view_commit(cupss_b200_plan *p, int field, int which);
/* Fourier-space callbacks: after step_stage(s), for every flagged field of that sweep (before its real-space callback).
 * The view is the reference's comp_array_d: the FULL float2[sz][sy][sx] spectrum.  commit keeps the Hermitian part of
 * what the callback wrote (what the reference's toReal -> normalize -> toComp keeps) and recomputes the dealiased copy. */
int cupss_b200_comp_view_begin(cupss_b200_plan *p, int field, void **dev_float2);
int cupss_b200_comp_view_commit(cupss_b200_plan *p, int field);
int cupss_b200_sync(cupss_b200_plan *p);

/* Field flags computed by finalize (field::needsaliasing / aliasing_order, src/term_init.cpp:123-126). */
int cupss_b200_field_alias(cupss_b200_plan *p, int field, int *needsaliasing, int *order);

/* Measurement hooks (bench.py): CUDA-event timing on the stream the kernels are launched on. */
int cupss_b200_time_steps(cupss_b200_plan *p, int nsteps, float *elapsed_ms);
/* Per-launch breakdown of ONE step (events between launches, graph bypassed).
 * names: buffer of n_max * 64 chars; ms / bytes: n_max floats / doubles (algorithmic bytes per launch).
 * Returns the number of launches in *n. */
int cupss_b200_profile_step(cupss_b200_plan *p, int n_max, char *names, float *ms, double *bytes, int *n);
int cupss_b200_launches_per_step(cupss_b200_plan *p);
double cupss_b200_bytes_per_step(cupss_b200_plan *p);   /* algorithmic HBM bytes of one step, this rank */
double cupss_b200_comm_bytes_per_step(cupss_b200_plan *p); /* bytes this rank sends per step */
void *cupss_b200_device_spectrum(cupss_b200_plan *p, int field);   /* borrowed device pointer (callbacks) */

/* Compiles (without loading) the run-time-specialised k stage of a synthetic sweep with NVRTC; needs no GPU. */
int cupss_b200_jit_selftest(char *log, int loglen);

const char *cupss_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
