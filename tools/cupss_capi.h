/*
 * Flat C facade over the public C++ `evolver` API (inc/cupss/evolver.h, which
 * mirrors /root/reference/inc/cupss/evolver.h:14-86).
 *
 * The SAME source (tools/cupss_capi.cpp) is compiled twice:
 *   - against this repo's inc/ + host library  -> lib/libcupss.so        (product, GPU)
 *   - against /root/reference/inc + sources    -> oracle/_ref/libcupss_ref_{u,f}.so (oracle, CPU)
 * so tests and bench.py drive both implementations through identical calls
 * (ctypes).  It touches nothing but public members a user's main() could touch.
 */
#ifndef CUPSS_CAPI_H
#define CUPSS_CAPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cupss_capi_pres { float preFactor; int q2n, iqx, iqy, iqz, invq; } cupss_capi_pres;

void *cupss_capi_create(int with_cuda, int sx, int sy, int sz, float dx, float dy, float dz, float dt, int write_every);
void cupss_capi_destroy(void *ev);
int cupss_capi_create_field(void *ev, const char *name, int dynamic);
int cupss_capi_add_parameter(void *ev, const char *name, float value);
int cupss_capi_add_equation(void *ev, const char *equation);
int cupss_capi_add_noise(void *ev, const char *field, const char *expr);
int cupss_capi_create_term(void *ev, const char *field, const cupss_capi_pres *pres, int npres, const char *const *product, int nproduct);
int cupss_capi_create_from_file(void *ev, const char *path);
void cupss_capi_prepare_problem(void *ev);
int cupss_capi_advance_time(void *ev, int nsteps);
void cupss_capi_copy_all_data_to_host(void *ev);
void cupss_capi_write_out(void *ev);
void cupss_capi_set_output_field(void *ev, const char *name, int on);
int cupss_capi_update_parameter(void *ev, const char *name, float value);
float cupss_capi_get_parameter(void *ev, const char *name);
int cupss_capi_get_timestep(void *ev);
float cupss_capi_get_time(void *ev);
void cupss_capi_set_write_precision(void *ev, int digits);
/* host mirrors: interleaved (re,im) float pairs, sx*sy*sz of them; valid for the evolver's lifetime */
float *cupss_capi_field_real(void *ev, const char *name);
float *cupss_capi_field_comp(void *ev, const char *name);
void cupss_capi_initialize_uniform(void *ev, const char *name, float value);
void cupss_capi_initialize_droplet(void *ev, const char *name, float v_out, float v_in, float radius, float width, int cx, int cy, int cz);
void cupss_capi_add_droplet(void *ev, const char *name, float value, float radius, float width, int cx, int cy, int cz);
void cupss_capi_initialize_half_system(void *ev, const char *name, float v1, float v2, float width, int direction);
void cupss_capi_initialize_from_file(void *ev, const char *name, const char *path, int skiprows, char delimiter);
/* installs a built-in host callback on a field (for RUN_CPU evolvers): odd = 0 / 1 mirror boundary condition, even / odd;
 * odd = 2 a non-symmetric clamp of two boundary strips */
int cupss_capi_set_mirror_callback(void *ev, const char *name, int odd);
/* installs a built-in Fourier-space callback (kinds: see cupss_capi.cpp); device_flavour = 1 needs the product and RUN_GPU */
int cupss_capi_set_fourier_callback(void *ev, const char *name, int kind, int device_flavour);
/* textual dump of the parsed system from public members (fields, implicit pres, terms, products, noise, aliasing) */
int cupss_capi_dump_plan(void *ev, char *buf, int buflen);
void cupss_capi_print_information(void *ev);                  /* evolver::printInformation (stdout) */
void cupss_capi_copy_host_to_device(void *ev, const char *name);   /* field::copyHostToDevice */


/* ---- product build only (-DCUPSS_B200_PRODUCT): access to the engine behind the evolver ---- */
void *cupss_capi_engine_plan(void *ev);                       /* cupss_b200_plan* (include/cupss_b200.h) */
void cupss_capi_set_noise_seed(void *ev, unsigned long long seed);
unsigned long long cupss_capi_get_noise_seed(void *ev);
void cupss_capi_set_partition(void *ev, int rank, int nranks, const void *nccl_id128);

#ifdef __cplusplus
}
#endif
#endif
