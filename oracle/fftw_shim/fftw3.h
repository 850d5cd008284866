/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product path.
 *
 * Minimal stand-in for the FFTW 3 single-precision API, exactly the subset the
 * reference's CPU path calls:
 *   plans     /root/reference/src/field_init.cpp:50-118, src/term_init.cpp:25-58
 *   executes  /root/reference/src/field.cpp:255,257,269,271,328, src/term.cpp:61
 * FFTW itself (libfftw3f, no version pinned by the reference: README.md:30-31,
 * CMakeLists.txt:32) is not installed in this image and there is no network, so
 * the oracle build of the unmodified reference sources links this shim instead.
 * Semantics restated from FFTW's published definition: unnormalised DFT,
 *   Y[k] = sum_j X[j] exp(sign * 2*pi*i * j*k / n),  FFTW_FORWARD = -1,
 * row-major multi-dimensional layout with the LAST dimension fastest.
 */
#ifndef CUPSS_ORACLE_FFTW3_SHIM_H
#define CUPSS_ORACLE_FFTW3_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct cupss_shim_plan *fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftwf_plan fftwf_plan_dft_1d(int n0, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
fftwf_plan fftwf_plan_dft_2d(int n0, int n1, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
fftwf_plan fftwf_plan_dft_3d(int n0, int n1, int n2, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);

/* shim-only knob: number of OpenMP threads fftwf_execute may use (default 1,
 * like the serial reference; the bench's CPU arm raises it and says so). */
void cupss_shim_set_threads(int n);
int cupss_shim_get_threads(void);
/* shim-only knob: 1 (default) = double arithmetic inside each transform (float in/out, one rounding);
 * 0 = float arithmetic throughout, like libfftw3f. */
void cupss_shim_set_double(int on);
int cupss_shim_get_double(void);

#ifdef __cplusplus
}
#endif
#endif
