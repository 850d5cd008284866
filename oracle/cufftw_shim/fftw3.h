/* ORACLE / TEST INFRASTRUCTURE ONLY.
 * FFTW 3 is not installed in this image; the reference's headers include <fftw3.h> unconditionally
 * (/root/reference/inc/cupss/field.h:6, term.h:5).  For the build of the reference's OWN GPU (cuFFT) path the
 * FFTW-3 API is taken from CUDA's cufftw.h; the RUN_GPU path never executes those plans. */
#include <cufftw.h>
