/* TEST INFRASTRUCTURE -- not product code.
 *
 * Minimal single-precision FFTW3 API surface, so that the UNMODIFIED reference
 * sources under /root/reference can be compiled in a container that has no
 * libfftw3f.  Only the six calls the reference makes are declared
 * (OfdmGenerator.cpp:106-153,228,334,370; Resampler.cpp:94-127,151,183;
 * DabMod.cpp:435-446).  Backed by fftw3_kiss.c (float build of the
 * reference's own vendored kiss/kiss_fft.c).
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct fftwf_plan_s *fftwf_plan;

#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

void *fftwf_malloc(size_t n);
void  fftwf_free(void *p);
void  fftwf_set_timelimit(double seconds);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out,
                             int sign, unsigned flags);
void  fftwf_execute(const fftwf_plan p);
void  fftwf_destroy_plan(fftwf_plan p);

#ifdef __cplusplus
}
#endif
#endif
