/* TEST INFRASTRUCTURE -- not product code.
 *
 * fftw3.h shim implemented over the reference's vendored KISS FFT compiled as
 * float (no -DFIXED_POINT).  The Makefile renames KISS's public symbols with
 * -Dkiss_fft=kf32_fft ... so this float build can coexist with the int16 KISS
 * build that OfdmGeneratorFixed links (Makefile.am:38).
 * Both FFTW and float KISS compute the unnormalised DFT; sign=+1 is the
 * inverse (e^{+j2pi kn/N}).
 */
#include "fftw3.h"
#include <stdlib.h>
#include "kiss_fft.h"

struct fftwf_plan_s {
    kiss_fft_cfg cfg;
    fftwf_complex *in;
    fftwf_complex *out;
};

void *fftwf_malloc(size_t n)
{
    void *p = NULL;
    if (posix_memalign(&p, 32, n ? n : 32) != 0) return NULL;
    return p;
}

void fftwf_free(void *p) { free(p); }

void fftwf_set_timelimit(double seconds) { (void)seconds; }

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out,
                             int sign, unsigned flags)
{
    (void)flags;
    fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
    if (!p) return NULL;
    p->cfg = kiss_fft_alloc(n, sign == FFTW_BACKWARD, NULL, NULL);
    p->in = in;
    p->out = out;
    return p;
}

void fftwf_execute(const fftwf_plan p)
{
    kiss_fft(p->cfg, (const kiss_fft_cpx *)p->in, (kiss_fft_cpx *)p->out);
}

void fftwf_destroy_plan(fftwf_plan p)
{
    if (!p) return;
    free(p->cfg);
    free(p);
}
