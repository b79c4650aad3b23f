/* TEST INFRASTRUCTURE -- not product code (see dabmod_oracle.c).
 *
 * CPU restatement of the FIXED-POINT engine of the path (FFTEngine::KISS,
 * src/DabModulator.cpp:144-224): complexfix = std::complex<fpm::fixed<int16, int32, 14>> carriers
 * (src/Buffer.h:42-43), the vendored KISS FFT built with FIXED_POINT=16 (kiss/kiss_fft.c,
 * kiss/_kiss_fft_guts.h), no GainControl, GuardIntervalInserter do_process<complexfix>.
 * Everything here is integer arithmetic: parity with the reference is BIT-EXACT and is pinned by
 * tests/test_fixed.py against the compiled reference (oracle/_ref) and tests/golden/fixed_*.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DABO_EXPORT __attribute__((visibility("default")))

typedef struct { int16_t r, i; } cfix;     /* kiss_fft_cpx with FIXED_POINT=16 == complexfix in memory */

/* ---- fpm::fixed<int16_t, int32_t, 14> (fpm/fixed.hpp) ------------------------------------ */
/* construction from a floating value with rounding (fixed.hpp:47-53) */
DABO_EXPORT int16_t dabo_fix_from_double(double v)
{
    return (int16_t)(v >= 0.0 ? v * 16384.0 + 0.5 : v * 16384.0 - 0.5);
}

/* operator*= with EnableRounding (fixed.hpp:156-169): one extra bit, then round half away from zero.
 * C division and remainder truncate toward zero, as in the reference. */
static int16_t fix_mul(int16_t a, int16_t b)
{
    const int32_t v = ((int32_t)a * b) / (16384 / 2);
    return (int16_t)(v / 2 + v % 2);
}

/* ---- carriers: the float chain's values are exactly representable --------------------------- */
/* QpskSymbolMapper.cpp:46-101 maps to +-fixed(M_SQRT1_2) = +-11585, PhaseReference.cpp:139-150 to
 * {+-16384, 0}; DifferentialModulator.cpp:45-76 multiplies them with fix_mul: 11585 * 11585 -> 8192,
 * 16384 * 11585 -> 11585, 16384 * 16384 -> 16384, so the product chain only ever visits
 * {0, +-11585, +-16384} -- the fixed images of the values {0, +-1/sqrt2, +-1} of the float chain
 * (dabmod_oracle.c stages 1-4, bit-exact against the reference).  The conversion below is therefore
 * exact; tests compare it with the reference's own fixed-point "mux" stage bit for bit. */
DABO_EXPORT void dabo_fix_carriers(const float *z, long n_floats, int16_t *out)
{
    for (long i = 0; i < n_floats; i++) out[i] = dabo_fix_from_double((double)z[i]);
}

/* ---- KISS FFT, FIXED_POINT=16, inverse (kiss/kiss_fft.c, kiss/_kiss_fft_guts.h) -------------- */
#define SROUND(x) ((int16_t)(((x) + (1 << 14)) >> 15))                     /* sround, FRACBITS = 15 */
static int16_t divscalar(int16_t x, int k) { return SROUND((int32_t)x * (32767 / k)); }   /* DIVSCALAR */
static cfix c_mul(cfix a, cfix b)                                            /* C_MUL */
{
    cfix m;
    m.r = SROUND((int32_t)a.r * b.r - (int32_t)a.i * b.i);
    m.i = SROUND((int32_t)a.r * b.i + (int32_t)a.i * b.r);
    return m;
}
static cfix c_add(cfix a, cfix b) { cfix c = {(int16_t)(a.r + b.r), (int16_t)(a.i + b.i)}; return c; }
static cfix c_sub(cfix a, cfix b) { cfix c = {(int16_t)(a.r - b.r), (int16_t)(a.i - b.i)}; return c; }

typedef struct {
    int n, nfac;
    int p[32], m[32];          /* kf_factor: 4s first, then 2, 3, 5, ... (kiss_fft.c:293-315) */
    cfix *tw;                  /* kiss_fft_alloc: twiddles[i] = kf_cexp(+2 pi i / n) for the inverse transform */
} kiss_fixed;

static kiss_fixed *kiss_fixed_new(int n)
{
    kiss_fixed *st = calloc(1, sizeof(*st));
    st->n = n;
    st->tw = malloc(sizeof(cfix) * n);
    for (int i = 0; i < n; i++) {
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        const double phase = 2 * pi * i / n;                 /* -2 pi i / n, negated for inverse_fft */
        st->tw[i].r = (int16_t)floor(.5 + 32767 * cos(phase));      /* kf_cexp, FIXED_POINT */
        st->tw[i].i = (int16_t)floor(.5 + 32767 * sin(phase));
    }
    int p = 4, rem = n;
    const double floor_sqrt = floor(sqrt((double)n));
    do {
        while (rem % p) {
            switch (p) {
                case 4: p = 2; break;
                case 2: p = 3; break;
                default: p += 2; break;
            }
            if (p > floor_sqrt) p = rem;
        }
        rem /= p;
        st->p[st->nfac] = p;
        st->m[st->nfac] = rem;
        st->nfac++;
    } while (rem > 1);
    return st;
}

/* kf_work (kiss_fft.c:236-291), radix 2 and 4 only (N = 256 ... 2048) */
static void kf_work(const kiss_fixed *st, cfix *Fout, const cfix *f, int fstride, int level)
{
    const int p = st->p[level], m = st->m[level];
    cfix *const Fout_beg = Fout;
    cfix *const Fout_end = Fout + p * m;
    if (m == 1) {
        do { *Fout = *f; f += fstride; } while (++Fout != Fout_end);
    }
    else {
        do {
            kf_work(st, Fout, f, fstride * p, level + 1);
            f += fstride;
        } while ((Fout += m) != Fout_end);
    }
    Fout = Fout_beg;
    if (p == 2) {                                            /* kf_bfly2 */
        cfix *F2 = Fout + m;
        const cfix *tw1 = st->tw;
        for (int k = 0; k < m; k++) {
            Fout->r = divscalar(Fout->r, 2); Fout->i = divscalar(Fout->i, 2);
            F2->r = divscalar(F2->r, 2); F2->i = divscalar(F2->i, 2);
            const cfix t = c_mul(*F2, *tw1);
            tw1 += fstride;
            *F2 = c_sub(*Fout, t);
            *Fout = c_add(*Fout, t);
            ++F2; ++Fout;
        }
    }
    else if (p == 4) {                                       /* kf_bfly4, st->inverse */
        const cfix *tw1 = st->tw, *tw2 = st->tw, *tw3 = st->tw;
        const int m2 = 2 * m, m3 = 3 * m;
        for (int k = 0; k < m; k++) {
            cfix s[6];
            for (int q = 0; q < 4; q++) {
                Fout[q * m].r = divscalar(Fout[q * m].r, 4);
                Fout[q * m].i = divscalar(Fout[q * m].i, 4);
            }
            s[0] = c_mul(Fout[m], *tw1);
            s[1] = c_mul(Fout[m2], *tw2);
            s[2] = c_mul(Fout[m3], *tw3);
            s[5] = c_sub(*Fout, s[1]);
            *Fout = c_add(*Fout, s[1]);
            s[3] = c_add(s[0], s[2]);
            s[4] = c_sub(s[0], s[2]);
            Fout[m2] = c_sub(*Fout, s[3]);
            tw1 += fstride; tw2 += fstride * 2; tw3 += fstride * 3;
            *Fout = c_add(*Fout, s[3]);
            Fout[m].r = (int16_t)(s[5].r - s[4].i);
            Fout[m].i = (int16_t)(s[5].i + s[4].r);
            Fout[m3].r = (int16_t)(s[5].r + s[4].i);
            Fout[m3].i = (int16_t)(s[5].i - s[4].r);
            ++Fout;
        }
    }
    else abort();
}

/* one inverse transform of n points (kiss_fft with fin != fout) */
DABO_EXPORT void dabo_kiss_fixed_ifft(int n, const int16_t *in, int16_t *out)
{
    kiss_fixed *st = kiss_fixed_new(n);
    kf_work(st, (cfix *)out, (const cfix *)in, 1, 0);
    free(st->tw);
    free(st);
}

/* ---- OfdmGeneratorFixed::process (src/OfdmGenerator.cpp:529-579) ---------------------------- */
/* in: nsym x K carriers, out: nsym x N samples (int16 pairs) */
DABO_EXPORT void dabo_ofdm_fixed(int N, int K, const int16_t *in, int nsym, int16_t *out)
{
    kiss_fixed *st = kiss_fixed_new(N);
    cfix *X = calloc(N, sizeof(cfix));
    const int pos_dst = (K & 1) ? 0 : 1, pos_size = (K + 1) / 2;
    const int neg_dst = N - K / 2, neg_src = (K + 1) / 2, neg_size = K / 2;
    for (int s = 0; s < nsym; s++) {
        const cfix *c = (const cfix *)in + (size_t)s * K;
        memset(X, 0, sizeof(cfix) * N);
        memcpy(X + pos_dst, c, sizeof(cfix) * pos_size);
        memcpy(X + neg_dst, c + neg_src, sizeof(cfix) * neg_size);
        kf_work(st, (cfix *)out + (size_t)s * N, X, 1, 0);
    }
    free(X);
    free(st->tw);
    free(st);
}

/* ---- GuardIntervalInserter do_process<complexfix> (src/GuardIntervalInserter.cpp:96-323) ---- */
/* in: (L+1) x N, out: null_size + L * sym_size samples.  W = windowOverlap.  The window is
 * fixed(0.5 (1 - cos(pi i / (2W - 1)))) (:103-112), applied with the rounding multiply to re and im
 * (std::complex<fixed> * fixed), and overlapping edges ADD (int16, :222-234). */
DABO_EXPORT void dabo_guard_fixed(int N, int L, int null_size, int sym_size, const int16_t *in_, int W, int16_t *out_)
{
    const cfix *in = (const cfix *)in_;
    cfix *out = (cfix *)out_;
    const long total = null_size + (long)L * sym_size;
    if (W == 0) {
        int size = null_size;
        for (int l = 0; l <= L; l++) {
            const int pre = size - N;
            memcpy(out, in + N - pre, sizeof(cfix) * pre);
            memcpy(out + pre, in, sizeof(cfix) * N);
            in += N; out += size; size = sym_size;
        }
        return;
    }
    int16_t *w = malloc(sizeof(int16_t) * 2 * W);
    for (int i = 0; i < 2 * W; i++) {
        const float value = (float)(0.5 * (1.0 - cos(M_PI * i / (2 * W - 1))));
        w[i] = dabo_fix_from_double((double)value);
    }
    cfix *acc = calloc(total + W, sizeof(cfix));
    long pos = 0;
    for (int l = 0; l <= L; l++) {
        const int size = l == 0 ? null_size : sym_size;
        const int pre = size - N;
        const cfix *x = in + (size_t)l * N;
        const int first = l == 0, last = l == L;
        for (long o = first ? 0 : -W; o < size + (last ? 0 : W); o++) {
            long ix = (o - pre) % N;
            if (ix < 0) ix += N;
            cfix v = x[ix];
            if (!first && o < W) {                          /* rising edge, added to what is there */
                v.r = fix_mul(v.r, w[o + W]); v.i = fix_mul(v.i, w[o + W]);
                acc[pos + o].r = (int16_t)(acc[pos + o].r + v.r);
                acc[pos + o].i = (int16_t)(acc[pos + o].i + v.i);
                continue;
            }
            if (!last && o >= size - W) {                   /* falling edge */
                const int16_t g = w[2 * W - 1 - (o - (size - W))];
                v.r = fix_mul(v.r, g); v.i = fix_mul(v.i, g);
            }
            acc[pos + o] = v;
        }
        pos += size;
    }
    memcpy(out, acc, sizeof(cfix) * total);
    free(acc);
    free(w);
}
