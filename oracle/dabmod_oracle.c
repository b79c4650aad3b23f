/* TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the ODR-DabMod COFDM
 * hot path.  Not product code: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load it.  The CUDA product never calls it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path
 * (SURVEY.md section 4), so this restatement is pinned against the reference
 * itself: oracle/_ref/libdabmod_ref.so is the unmodified reference code
 * (built by oracle/Makefile with an fftw3 shim over the reference's vendored
 * float KISS FFT -- FFTW3 itself is not installed here), and
 * tests/test_oracle_vs_reference.py + tests/golden/ compare every stage.
 *
 * Written from the behaviour of the reference files cited at each function
 * (paths relative to /root/reference).  All stage boundaries are float32 like
 * the reference's Buffers; DFTs are evaluated in double and rounded once, so
 * the oracle sits between any two float32 FFT libraries (FFTW, KISS, ours).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared dabmod_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

typedef struct { float re, im; } cf;
typedef struct { double re, im; } cd;

#define DABO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* Mode table: src/DabModulator.cpp:84-122; FrequencyInterleaver.cpp:41-66;  */
/* BlockPartitioner.cpp:44-73                                               */
/* ------------------------------------------------------------------------ */
typedef struct {
    int32_t mode;       /* 1..4 */
    int32_t L;          /* nbSymbols incl. phase reference, excl. null */
    int32_t K;          /* carriers */
    int32_t N;          /* FFT size (spacing) */
    int32_t null_size;
    int32_t sym_size;
    int32_t beta;       /* frequency interleaver LCG constant */
    int32_t tf_bytes;   /* BlockPartitioner output bytes per TF = (L-1)*K/4 */
    int32_t tf_samples; /* null_size + L*sym_size */
} dabo_mode;

DABO_EXPORT int dabo_mode_params(int mode, dabo_mode *m)
{
    switch (mode) {
        case 1: *m = (dabo_mode){1, 76, 1536, 2048, 2656, 2552, 511, 0, 0}; break;
        case 2: *m = (dabo_mode){2, 76, 384, 512, 664, 638, 127, 0, 0}; break;
        case 3: *m = (dabo_mode){3, 153, 192, 256, 345, 319, 63, 0, 0}; break;
        case 4: *m = (dabo_mode){4, 76, 768, 1024, 1328, 1276, 255, 0, 0}; break;
        default: return -1;
    }
    m->tf_bytes = (m->L - 1) * m->K / 4;
    m->tf_samples = m->null_size + m->L * m->sym_size;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* a1  QpskSymbolMapper::process  (src/QpskSymbolMapper.cpp:105-156)         */
/* Per symbol K/4 bytes: first K/8 bytes = I bits, next K/8 = Q bits, MSB    */
/* first.  bit 0 -> +1/sqrt2, bit 1 -> -1/sqrt2.                             */
/* ------------------------------------------------------------------------ */
DABO_EXPORT void dabo_qpsk(const dabo_mode *m, const uint8_t *bits, int nsym, cf *out)
{
    const float v = (float)M_SQRT1_2;
    const int K = m->K;
    for (int s = 0; s < nsym; s++) {
        const uint8_t *bi = bits + (size_t)s * (K / 4);
        const uint8_t *bq = bi + K / 8;
        for (int n = 0; n < K; n++) {
            const int sh = 7 - (n & 7);
            const int i = (bi[n >> 3] >> sh) & 1;
            const int q = (bq[n >> 3] >> sh) & 1;
            out[(size_t)s * K + n].re = i ? -v : v;
            out[(size_t)s * K + n].im = q ? -v : v;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* a2  FrequencyInterleaver ctor + do_process                                */
/* (src/FrequencyInterleaver.cpp:31-93, 103-126)                             */
/* ------------------------------------------------------------------------ */
DABO_EXPORT void dabo_freq_index(const dabo_mode *m, int32_t *idx)
{
    const int N = m->N, K = m->K;
    int perm = 0, n = 0;
    for (int j = 1; j < N; j++) {
        perm = (13 * perm + m->beta) & (N - 1);
        if (perm >= (N - K) / 2 && perm <= N - (N - K) / 2 && perm != N / 2) {
            idx[n++] = perm > N / 2 ? perm - (1 + N / 2) : perm + (K - N / 2);
        }
    }
}

DABO_EXPORT void dabo_freq_interleave(const dabo_mode *m, const cf *in, int nsym, cf *out)
{
    const int K = m->K;
    int32_t *idx = malloc(sizeof(int32_t) * K);
    dabo_freq_index(m, idx);
    for (int s = 0; s < nsym; s++)
        for (int j = 0; j < K; j++)
            out[(size_t)s * K + idx[j]] = in[(size_t)s * K + j];
    free(idx);
}

/* ------------------------------------------------------------------------ */
/* a3  PhaseReference (src/PhaseReference.cpp:35-44, 91-124, 152-171)        */
/* ETSI EN 300 401 table 43 (h) and tables 44-47 (i, n per 32-carrier block) */
/* Returns the quarter-turn index 0..3 per carrier; value = j^index.         */
/* ------------------------------------------------------------------------ */
static const uint8_t H_TAB[4][32] = {
    {0,2,0,0,0,0,1,1,2,0,0,0,2,2,1,1,0,2,0,0,0,0,1,1,2,0,0,0,2,2,1,1},
    {0,3,2,3,0,1,3,0,2,1,2,3,2,3,3,0,0,3,2,3,0,1,3,0,2,1,2,3,2,3,3,0},
    {0,0,0,2,0,2,1,3,2,2,0,2,2,0,1,3,0,0,0,2,0,2,1,3,2,2,0,2,2,0,1,3},
    {0,1,2,1,0,3,3,2,2,3,2,1,2,1,3,2,0,1,2,1,0,3,3,2,2,3,2,1,2,1,3,2},
};
/* {i, n} per block of 32 carriers, in carrier-buffer order */
static const uint8_t IN_TM1[48][2] = {
    {0,3},{3,1},{2,1},{1,1},{0,2},{3,2},{2,1},{1,0},{0,2},{3,2},{2,3},{1,3},
    {0,0},{3,2},{2,1},{1,3},{0,3},{3,3},{2,3},{1,0},{0,3},{3,0},{2,1},{1,1},
    {0,1},{1,2},{2,0},{3,1},{0,3},{1,2},{2,2},{3,3},{0,2},{1,1},{2,2},{3,3},
    {0,1},{1,2},{2,3},{3,3},{0,2},{1,2},{2,2},{3,1},{0,1},{1,3},{2,1},{3,2},
};
static const uint8_t IN_TM2[12][2] = {
    {2,0},{1,2},{0,2},{3,1},{2,0},{1,3},{0,2},{1,3},{2,2},{3,2},{0,1},{1,2},
};
static const uint8_t IN_TM3[6][2] = {
    {3,2},{2,2},{1,2},{0,2},{1,3},{2,0},
};
static const uint8_t IN_TM4[24][2] = {
    {0,0},{3,1},{2,0},{1,2},{0,0},{3,1},{2,2},{1,2},{0,2},{3,1},{2,3},{1,0},
    {0,0},{1,1},{2,1},{3,2},{0,2},{1,2},{2,0},{3,3},{0,3},{1,1},{2,3},{3,2},
};

DABO_EXPORT void dabo_phase_ref_index(const dabo_mode *m, uint8_t *q)
{
    const uint8_t (*tab)[2] = m->mode == 1 ? IN_TM1 : m->mode == 2 ? IN_TM2
                            : m->mode == 3 ? IN_TM3 : IN_TM4;
    for (int blk = 0; blk < m->K / 32; blk++)
        for (int k = 0; k < 32; k++)
            q[blk * 32 + k] = (H_TAB[tab[blk][0]][k] + tab[blk][1]) & 3;
}

static cf quarter_turn(int q)
{
    static const cf v[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
    return v[q & 3];
}

DABO_EXPORT void dabo_phase_ref(const dabo_mode *m, cf *out)
{
    uint8_t *q = malloc(m->K);
    dabo_phase_ref_index(m, q);
    for (int k = 0; k < m->K; k++) out[k] = quarter_turn(q[k]);
    free(q);
}

/* ------------------------------------------------------------------------ */
/* a4  DifferentialModulator::do_process (src/DifferentialModulator.cpp:45-76)*/
/* out[0] = phase ref, out[l+1] = out[l] * in[l]; float32 complex multiply,   */
/* sequential in l like the reference.                                       */
/* ------------------------------------------------------------------------ */
static inline cf cmulf(cf a, cf b)
{
    cf r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}

DABO_EXPORT void dabo_diff_mod(const dabo_mode *m, const cf *phase, const cf *in, int nsym, cf *out)
{
    const int K = m->K;
    memcpy(out, phase, sizeof(cf) * K);
    for (int l = 0; l < nsym; l++)
        for (int k = 0; k < K; k++)
            out[(size_t)(l + 1) * K + k] = cmulf(out[(size_t)l * K + k], in[(size_t)l * K + k]);
}

/* ------------------------------------------------------------------------ */
/* a5  TII carrier set (src/TII.cpp:247-337, enable_carrier :229-245) and    */
/*     TII::process / do_process (:172-245)                                  */
/* The pattern table (TII.cpp:34-104, EN 300 401 table 52) lists the 70      */
/* 8-bit words of weight 4 in ascending order, MSB = b0.                     */
/* ------------------------------------------------------------------------ */
static int tii_pattern_bit(int pattern, int b)
{
    int n = -1;
    for (int w = 0; w < 256; w++) {
        if (__builtin_popcount(w) == 4 && ++n == pattern) return (w >> (7 - b)) & 1;
    }
    return 0;
}

/* acp[K]: 1 where a carrier pair starts. Returns -1 for unsupported modes. */
DABO_EXPORT int dabo_tii_carriers(const dabo_mode *m, int comb, int pattern, uint8_t *acp)
{
    const int K = m->K;
    memset(acp, 0, K);
    if (pattern < 0 || pattern > 69 || comb < 0 || comb > 23) return -1;
    if (m->mode == 1) {
        static const int base[4] = {-768, -384, 1, 385};
        for (int g = 0; g < 4; g++)
            for (int b = 0; b < 8; b++)
                if (tii_pattern_bit(pattern, b)) {
                    const int k = base[g] + 2 * comb + 48 * b;
                    /* each group only spans 384 carriers (loop bounds :269-307) */
                    if (k < base[g] || k >= base[g] + 384) continue;
                    const int ix = K / 2 + k + (k >= 0 ? -1 : 0);
                    if (ix < 0 || ix + 1 >= K) return -1;
                    acp[ix] = 1;
                }
    }
    else if (m->mode == 2) {
        for (int b = 0; b < 8; b++)
            if (tii_pattern_bit(pattern, b)) {
                const int k = (b < 4 ? -192 : -191) + 2 * comb + 48 * b;
                if (k < -192 || k > 192) continue;
                const int ix = K / 2 + k + (k >= 0 ? -1 : 0);
                if (ix < 0 || ix + 1 >= K) return -1;
                acp[ix] = 1;
            }
    }
    else {
        return -1;
    }
    return 0;
}

DABO_EXPORT void dabo_tii_symbol(const dabo_mode *m, const uint8_t *acp, int old_variant,
                                 const cf *phase, cf *out)
{
    memset(out, 0, sizeof(cf) * m->K);
    for (int i = 0; i < m->K; i++)
        if (acp[i]) {
            out[i] = phase[i];
            out[i + 1] = old_variant ? phase[i + 1] : phase[i];
        }
}

/* ------------------------------------------------------------------------ */
/* a6  CicEqualizer ctor + process (src/CicEqualizer.cpp:29-57, 66-91)       */
/* ------------------------------------------------------------------------ */
DABO_EXPORT void dabo_cic_filter(int K, float spacing_f, int R, float *filt)
{
    /* CicEqualizer(size_t nbCarriers, size_t spacing, int R): DabModulator.cpp:172-175 passes a float,
     * the parameter truncates it (TM III at 2.5 Msps: 312.5 -> 312) */
    const size_t spacing = (size_t)spacing_f;
    const int M = 1, Npow = 4;
    const float pi = 4.0f * atanf(1.0f);
    for (int i = 0; i < K; i++) {
        const int k = i < (K + 1) / 2 ? i + ((K & 1) ^ 1) : i - K;
        const float angle = pi * k / spacing;
        if (k == 0) {
            filt[i] = 1.0f;
        }
        else {
            float f = sinf(angle / R) / sinf(angle * M);
            f = fabsf(f) * R * M;
            filt[i] = powf(f, Npow);
        }
    }
}

/* ------------------------------------------------------------------------ */
/* DFT helper: mixed-radix (any N), double precision, unnormalised like      */
/* FFTW/KISS.  sign=+1: e^{+j2pi kn/N} (FFTW_BACKWARD), -1: forward.         */
/* ------------------------------------------------------------------------ */
typedef struct {
    int n;
    cd *tw;      /* tw[k] = exp(sign * 2 pi i k / n) for sign=+1 */
    cd *scratch;
} dft_plan;

static dft_plan *dft_plan_new(int n)
{
    dft_plan *p = malloc(sizeof(*p));
    p->n = n;
    p->tw = malloc(sizeof(cd) * n);
    p->scratch = malloc(sizeof(cd) * n);
    for (int k = 0; k < n; k++) {
        const double a = 2.0 * M_PI * (double)k / (double)n;
        p->tw[k].re = cos(a);
        p->tw[k].im = sin(a);
    }
    return p;
}

static void dft_plan_free(dft_plan *p)
{
    if (!p) return;
    free(p->tw);
    free(p->scratch);
    free(p);
}

static int smallest_factor(int n)
{
    if (n % 4 == 0) return 4;
    for (int f = 2; f * f <= n; f++)
        if (n % f == 0) return f;
    return n;
}

/* out[0..n) = DFT of in[0], in[stride], ...; tws = N/n twiddle step */
static void dft_rec(const dft_plan *p, int sign, const cd *in, int stride, cd *out, int n)
{
    if (n == 1) { out[0] = in[0]; return; }
    const int r = smallest_factor(n);
    const int m = n / r;
    const int tws = p->n / n;
    for (int q = 0; q < r; q++)
        dft_rec(p, sign, in + (size_t)q * stride, stride * r, out + (size_t)q * m, m);
    /* combine: X[k + m*t] = sum_q W_n^{q(k+m t)} Y_q[k] */
    cd tmp[64];
    cd *t = r <= 64 ? tmp : malloc(sizeof(cd) * r);
    for (int k = 0; k < m; k++) {
        for (int q = 0; q < r; q++) {
            const int ti = (int)(((long)q * k * tws) % p->n);
            cd w = p->tw[ti];
            if (sign < 0) w.im = -w.im;
            const cd y = out[(size_t)q * m + k];
            t[q].re = y.re * w.re - y.im * w.im;
            t[q].im = y.re * w.im + y.im * w.re;
        }
        for (int u = 0; u < r; u++) {
            double sr = 0, si = 0;
            for (int q = 0; q < r; q++) {
                const int ti = (int)(((long)q * u * m * tws) % p->n);
                cd w = p->tw[ti];
                if (sign < 0) w.im = -w.im;
                sr += t[q].re * w.re - t[q].im * w.im;
                si += t[q].re * w.im + t[q].im * w.re;
            }
            p->scratch[k + (size_t)u * m].re = sr;
            p->scratch[k + (size_t)u * m].im = si;
        }
    }
    /* scratch is shared across recursion levels but each level finishes its
     * use before returning, and children are complete before we write it */
    memcpy(out, p->scratch, sizeof(cd) * n);
    if (t != tmp) free(t);
}

static void dft_exec(const dft_plan *p, int sign, const cd *in, cd *out)
{
    dft_rec(p, sign, in, 1, out, p->n);
}

/* exported for the tests (checks the helper against numpy.fft) */
DABO_EXPORT void dabo_dft(int n, int sign, const double *in, double *out)
{
    dft_plan *p = dft_plan_new(n);
    dft_exec(p, sign, (const cd *)in, (cd *)out);
    dft_plan_free(p);
}

/* ------------------------------------------------------------------------ */
/* a7  OfdmGeneratorCF32::process (src/OfdmGenerator.cpp:157-308; carrier    */
/* placement :77-94) and a7' cfr_one_iteration (:310-373)                    */
/* in: nsym x K carriers, out: nsym x N samples.                             */
/* cfr_stats (optional, 2 x uint64): clipped samples, clipped errors.        */
/* ------------------------------------------------------------------------ */
/* PAPRStats::process_block (src/PAPRStats.cpp:41-70): peak and mean of |x|^2 */
static void papr_block(const cf *x, int n, double *peak, double *mean)
{
    double pk = 0, rms2 = 0;
    for (int i = 0; i < n; i++) {
        const double v = (double)(x[i].re * x[i].re + x[i].im * x[i].im);   /* std::norm of a complexf */
        if (v > pk) pk = v;
        rms2 += v;
    }
    *peak = pk; *mean = rms2 / n;
}

/* sym_stats (optional, 8 doubles per symbol, CFR only; OfdmGenerator.cpp:228-275): peak and mean power
 * before CFR, the same after, sum |before|^2, sum |after - before|^2 (MER), clipped samples, clipped errors */
#define DABO_SYM_STATS 8
DABO_EXPORT void dabo_ofdm_stats(const dabo_mode *m, const cf *in, int nsym, int cfr, float clip,
                                 float errclip, cf *out, uint64_t *cfr_stats, double *sym_stats);

DABO_EXPORT void dabo_ofdm(const dabo_mode *m, const cf *in, int nsym, int cfr, float clip,
                           float errclip, cf *out, uint64_t *cfr_stats)
{
    dabo_ofdm_stats(m, in, nsym, cfr, clip, errclip, out, cfr_stats, NULL);
}

DABO_EXPORT void dabo_ofdm_stats(const dabo_mode *m, const cf *in, int nsym, int cfr, float clip,
                                 float errclip, cf *out, uint64_t *cfr_stats, double *sym_stats)
{
    const int N = m->N, K = m->K;
    dft_plan *p = dft_plan_new(N);
    cd *X = malloc(sizeof(cd) * N), *x = malloc(sizeof(cd) * N);
    cf *Xf = malloc(sizeof(cf) * N), *sym = malloc(sizeof(cf) * N), *before = malloc(sizeof(cf) * N);
    const int pos_dst = (K & 1) ? 0 : 1, pos_size = (K + 1) / 2;
    const int neg_dst = N - K / 2, neg_src = (K + 1) / 2, neg_size = K / 2;
    uint64_t nclip = 0, nerr = 0;
    for (int s = 0; s < nsym; s++) {
        const cf *c = in + (size_t)s * K;
        memset(Xf, 0, sizeof(cf) * N);
        memcpy(Xf + pos_dst, c, sizeof(cf) * pos_size);
        memcpy(Xf + neg_dst, c + neg_src, sizeof(cf) * neg_size);
        for (int k = 0; k < N; k++) { X[k].re = Xf[k].re; X[k].im = Xf[k].im; }
        dft_exec(p, +1, X, x);
        for (int n = 0; n < N; n++) { sym[n].re = (float)x[n].re; sym[n].im = (float)x[n].im; }
        if (cfr) {
            const float clip_sq = clip * clip, err_sq = errclip * errclip;
            double *st = sym_stats ? sym_stats + (size_t)s * DABO_SYM_STATS : NULL;
            const uint64_t nclip0 = nclip, nerr0 = nerr;
            if (st) {
                papr_block(sym, N, &st[0], &st[1]);
                memcpy(before, sym, sizeof(cf) * N);
            }
            for (int n = 0; n < N; n++) {
                const float mag = sym[n].re * sym[n].re + sym[n].im * sym[n].im;
                if (mag > clip_sq) {
                    const float f = sqrtf(clip_sq / mag);
                    sym[n].re *= f; sym[n].im *= f;
                    nclip++;
                }
                x[n].re = sym[n].re; x[n].im = sym[n].im;
            }
            dft_exec(p, -1, x, X);
            for (int k = 0; k < N; k++) {
                cf pt = {(float)X[k].re / (float)N, (float)X[k].im / (float)N};
                cf e = {Xf[k].re - pt.re, Xf[k].im - pt.im};
                const float mag = e.re * e.re + e.im * e.im;
                if (mag > err_sq) {
                    const float f = sqrtf(err_sq / mag);
                    e.re *= f; e.im *= f;
                    nerr++;
                }
                X[k].re = pt.re + e.re; X[k].im = pt.im + e.im;
                /* the reference stores this sum as float32 before the IFFT */
                X[k].re = (float)X[k].re; X[k].im = (float)X[k].im;
            }
            dft_exec(p, +1, X, x);
            for (int n = 0; n < N; n++) { sym[n].re = (float)x[n].re; sym[n].im = (float)x[n].im; }
            if (st) {
                papr_block(sym, N, &st[2], &st[3]);
                double sum_iq = 0, sum_delta = 0;       /* OfdmGenerator.cpp:262-267 */
                for (int n = 0; n < N; n++) {
                    const float dr = sym[n].re - before[n].re, di = sym[n].im - before[n].im;
                    sum_iq += (double)(before[n].re * before[n].re + before[n].im * before[n].im);
                    sum_delta += (double)(dr * dr + di * di);
                }
                st[4] = sum_iq; st[5] = sum_delta;
                st[6] = (double)(nclip - nclip0); st[7] = (double)(nerr - nerr0);
            }
        }
        memcpy(out + (size_t)s * N, sym, sizeof(cf) * N);
    }
    if (cfr_stats) { cfr_stats[0] = nclip; cfr_stats[1] = nerr; }
    free(X); free(x); free(Xf); free(sym); free(before);
    dft_plan_free(p);
}

/* ------------------------------------------------------------------------ */
/* a8  GainControl::internal_process + computeGain{Fix,Max,Var}              */
/* (src/GainControl.cpp:82-192, 196-340; scalar spec :344-502)               */
/* gain_mode: 0 fix, 1 max, 2 var (GainControl.h:45).                        */
/* ------------------------------------------------------------------------ */
static float gain_of_symbol(const cf *x, int N, int gain_mode, float var_factor)
{
    if (gain_mode == 0) return 512.0f;
    if (gain_mode == 1) {
        float mn = x[0].re, mx = x[0].re;
        for (int n = 0; n < N; n++) {
            mn = fminf(mn, fminf(x[n].re, x[n].im));
            mx = fmaxf(mx, fmaxf(x[n].re, x[n].im));
        }
        const float a = fmaxf(-mn, mx);
        return (int)a != 0 ? 32767.0f / a : 1.0f;
    }
    double mr = 0, mi = 0;
    for (int n = 0; n < N; n++) { mr += x[n].re; mi += x[n].im; }
    mr /= N; mi /= N;
    double vr = 0, vi = 0;
    for (int n = 0; n < N; n++) {
        const double dr = x[n].re - mr, di = x[n].im - mi;
        vr += dr * dr; vi += di * di;
    }
    const float sr = var_factor * (float)sqrt(vr / N);
    const float si = var_factor * (float)sqrt(vi / N);
    /* NULL detection looks at lane 0 = real part only (GainControl.cpp:331) */
    return (int)sr != 0 ? 32767.0f / fmaxf(sr, si) : 1.0f;
}

DABO_EXPORT void dabo_gain(const dabo_mode *m, const cf *in, int nsym, int gain_mode,
                           float digital_gain, float normalise, float var_factor, cf *out)
{
    const int N = m->N;
    const float constant = normalise * digital_gain;
    for (int s = 0; s < nsym; s++) {
        /* the null symbol borrows the gain of the next symbol (:139-144) */
        const cf *stat = in + (size_t)(s > 0 ? s : 1) * N;
        const float g = gain_of_symbol(stat, N, gain_mode, var_factor) * constant;
        for (int n = 0; n < N; n++) {
            out[(size_t)s * N + n].re = in[(size_t)s * N + n].re * g;
            out[(size_t)s * N + n].im = in[(size_t)s * N + n].im * g;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* a9  GuardIntervalInserter do_process<complexf>                            */
/* (src/GuardIntervalInserter.cpp:96-113 window, :115-323)                   */
/* in: (L+1) x N, out: null_size + L*sym_size samples.  W = windowOverlap.   */
/* ------------------------------------------------------------------------ */
DABO_EXPORT void dabo_guard(const dabo_mode *m, const cf *in, int W, cf *out)
{
    const int N = m->N, L = m->L;
    if (W == 0) {
        const cf *x = in;
        cf *o = out;
        int size = m->null_size;
        for (int l = 0; l <= L; l++) {
            const int pre = size - N;
            memcpy(o, x + N - pre, sizeof(cf) * pre);
            memcpy(o + pre, x, sizeof(cf) * N);
            x += N; o += size; size = m->sym_size;
        }
        return;
    }
    float *w = malloc(sizeof(float) * 2 * W);
    for (int i = 0; i < 2 * W; i++)
        w[i] = (float)(0.5 * (1.0 - cos(M_PI * i / (2 * W - 1))));
    const long total = m->tf_samples;
    cf *acc = calloc(total + W, sizeof(cf));
    long pos = 0;
    for (int l = 0; l <= L; l++) {
        const int size = l == 0 ? m->null_size : m->sym_size;
        const int pre = size - N;
        const cf *x = in + (size_t)l * N;
        const int first = l == 0, last = l == L;
        /* extended symbol covers [pos - (first?0:W), pos + size + (last?0:W)) */
        for (long o = first ? 0 : -W; o < size + (last ? 0 : W); o++) {
            /* sample index in the cyclic extension: output offset o maps to
             * x[(o - pre) mod N] */
            long ix = (o - pre) % N;
            if (ix < 0) ix += N;
            float g = 1.0f;
            if (!first && o < W) g = w[o + W];
            if (!last && o >= size - W) g = w[2 * W - 1 - (o - (size - W))];
            cf v = {x[ix].re * g, x[ix].im * g};
            if (!first && o < W) {
                acc[pos + o].re += v.re; acc[pos + o].im += v.im;
            }
            else {
                acc[pos + o] = v;
            }
        }
        pos += size;
    }
    memcpy(out, acc, sizeof(cf) * total);
    free(acc);
    free(w);
}

/* ------------------------------------------------------------------------ */
/* a10  FIRFilter::internal_process (src/FIRFilter.cpp:144-192)              */
/* Anti-causal, truncated at the frame end, I and Q filtered separately;     */
/* float32 accumulation in tap order like the reference's inner loop.        */
/* ------------------------------------------------------------------------ */
DABO_EXPORT void dabo_fir(const cf *in, long nsamp, const float *taps, int ntaps, cf *out)
{
    const float *x = (const float *)in;
    float *y = (float *)out;
    const long n = 2 * nsamp;
    for (long i = 0; i < n; i++) {
        float acc = 0.0f;
        for (int j = 0; j < ntaps && i + 2L * j < n; j++)
            acc += x[i + 2L * j] * taps[j];
        y[i] = acc;
    }
}

/* The built-in taps of FIRFilter.cpp:59-71 ("default"), i.e. the output of
 * doc/fir-filter/generate-filter.py for fs 2.048e6, cutoff 810e3, transition
 * 250e3: symmetric, 45 taps.  Half + centre listed; mirrored on load. */
static const float FIR_DEFAULT_HALF[23] = {
    -0.00110450468492f, 0.00120703084394f, -0.000840645749122f, -0.000187368263141f,
    0.00184351124335f, -0.00355578539893f, 0.00419321097434f, -0.00254214904271f,
    -0.00183473504148f, 0.00781436730176f, -0.0125957569107f, 0.0126200336963f,
    -0.00537294941023f, -0.00866683479398f, 0.0249746385962f, -0.0356550291181f,
    0.0319730602205f, -0.00795613788068f, -0.0363943465054f, 0.0938014090061f,
    -0.151176810265f, 0.193567320704f, 0.791776955128f,
};

DABO_EXPORT int dabo_fir_default_taps(float *taps)
{
    for (int i = 0; i < 23; i++) { taps[i] = FIR_DEFAULT_HALF[i]; taps[44 - i] = FIR_DEFAULT_HALF[i]; }
    return 45;
}

/* ------------------------------------------------------------------------ */
/* a11  Resampler ctor + process (src/Resampler.cpp:51-112, 131-195)         */
/* ------------------------------------------------------------------------ */
typedef struct {
    int ni, no;
    long L, M;
    float factor;
    float *window;
    cf *buf_in;    /* ni/2 */
    cf *buf_out;   /* no/2 */
    dft_plan *pin, *pout;
} dabo_resampler;

static long gcd_l(long a, long b) { return b == 0 ? a : gcd_l(b, a % b); }

DABO_EXPORT dabo_resampler *dabo_resampler_new(long in_rate, long out_rate, int resolution)
{
    dabo_resampler *r = calloc(1, sizeof(*r));
    const long g = gcd_l(in_rate, out_rate);
    r->L = out_rate / g;
    r->M = in_rate / g;
    long factor = (long)resolution * 2 / r->M;
    if (factor & 1) ++factor;
    r->ni = (int)(factor * r->M);
    r->no = (int)(factor * r->L);
    if (r->ni > r->no) r->factor = 1.0f / r->ni * out_rate / in_rate;
    else               r->factor = 1.0f / r->no * out_rate / in_rate;
    r->window = malloc(sizeof(float) * r->ni);
    for (int i = 0; i < r->ni; i++)
        r->window[i] = (float)(0.5 * (1.0 - cos(2.0 * M_PI * i / (r->ni - 1))));
    r->buf_in = calloc(r->ni / 2, sizeof(cf));
    r->buf_out = calloc(r->no / 2, sizeof(cf));
    r->pin = dft_plan_new(r->ni);
    r->pout = dft_plan_new(r->no);
    return r;
}

DABO_EXPORT void dabo_resampler_free(dabo_resampler *r)
{
    if (!r) return;
    free(r->window); free(r->buf_in); free(r->buf_out);
    dft_plan_free(r->pin); dft_plan_free(r->pout);
    free(r);
}

DABO_EXPORT int dabo_resampler_sizes(const dabo_resampler *r, int *ni, int *no)
{
    *ni = r->ni; *no = r->no;
    return 0;
}

/* nsamp must be a multiple of ni/2 (true for every DAB mode); returns the
 * number of output samples = nsamp * L / M */
DABO_EXPORT long dabo_resampler_process(dabo_resampler *r, const cf *in, long nsamp, cf *out)
{
    const int ni = r->ni, no = r->no, hi = ni / 2, ho = no / 2;
    cd *a = malloc(sizeof(cd) * ni), *F = malloc(sizeof(cd) * ni);
    cd *B = malloc(sizeof(cd) * no), *y = malloc(sizeof(cd) * no);
    cf *Ff = malloc(sizeof(cf) * ni), *Bf = malloc(sizeof(cf) * no);
    long j = 0;
    for (long i = 0; i + hi <= nsamp; i += hi, j += ho) {
        for (int k = 0; k < hi; k++) {
            a[k].re = r->buf_in[k].re * r->window[k];
            a[k].im = r->buf_in[k].im * r->window[k];
            a[hi + k].re = in[i + k].re * r->window[hi + k];
            a[hi + k].im = in[i + k].im * r->window[hi + k];
            /* products are float32 in the reference */
            a[k].re = (float)a[k].re; a[k].im = (float)a[k].im;
            a[hi + k].re = (float)a[hi + k].re; a[hi + k].im = (float)a[hi + k].im;
        }
        memcpy(r->buf_in, in + i, sizeof(cf) * hi);
        dft_exec(r->pin, -1, a, F);
        for (int k = 0; k < ni; k++) { Ff[k].re = (float)F[k].re; Ff[k].im = (float)F[k].im; }
        memset(Bf, 0, sizeof(cf) * no);
        if (no > ni) {
            memcpy(Bf, Ff, sizeof(cf) * hi);
            memcpy(Bf + no - hi, Ff + hi, sizeof(cf) * hi);
            Bf[hi] = Ff[hi];
        }
        else {
            memcpy(Bf, Ff, sizeof(cf) * ho);
            memcpy(Bf + ho, Ff + ni - ho, sizeof(cf) * ho);
            Bf[ho].re += Ff[ho].re; Bf[ho].im += Ff[ho].im;
            Bf[ho].re *= 0.5f; Bf[ho].im *= 0.5f;
        }
        for (int k = 0; k < no; k++) {
            B[k].re = Bf[k].re * r->factor; B[k].im = Bf[k].im * r->factor;
            B[k].re = (float)B[k].re; B[k].im = (float)B[k].im;
        }
        dft_exec(r->pout, +1, B, y);
        for (int k = 0; k < ho; k++) {
            out[j + k].re = r->buf_out[k].re + (float)y[k].re;
            out[j + k].im = r->buf_out[k].im + (float)y[k].im;
            r->buf_out[k].re = (float)y[ho + k].re;
            r->buf_out[k].im = (float)y[ho + k].im;
        }
    }
    free(a); free(F); free(B); free(y); free(Ff); free(Bf);
    return j;
}

/* ------------------------------------------------------------------------ */
/* a12  MemlessPoly apply_coeff / apply_lut (src/MemlessPoly.cpp:237-309)    */
/* The "cos" series constants are the reference's, verbatim (0.486666, not   */
/* 1/24).                                                                    */
/* ------------------------------------------------------------------------ */
DABO_EXPORT void dabo_poly(const cf *in, long n, const float *am, const float *pm, cf *out)
{
    for (long i = 0; i < n; i++) {
        const float mag = in[i].re * in[i].re + in[i].im * in[i].im;
        const float amp = am[0] + mag * (am[1] + mag * (am[2] + mag * (am[3] + mag * am[4])));
        const float ph = -1 * (pm[0] + mag * (pm[1] + mag * (pm[2] + mag * (pm[3] + mag * pm[4]))));
        const float p2 = ph * ph;
        const float re = 1.0f - p2 * (-0.5f + p2 * (0.486666f + p2 * (-0.00138888f)));
        const float im = ph * (1.0f + p2 * (0.166666f + p2 * 0.00833333f));
        const cf a = {in[i].re * amp, in[i].im * amp};
        const cf c = {re, im};
        out[i] = cmulf(a, c);
    }
}

/* LUT of 32 REAL entries (the loader assigns a float to each complex entry,
 * MemlessPoly.cpp:203-210) */
DABO_EXPORT void dabo_lut(const cf *in, long n, const float *lut, float scalefactor, cf *out)
{
    for (long i = 0; i < n; i++) {
        const float mag = hypotf(in[i].re, in[i].im);
        const uint32_t scaled = (uint32_t)lrintf(mag * scalefactor);
        const float g = lut[scaled >> 27];
        out[i].re = in[i].re * g;
        out[i].im = in[i].im * g;
    }
}

/* ------------------------------------------------------------------------ */
/* a13  FormatConverter::process, float input (src/FormatConverter.cpp:112-165)*/
/* fmt: 1 = s16, 2 = u8, 3 = s8.  Returns the clipped-sample count.          */
/* ------------------------------------------------------------------------ */
DABO_EXPORT uint64_t dabo_format(const float *in, long n, int fmt, void *out)
{
    uint64_t clipped = 0;
    for (long i = 0; i < n; i++) {
        if (fmt == 1) {
            int16_t *o = out;
            if (in[i] < -32768.0f) { o[i] = INT16_MIN; clipped++; }
            else if (in[i] > 32767.0f) { o[i] = INT16_MAX; clipped++; }
            else o[i] = (int16_t)in[i];
        }
        else if (fmt == 2) {
            uint8_t *o = out;
            const float s = in[i] + 128.0f;
            if (s < 0) { o[i] = 0; clipped++; }
            else if (s > 255.0f) { o[i] = 255; clipped++; }
            else o[i] = (uint8_t)s;
        }
        else {
            int8_t *o = out;
            if (in[i] < -128.0f) { o[i] = INT8_MIN; clipped++; }
            else if (in[i] > 127.0f) { o[i] = INT8_MAX; clipped++; }
            else o[i] = (int8_t)in[i];
        }
    }
    return clipped;
}

/* ------------------------------------------------------------------------ */
/* Whole chain in the reference's wiring order (src/DabModulator.cpp:386-417)*/
/* ------------------------------------------------------------------------ */
typedef struct {
    int32_t  mode;
    int32_t  gain_mode;
    uint64_t output_rate;     /* 0 or 2048000: no resampler */
    uint64_t clock_rate;      /* 0: no CicEqualizer */
    float    digital_gain;
    float    normalise;
    float    gain_variance;
    int32_t  window_overlap;
    int32_t  cfr_enable;
    float    cfr_clip;
    float    cfr_errclip;
    int32_t  tii_enable;
    int32_t  tii_comb;
    int32_t  tii_pattern;
    int32_t  tii_old_variant;
    int32_t  fir_ntaps;       /* 0: no FIR */
    const float *fir_taps;
    int32_t  poly_mode;       /* 0 none, 1 odd polynomial (10 coefs), 2 LUT (scale + 32) */
    const float *poly_coefs;
    int32_t  format;          /* 0 complexf, 1 s16, 2 u8, 3 s8 */
    int32_t  fixed_point;     /* 1 = FFTEngine::KISS: fixed_oracle.c after the multiplexer, int16 I/Q out */
} dabo_cfg;

/* fixed_oracle.c */
void dabo_fix_carriers(const float *z, long n_floats, int16_t *out);
void dabo_ofdm_fixed(int N, int K, const int16_t *in, int nsym, int16_t *out);
void dabo_guard_fixed(int N, int L, int null_size, int sym_size, const int16_t *in, int W, int16_t *out);

typedef struct {
    dabo_cfg c;
    dabo_mode m;
    float *fir_taps;
    float poly[33];
    int use_cic;
    float *cic;
    int tii_ok;
    uint8_t *acp;
    int tii_insert;
    dabo_resampler *rs;
    uint64_t clipped, cfr_clip_count, cfr_err_count;
    /* CFR read-outs (OfdmGenerator.cpp:186-306): PAPR windows of (L+1)*50 blocks, the last 10 frames'
     * clip / error-clip ratios and MERs */
    double *papr_b, *papr_a;          /* [2 * window]: (peak, mean) pairs, oldest first */
    int papr_nb, papr_na, papr_window;
    double clip_r[10], err_r[10], mer[10];
    int n_clip_r, n_err_r, n_mer;
    int mer_index, papr_clear;
} dabo_chain;

static void readout_push(double *d, int *n, int cap, double v)
{
    if (*n == cap) { memmove(d, d + 1, sizeof(double) * (cap - 1)); (*n)--; }
    d[(*n)++] = v;
}

static void papr_push(double *w, int *n, int window, double peak, double mean)
{
    if (*n == window) { memmove(w, w + 2, sizeof(double) * 2 * (window - 1)); (*n)--; }
    w[2 * *n] = peak; w[2 * *n + 1] = mean; (*n)++;
}

/* PAPRStats::calculate_papr (src/PAPRStats.cpp:72-101) */
static double papr_calc(const double *w, int n, int window)
{
    if (n < window) return 0;
    double peak = 0, rms2 = 0;
    for (int i = 0; i < n; i++) {
        if (w[2 * i] > peak) peak = w[2 * i];
        rms2 += w[2 * i + 1];
    }
    rms2 /= n;
    return 10.0 * log10(peak / rms2);
}

/* one frame of per-symbol statistics into the read-outs (OfdmGenerator.cpp:196-306) */
static void chain_cfr_frame(dabo_chain *h, const double *st, int nsym)
{
    const int N = h->m.N;
    h->mer_index = (h->mer_index + 1) % nsym;
    if (h->papr_clear) { h->papr_nb = h->papr_na = 0; h->papr_clear = 0; }
    double nclip = 0, nerr = 0;
    for (int i = 0; i < nsym; i++) {
        const double *r = st + (size_t)i * DABO_SYM_STATS;
        papr_push(h->papr_b, &h->papr_nb, h->papr_window, r[0], r[1]);
        if (i > 0) papr_push(h->papr_a, &h->papr_na, h->papr_window, r[2], r[3]);
        if (i > 0 && h->mer_index == i)
            readout_push(h->mer, &h->n_mer, 10, r[5] > 0 ? 10.0 * log10(r[4] / r[5]) : 90);
        nclip += r[6]; nerr += r[7];
    }
    readout_push(h->clip_r, &h->n_clip_r, 10, nclip / ((double)nsym * N));
    readout_push(h->err_r, &h->n_err_r, 10, nerr / ((double)nsym * N));
}

/* OfdmGeneratorCF32::get_parameter "clip_stats" / "papr" (OfdmGenerator.cpp:419-453) */
DABO_EXPORT int dabo_chain_readout(const dabo_chain *h, const char *name, char *buf, int cap)
{
    if (!strcmp(name, "clip_stats")) {
        if (!h->n_clip_r || !h->n_err_r || !h->n_mer) return snprintf(buf, cap, "No stats available");
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < h->n_clip_r; i++) a += h->clip_r[i];
        for (int i = 0; i < h->n_err_r; i++) b += h->err_r[i];
        for (int i = 0; i < h->n_mer; i++) c += h->mer[i];
        return snprintf(buf, cap, "Statistics : %f%% samples clipped, %f%% errors clipped. MER after CFR: %f dB",
                        a / h->n_clip_r * 100, b / h->n_err_r * 100, c / h->n_mer);
    }
    if (!strcmp(name, "papr")) {
        const double pb = papr_calc(h->papr_b, h->papr_nb, h->papr_window);
        const double pa = papr_calc(h->papr_a, h->papr_na, h->papr_window);
        char sb[64], sa[64];
        if (pb == 0) snprintf(sb, sizeof sb, "N/A"); else snprintf(sb, sizeof sb, "%f", pb);
        if (pa == 0) snprintf(sa, sizeof sa, "N/A"); else snprintf(sa, sizeof sa, "%f", pa);
        return snprintf(buf, cap, "PAPR [dB]: %s, %s", sb, sa);
    }
    return -1;
}

DABO_EXPORT dabo_chain *dabo_chain_new(const dabo_cfg *c)
{
    dabo_chain *h = calloc(1, sizeof(*h));
    h->c = *c;
    if (dabo_mode_params(c->mode, &h->m)) { free(h); return NULL; }
    if (c->fir_ntaps > 0) {
        h->fir_taps = malloc(sizeof(float) * c->fir_ntaps);
        memcpy(h->fir_taps, c->fir_taps, sizeof(float) * c->fir_ntaps);
    }
    if (c->poly_mode == 1) memcpy(h->poly, c->poly_coefs, sizeof(float) * 10);
    if (c->poly_mode == 2) memcpy(h->poly, c->poly_coefs, sizeof(float) * 33);
    h->papr_window = (h->m.L + 1) * 50;      /* OfdmGenerator.cpp:59-61 */
    h->papr_b = calloc((size_t)2 * h->papr_window, sizeof(double));
    h->papr_a = calloc((size_t)2 * h->papr_window, sizeof(double));
    const uint64_t rate = c->output_rate ? c->output_rate : 2048000;
    /* DabModulator.cpp:154-176 */
    if (c->clock_rate) {
        unsigned ratio = (unsigned)(c->clock_rate / rate) / 4;
        if (c->clock_rate == 400000000) h->use_cic = ratio & 1;
        else h->use_cic = 1;
        if (h->use_cic) {
            h->cic = malloc(sizeof(float) * h->m.K);
            dabo_cic_filter(h->m.K, (float)h->m.N * (float)rate / 2048000.0f, (int)ratio, h->cic);
        }
    }
    /* DabModulator.cpp:178-190: TII exists for TM I/II, NullSymbol otherwise */
    h->acp = calloc(h->m.K, 1);
    h->tii_ok = dabo_tii_carriers(&h->m, c->tii_comb, c->tii_pattern, h->acp) == 0;
    h->tii_insert = 1;
    if (rate != 2048000) h->rs = dabo_resampler_new(2048000, (long)rate, h->m.N);
    return h;
}

DABO_EXPORT void dabo_chain_free(dabo_chain *h)
{
    if (!h) return;
    free(h->fir_taps); free(h->cic); free(h->acp); free(h->papr_b); free(h->papr_a);
    dabo_resampler_free(h->rs);
    free(h);
}

DABO_EXPORT long dabo_chain_out_samples(const dabo_chain *h)
{
    long n = h->m.tf_samples;
    if (h->rs) n = n * h->rs->L / h->rs->M;
    return n;
}

DABO_EXPORT uint64_t dabo_chain_clipped(const dabo_chain *h) { return h->clipped; }

/* stage: 0 = final; otherwise stop after 1 qpsk, 2 freq, 3 diff, 4 mux,
 * 5 ciceq, 6 ofdm, 7 gain, 8 guard, 9 fir, 10 resampler, 11 poly.
 * Returns the number of bytes written to out (or -1). */
DABO_EXPORT long dabo_chain_process(dabo_chain *h, const uint8_t *bits, int stage, void *out)
{
    const dabo_mode *m = &h->m;
    const int K = m->K, N = m->N, L = m->L;
    long ret = -1;
    cf *q = malloc(sizeof(cf) * (size_t)(L - 1) * K);
    cf *f = malloc(sizeof(cf) * (size_t)(L - 1) * K);
    cf *ref = malloc(sizeof(cf) * K);
    cf *z = malloc(sizeof(cf) * (size_t)(L + 1) * K);
    cf *x = malloc(sizeof(cf) * (size_t)(L + 1) * N);
    cf *y = malloc(sizeof(cf) * (size_t)(L + 1) * N);
    cf *tf = malloc(sizeof(cf) * (size_t)m->tf_samples);
    cf *tf2 = NULL, *rs = NULL;
    long nout = m->tf_samples;

    dabo_qpsk(m, bits, L - 1, q);
    if (stage == 1) { ret = sizeof(cf) * (L - 1) * K; memcpy(out, q, ret); goto done; }
    dabo_freq_interleave(m, q, L - 1, f);
    if (stage == 2) { ret = sizeof(cf) * (L - 1) * K; memcpy(out, f, ret); goto done; }
    dabo_phase_ref(m, ref);
    dabo_diff_mod(m, ref, f, L - 1, z + K);
    if (stage == 3) { ret = sizeof(cf) * L * K; memcpy(out, z + K, ret); goto done; }
    /* null or TII symbol in slot 0 (SignalMultiplexer.cpp:61-68); TII toggles
     * every call (TII.cpp:225-242) */
    memset(z, 0, sizeof(cf) * K);
    if (h->tii_ok) {
        if (h->c.tii_enable && h->tii_insert)
            dabo_tii_symbol(m, h->acp, h->c.tii_old_variant, ref, z);
        h->tii_insert = !h->tii_insert;
    }
    if (h->c.fixed_point) {
        /* DabModulator.cpp:144-224 with fftEngine == KISS: complexfix carriers, OfdmGeneratorFixed, no
         * GainControl, GuardIntervalInserter<complexfix>; the output is already s16 */
        int16_t *zi = malloc(sizeof(int16_t) * 2 * (size_t)(L + 1) * K);
        int16_t *xi = malloc(sizeof(int16_t) * 2 * (size_t)(L + 1) * N);
        dabo_fix_carriers((const float *)z, 2L * (L + 1) * K, zi);
        if (stage >= 1 && stage <= 4) { ret = 4L * (L + 1) * K; memcpy(out, zi, ret); }
        else {
            dabo_ofdm_fixed(N, K, zi, L + 1, xi);
            if (stage == 6) { ret = 4L * (L + 1) * N; memcpy(out, xi, ret); }
            else {
                dabo_guard_fixed(N, L, m->null_size, m->sym_size, xi, h->c.window_overlap, out);
                ret = 4L * m->tf_samples;
            }
        }
        free(zi); free(xi);
        goto done;
    }
    if (stage == 4) { ret = sizeof(cf) * (L + 1) * K; memcpy(out, z, ret); goto done; }
    if (h->use_cic) {
        for (int s = 0; s <= L; s++)
            for (int k = 0; k < K; k++) {
                z[(size_t)s * K + k].re *= h->cic[k];
                z[(size_t)s * K + k].im *= h->cic[k];
            }
    }
    if (stage == 5) { ret = sizeof(cf) * (L + 1) * K; memcpy(out, z, ret); goto done; }
    {
        uint64_t st[2];
        double *ss = h->c.cfr_enable ? calloc((size_t)(L + 1) * DABO_SYM_STATS, sizeof(double)) : NULL;
        dabo_ofdm_stats(m, z, L + 1, h->c.cfr_enable, h->c.cfr_clip, h->c.cfr_errclip, x, st, ss);
        h->cfr_clip_count = st[0]; h->cfr_err_count = st[1];
        if (ss) { chain_cfr_frame(h, ss, L + 1); free(ss); }
    }
    if (stage == 6) { ret = sizeof(cf) * (L + 1) * N; memcpy(out, x, ret); goto done; }
    dabo_gain(m, x, L + 1, h->c.gain_mode, h->c.digital_gain, h->c.normalise, h->c.gain_variance, y);
    if (stage == 7) { ret = sizeof(cf) * (L + 1) * N; memcpy(out, y, ret); goto done; }
    dabo_guard(m, y, h->c.window_overlap, tf);
    if (stage == 8) { ret = sizeof(cf) * nout; memcpy(out, tf, ret); goto done; }
    if (h->fir_taps) {
        tf2 = malloc(sizeof(cf) * nout);
        dabo_fir(tf, nout, h->fir_taps, h->c.fir_ntaps, tf2);
        cf *t = tf; tf = tf2; tf2 = t;
    }
    if (stage == 9) { ret = sizeof(cf) * nout; memcpy(out, tf, ret); goto done; }
    if (h->rs) {
        rs = malloc(sizeof(cf) * (size_t)dabo_chain_out_samples(h));
        nout = dabo_resampler_process(h->rs, tf, nout, rs);
    }
    {
        cf *cur = rs ? rs : tf;
        if (stage == 10) { ret = sizeof(cf) * nout; memcpy(out, cur, ret); goto done; }
        if (h->c.poly_mode == 1) dabo_poly(cur, nout, h->poly, h->poly + 5, cur);
        else if (h->c.poly_mode == 2) dabo_lut(cur, nout, h->poly + 1, h->poly[0], cur);
        if (stage == 11 || h->c.format == 0) { ret = sizeof(cf) * nout; memcpy(out, cur, ret); goto done; }
        h->clipped = dabo_format((const float *)cur, 2 * nout, h->c.format, out);
        ret = 2 * nout * (h->c.format == 1 ? 2 : 1);
    }
done:
    free(q); free(f); free(ref); free(z); free(x); free(y); free(tf); free(tf2); free(rs);
    return ret;
}
