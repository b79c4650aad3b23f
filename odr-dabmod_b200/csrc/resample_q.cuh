// k_resample_q: the TM I resampler for the rational up-sampling ratios whose output
// transform is a multiple of 4000 points: Ni = 4096, No = P * 4000, P = 2..5, i.e.
// 4.0 / 6.0 / 8.0 / 10.0 Msps (10 Msps: L/M = 625/128, No = 20000 -- BASELINE config 5).
// Reference: src/Resampler.cpp:51-112 (geometry, window, factor), :131-195 (process).
//
// Same hop-independent formulation as resample.cuh (block c_b from the input halves
// b-2, b-1, b; no overlap-add state), and the same idea as resample_up.cuh: the
// No-point inverse transform of a spectrum that occupies only the bins k' in
// [-Ni/2, Ni/2] splits by output phase.  With n = P m + rho, Q = No / P = 4000:
//     y[P m + rho] = factor * IFFT_Q( G_rho )[m],
//     G_rho[q]     = sum over k' = q (mod Q) of B[k'] e^{j 2 pi k' rho / No}
// B[k'] = F[k' mod Ni] is the re-laid-out spectrum (Resampler.cpp:153-164: the input
// Nyquist bin appears at both k' = +Ni/2 and k' = -Ni/2).  4097 bins on 4000 slots:
// the slots q in [1952, 2048] receive two bins each (k' = q and k' = q - 4000), every
// other slot exactly one, none is empty.  Only m < Q/2 is kept (the first half of the
// No-point transform).  Per hop: one forward 4096-point FFT (three radix-16 passes,
// shared with k_resample_up) and P inverse 4000-point FFTs = radix 20 x 20 x 10, each
// radix a prime-factor butterfly without internal twiddles (fft.cuh), instead of one
// 20000-point transform through an L2 scratch: 1.45 instead of 1.67 MFLOP per hop and
// nothing leaves the SM but the result.
//
// A CTA is two teams of 256 threads, each on its own hop with its own buffers and
// named barrier (200 threads of a team carry the 20-point butterflies).  The phase
// twiddles e^{j 2 pi k' rho / No} of a thread's 16 bins are a geometric sequence
// (ratio e^{j 2 pi 256 rho / No}): one table value per thread and phase, the rest by
// multiplication.  The phases 0..P-2 are staged in shared memory, the last one stays in
// the FFT buffer (the last pass works in place on its own slots), and the hop's P*2000
// output samples go out interleaved as fully coalesced stores with the MemlessPoly /
// FormatConverter epilogue.
//
// The per-thread stages are plain functions over a `float2 *buf`; with DABMOD_FN
// redefined they compile for the host, where tests/test_fft_host.py runs a whole hop
// thread by thread against numpy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft.cuh"

namespace dabmod {

constexpr int RQ_NI = 4096;
constexpr int RQ_HI = RQ_NI / 2;
constexpr int RQ_Q = 4000;                    // points of a phase transform
constexpr int RQ_KEEP = RQ_Q / 2;             // outputs kept per phase
constexpr int RQ_ACTIVE = RQ_Q / 20;          // threads of a team that carry a 20-point butterfly
constexpr int RQ_MIN_P = 2, RQ_MAX_P = 5;
constexpr int RQ_FOLD0 = RQ_Q - RQ_HI;        // 1952: first slot that receives two bins
constexpr int RQ_NFOLD = RQ_HI - RQ_FOLD0 + 1; // 97
constexpr int RQ_SHIFT = RQ_NI - RQ_Q;        // 96: slot of a negative bin = its FFT index - 96
constexpr int RQ_S1 = 201;                    // row stride of the layout between pass 1 and pass 2
constexpr int RQ_S2 = 404;                    // row stride of the layout between pass 2 and pass 3, and of the results
constexpr int RQ_STAGE_STRIDE = 5 * RQ_S2;    // one phase's 2000 results: element m at 404 (m / 400) + m % 400
constexpr int RQ_BUF = RQ_NI + RQ_NI / 16;    // the forward transform's spad(4096); the layouts below need <= 4036

// Shared-memory layouts of the 4000-point transform (8-byte elements, a half-warp is conflict free when its
// 16 slots differ mod 16).  Every exchange has its own layout so that both sides are contiguous or odd-strided
// without gaps (a padded natural order costs a second wavefront wherever a run of lanes crosses a pad slot):
//   spread -> pass 1:  slot q                    stores: consecutive q;  loads: q = t + 200 r
//   pass 1 -> pass 2:  element 20 t + r at 201 r + t         stores: consecutive t;  loads: 201 k + g + 10 r'
//   pass 2 -> pass 3:  element 400 g + j at 404 g + j        stores: j = k + 20 r, k fastest over the lanes, the
//                      next g continues the run mod 16 (404 = 4 mod 16, 20 = 4 mod 16);  loads: 404 r + b
//   results:           X[400 r + b] at 404 r + b, in place on the slots the butterfly has just read
__device__ __host__ __forceinline__ constexpr int rq_result_slot(int m) { return RQ_S2 * (m / 400) + m % 400; }

// Per-phase uniform constants (shared memory, RqPhase ph[P]): c = w^(256 rho), d = w^(-160 rho), e = w^(96 rho),
// w = e^{+j 2 pi / No}
struct RqPhase { float2 c, d, e; float2 pad_; };

// G_rho into buf (natural order).  Thread t holds F[t + 256 r] in F[r]; fp7 / fp8 are the bins that
// fold onto its slots 1792 + t (t >= 160) and 2048 (t == 0): F[k + 96].  wt = w^(t rho).  The twiddles
// w^(k' rho) of the thread's 16 bins k' = t + 256 r (r < 8), t + 256 (r - 16) (r >= 8) follow from wt by
// repeated multiplication with c resp. its conjugate: no table look-ups in the loop.
template <bool FIRST>
DABMOD_FN void rq_spread(const float2 (&F)[16], float2 fp7, float2 fp8, int t, float2 wt, const RqPhase &ph,
                         float2 *buf)
{
    if (FIRST) {                              // rho = 0: all twiddles are 1
#pragma unroll
        for (int r = 0; r < 8; r++) {
            float2 g = F[r];
            if (r == 7 && t >= 160) g = cadd(g, fp7);
            buf[t + 256 * r] = g;
        }
#pragma unroll
        for (int r = 8; r < 16; r++) {
            if (r == 8 && t < RQ_NFOLD) {
                if (t == 0) buf[RQ_HI] = cadd(F[8], fp8);
            }
            else buf[t + 256 * r - RQ_SHIFT] = F[r];
        }
        return;
    }
    const float2 c = ph.c, cc = make_float2(c.x, -c.y);
    float2 root = wt;
#pragma unroll
    for (int r = 0; r < 7; r++) {             // k' = k = t + 256 r >= 0, slot k
        buf[t + 256 * r] = cmul(F[r], root);
        root = cmul(root, c);
    }
    float2 g7 = cmul(F[7], root);             // slots >= 1952 also receive the bin k' = k - 4000: below
    root = cmul(wt, cc);                      // k' = t - 256
#pragma unroll
    for (int r = 15; r > 8; r--) {            // k' = k - 4096 < -1952, slot k' + 4000 = k - 96
        buf[t + 256 * r - RQ_SHIFT] = cmul(F[r], root);
        root = cmul(root, cc);
    }
    // root = w^((t - 2048) rho): FFT indices 2048..2303
    if (t >= RQ_NFOLD) buf[t + 256 * 8 - RQ_SHIFT] = cmul(F[8], root);
    else if (t == 0) {
        // the Nyquist bin as k' = +2048 (twiddle = conjugate of w^(-2048 rho)) plus the bin k' = -1952 (FFT index 2144)
        const float2 g = cmul(F[8], make_float2(root.x, -root.y));
        const float2 h = cmul(fp8, cmul(root, ph.e));
        buf[RQ_HI] = cadd(g, h);
    }
    // (FFT indices 2049..2144 are negative bins that fold: added by the owners of the slots 1953..2047 here)
    if (t >= 160) g7 = cadd(g7, cmul(fp7, cmul(root, ph.d)));   // k' = (t - 2048) - 160
    buf[t + 256 * 7] = g7;
}

// twiddle powers w^r from the tabled w^(2^i): tab[i * Ns + k]
DABMOD_FN void rq_powers20(const float2 *tab, int k, float2 (&w)[20])
{
    w[1] = tab[k]; w[2] = tab[20 + k]; w[4] = tab[40 + k]; w[8] = tab[60 + k]; w[16] = tab[80 + k];
    w[3] = cmul(w[1], w[2]);
    w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]);
#pragma unroll
    for (int r = 1; r < 8; r++) w[8 + r] = cmul(w[8], w[r]);
    w[17] = cmul(w[16], w[1]); w[18] = cmul(w[16], w[2]); w[19] = cmul(w[16], w[3]);
}

// Inverse 4000-point transform, Stockham radix 20 (Ns = 1), 20 (Ns = 20), 10 (Ns = 400).
// pass 1, butterfly t < 200: x[t + 200 r] -> 20-point transform (no twiddles)
DABMOD_FN void rq_pass1_load(const float2 *buf, int t, float2 (&v)[20])
{
#pragma unroll
    for (int r = 0; r < 20; r++) v[r] = buf[t + 200 * r];
    fft20<true>(v);
}
DABMOD_FN void rq_pass1_store(float2 *buf, int t, const float2 (&v)[20])
{
#pragma unroll
    for (int r = 0; r < 20; r++) buf[RQ_S1 * r + t] = v[r];
}
// pass 2, butterfly t = 20 g + k < 200: elements t + 200 r' = 20 (g + 10 r') + k, twiddles e^{+j 2 pi k r' / 400}
DABMOD_FN void rq_pass2_load(const float2 *buf, const float2 *itw2, int t, float2 (&v)[20])
{
    const int g = t / 20, k = t - 20 * g;
    const float2 *src = buf + RQ_S1 * k + g;
#pragma unroll
    for (int r = 0; r < 20; r++) v[r] = src[10 * r];
    float2 w[20];
    rq_powers20(itw2, k, w);
#pragma unroll
    for (int r = 1; r < 20; r++) v[r] = cmul(v[r], w[r]);
    fft20<true>(v);
}
DABMOD_FN void rq_pass2_store(float2 *buf, int t, const float2 (&v)[20])
{
    const int g = t / 20, k = t - 20 * g;
    float2 *dst = buf + RQ_S2 * g + k;        // element 400 g + k + 20 r
#pragma unroll
    for (int r = 0; r < 20; r++) dst[20 * r] = v[r];
}
// pass 3, butterfly b < 400 (k = b): elements b + 400 r, twiddles e^{+j 2 pi k r / 4000};
// on return x[r] = X[b + 400 r], r < 5, to be stored at dst[404 r + b]
DABMOD_FN void rq_pass3(const float2 *buf, const float2 *itw3, int b, float2 (&x)[10])
{
#pragma unroll
    for (int r = 0; r < 10; r++) x[r] = buf[RQ_S2 * r + b];
    const float2 w1 = itw3[b], w2 = itw3[400 + b], w4 = itw3[800 + b], w8 = itw3[1200 + b];
    const float2 w3 = cmul(w1, w2);
    x[1] = cmul(x[1], w1); x[2] = cmul(x[2], w2); x[3] = cmul(x[3], w3); x[4] = cmul(x[4], w4);
    x[5] = cmul(x[5], cmul(w4, w1)); x[6] = cmul(x[6], cmul(w4, w2)); x[7] = cmul(x[7], cmul(w4, w3));
    x[8] = cmul(x[8], w8); x[9] = cmul(x[9], cmul(w8, w1));
    fft10_lo<true>(x);
}

#if defined(__CUDACC__) && !defined(RQ_HOST_ONLY)
} // namespace dabmod
#include "kernels.cuh"
#include "resample.cuh"
#include "resample_up.cuh"
namespace dabmod {

constexpr int RQ_TEAM = RU_TEAM;              // 256
constexpr int RQ_TEAMS = 2;
constexpr int RQ_THREADS = RQ_TEAM * RQ_TEAMS;

struct RqSmem {
    float2 tw2[4 * 16];                       // forward pass 2 / 3 tables, as in RuSmem
    float2 tw3[4 * 256];
    float2 itw2[5 * 20];                      // itw2[i*20 + k]  = e^{+j 2 pi k 2^i / 400}
    float2 itw3[4 * 400];                     // itw3[i*400 + k] = e^{+j 2 pi k 2^i / 4000}
    RqPhase ph[RQ_MAX_P];                     // per-phase twiddle constants
    float2 fold[RQ_TEAMS][RQ_NFOLD + 1];      // F[2048 .. 2144] of the team's current hop
    float2 buf[RQ_TEAMS][RQ_BUF];
    float2 stage[RQ_TEAMS][RQ_MAX_P - 1][RQ_STAGE_STRIDE];  // phase rho < P-1 at [rho][m]; phase P-1 ends in buf
};

struct RqParams {
    ResParams r;                              // ni, no, factor, in, hist, win, tw_in, tw_out, out, post
    int P;                                    // no / 4000
};

template <bool POST>
__global__ void __launch_bounds__(RQ_THREADS, 1) k_resample_q(const __grid_constant__ RqParams pq)
{
    const ResParams &p = pq.r;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RqSmem &sm = *reinterpret_cast<RqSmem *>(smem_raw);
    const int tid = threadIdx.x, team = tid / RQ_TEAM, t = tid - team * RQ_TEAM;
    const int P = pq.P, no = p.no;
    constexpr int hi = RQ_HI;

    // tables: tw_in[k] = e^{+j 2 pi k / 4096}, tw_out[i] = e^{+j 2 pi i / No}, No = 400 * 10 P = 4000 * P
    for (int i = tid; i < 4 * 16; i += RQ_THREADS) sm.tw2[i] = __ldg(p.tw_in + ((i & 15) << (i >> 4)) * 16);
    for (int i = tid; i < 4 * 256; i += RQ_THREADS) sm.tw3[i] = __ldg(p.tw_in + (((i & 255) << (i >> 8)) & (RQ_NI - 1)));
    for (int i = tid; i < 5 * 20; i += RQ_THREADS) sm.itw2[i] = __ldg(p.tw_out + (((i % 20) << (i / 20)) % 400) * 10 * P);
    for (int i = tid; i < 4 * 400; i += RQ_THREADS) sm.itw3[i] = __ldg(p.tw_out + (((i % 400) << (i / 400)) % 4000) * P);
    if (tid < P) {
        const int rho = tid;
        sm.ph[rho].c = __ldg(p.tw_out + 256 * rho);
        sm.ph[rho].d = __ldg(p.tw_out + (no - 160 * rho) % no);
        sm.ph[rho].e = __ldg(p.tw_out + 96 * rho);
    }
    __syncthreads();

    float2 *buf = sm.buf[team];
    unsigned clip = 0;
    // The two teams run the same code on the same amount of data: started together they load together, transform
    // together and store together, and nobody computes while both wait for memory.  Half a hop apart they overlap
    // (measured: 1.13 -> 1.10 ms per 128 TFs).
    if (team == 1) __nanosleep(8192);
    const long long team0 = (long long)blockIdx.x * RQ_TEAMS + team;
    const long long n_teams = (long long)gridDim.x * RQ_TEAMS;
    const long long rounds = (p.total_hops + n_teams - 1) / n_teams;   // idle teams still walk the barriers
    for (long long rd = 0; rd < rounds; rd++) {
        const long long hop = team0 + rd * n_teams;
        const bool live = hop < p.total_hops;
        const long long base = (live ? hop : 0) * hi;

        // ---- block c_b into the registers of the forward transform's first pass: element t + 256 r ----
        float2 F[16];
        if (p.dbg & 1) {
#pragma unroll
            for (int r = 0; r < 16; r++) F[r] = make_float2((float)(t + r), (float)(t - r));
        }
        else if (base >= 2 * hi) {
            const float2 *src = p.in + base + t;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int m = t + 256 * r;
                const float w0 = __ldg(p.win + m), w1 = __ldg(p.win + hi + m);
                const float2 a = __ldg(src + 256 * r - hi);        // H_{b-1}[m]
                const float2 b = __ldg(src + 256 * r);             // H_b[m]
                const float2 c = __ldg(src + 256 * r - 2 * hi);    // H_{b-2}[m]
                F[r] = cfma(w0, a, cscale(a, w1));
                F[r + 8] = cfma(w1, b, cscale(c, w0));
            }
        }
        else {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int m = t + 256 * r;
                const float w0 = __ldg(p.win + m), w1 = __ldg(p.win + hi + m);
                const float2 a = res_load(p, base - hi + m);
                const float2 b = res_load(p, base + m);
                const float2 c = res_load(p, base - 2 * hi + m);
                F[r] = cfma(w0, a, cscale(a, w1));
                F[r + 8] = cfma(w1, b, cscale(c, w0));
            }
        }
        {
            // the team's next hop: its three input halves (48 KB, contiguous) towards L2 while this one computes
            const long long nb = (hop + n_teams) * hi;
            if (hop + n_teams < p.total_hops && t < 192 && !(p.dbg & 32)) {
                const char *q = reinterpret_cast<const char *>(p.in + nb - 2 * hi) + t * 256;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 128));
            }
        }
        if (!(p.dbg & 8)) ru_fft4096<false>(F, buf, sm.tw2, sm.tw3, t, team);
        // F[r] = spectrum bin t + 256 r, scaled once here (Resampler.cpp:179-181)
#pragma unroll
        for (int r = 0; r < 16; r++) F[r] = cscale(F[r], p.factor);
        // the bins that fold onto the slots 1952..2048 change hands
        if (t < RQ_NFOLD) sm.fold[team][t] = F[8];
        ru_bar(team);
        const float2 fp7 = t >= 160 ? sm.fold[team][t - 160] : make_float2(0.f, 0.f);
        const float2 fp8 = sm.fold[team][RQ_NFOLD - 1];

        float2 wt = __ldg(p.tw_out + t);      // w^(t rho) of the next phase, fetched one phase ahead
        for (int rho = 0; rho < ((p.dbg & 4) ? 1 : P); rho++) {
            ru_bar(team);                     // the previous user of buf is done reading
            if (rho == 0) rq_spread<true>(F, fp7, fp8, t, wt, sm.ph[0], buf);
            else {
                rq_spread<false>(F, fp7, fp8, t, wt, sm.ph[rho], buf);
                wt = __ldg(p.tw_out + t * (rho + 1));   // < 256 * 5 entries: stays in L1
            }
            ru_bar(team);
            float2 v[20];
            if (t < RQ_ACTIVE) rq_pass1_load(buf, t, v);
            ru_bar(team);
            if (t < RQ_ACTIVE) rq_pass1_store(buf, t, v);
            ru_bar(team);
            if (t < RQ_ACTIVE) rq_pass2_load(buf, sm.itw2, t, v);
            ru_bar(team);
            if (t < RQ_ACTIVE) rq_pass2_store(buf, t, v);
            ru_bar(team);
            // pass 3 works in place on its own ten slots, so the last phase may stay in buf
            float2 *dst = rho < P - 1 ? sm.stage[team][rho] : buf;
            if (t < RQ_ACTIVE) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    float2 x[10];
                    rq_pass3(buf, sm.itw3, t + 200 * u, x);
#pragma unroll
                    for (int r = 0; r < 5; r++) dst[RQ_S2 * r + t + 200 * u] = x[r];
                }
            }
        }
        ru_bar(team);
        // ---- interleave the P phases and store: out[hop * P * 2000 + P m + rho] ----
        if (live && !(p.dbg & 2)) {
            const int n_out = P * RQ_KEEP;
            const size_t obase = (size_t)hop * n_out;
            // e / P by multiplication: exact for e < 16384 and P = 2..5 (ceil(2^16 / P) over-estimates by < 1 / (P e))
            const unsigned inv = (65536u + P - 1) / P;
            auto result = [&](int e) {                       // output sample e of the hop = phase e % P of m = e / P
                const int m = (int)(((unsigned)e * inv) >> 16), rho = e - m * P;
                return (rho == P - 1 ? buf : sm.stage[team][rho])[rq_result_slot(m)];
            };
            for (int e = 2 * t; e < n_out; e += 2 * RQ_TEAM)  // n_out and the hop's first output index are even
                store_run2<POST>(p.out, obase + e, result(e), result(e + 1), p.post, clip);
        }
        // (the next round's first write to buf / stage comes after several team barriers)
    }
    if (POST && p.post.format != 0) flush_clip(p.post, clip);
}
#endif

} // namespace dabmod
