// k_resample_up: the TM I resampler for integer up-sampling ratios (M == 1:
// 4.096 / 6.144 / 8.192 Msps ...), Ni = 4096, No = L * Ni.
// Reference: src/Resampler.cpp:51-112 (geometry, window, factor), :131-195 (process).
//
// Same hop-independent formulation as resample.cuh (block c_b from input halves
// b-2, b-1, b), plus the structure of zero padding in frequency: with
// F = FFT_Ni(c_b) re-laid out on No = L*Ni bins (Resampler.cpp:153-164, the Nyquist
// bin copied to both sides), output sample n = L*m + rho is
//     y[L m + rho] = factor * IFFT_Ni( F[k] * e^{j 2 pi k' rho / No} )[m]
// k' = k for k < Ni/2, k - Ni above, and bin Ni/2 weighted 2 cos(pi rho / L).
// So the No-point inverse transform is L transforms of Ni points, and the one for
// rho = 0 is the block itself: y[L m] = factor * (Ni c_b[m] + (-1)^m F[Ni/2]).
// Per hop: 1 forward + (L-1) inverse 4096-point FFTs instead of 4096 + 16384 points
// (-30 % flops at L = 4), everything in shared memory, only the first Ni/2 outputs
// of every transform are kept (the hop-independent form needs no overlap-add).
//
// A CTA is two teams of 256 threads, each working on its own hop with its own
// buffers and named barrier; a 4096-point transform is three radix-16 passes
// (fft.cuh fft16), 16 points per thread.  Twiddles: only W^(k 2^i), i < 4, are tabled,
// the other powers are products (11 complex multiplies per pass) -- the LSU pipe,
// not the FP32 pipe, is the scarcer resource here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft.cuh"
#include "kernels.cuh"
#include "resample.cuh"

namespace dabmod {

constexpr int RU_NI = 4096;
constexpr int RU_HI = RU_NI / 2;
constexpr int RU_TEAM = 256;                  // threads per team = butterflies per pass
constexpr int RU_TEAMS = 2;
constexpr int RU_THREADS = RU_TEAM * RU_TEAMS;
constexpr int RU_MAX_L = 4;
constexpr int RU_BUF = RU_NI + RU_NI / 16;    // spad

struct RuSmem {
    float2 tw2[4 * 16];                       // pass 2 (Ns = 16):  tw2[i*16 + k]  = e^{+j 2 pi k 2^i / 256}
    float2 tw3[4 * 256];                      // pass 3 (Ns = 256): tw3[i*256 + k] = e^{+j 2 pi k 2^i / 4096}
    float2 nyq[RU_TEAMS];                     // F[Ni/2] of the team's current hop
    float2 buf[RU_TEAMS][RU_BUF];
    float2 stage[RU_TEAMS][RU_MAX_L][RU_HI];  // y[L m + rho] at [rho][m]
};

struct RuParams {
    ResParams r;                              // shared with the generic kernel (ni, no, factor, in, hist, win, tw_out, out, post)
    int L;                                    // no / ni
};

__device__ __forceinline__ void ru_bar(int team)
{
    asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(RU_TEAM) : "memory");
}

// w[r] = w1^r for r = 1..15 from w1, w2, w4, w8
__device__ __forceinline__ void ru_powers(const float2 *tab, int Ns, int k, float2 (&w)[16])
{
    w[1] = tab[k]; w[2] = tab[Ns + k]; w[4] = tab[2 * Ns + k]; w[8] = tab[3 * Ns + k];
    w[3] = cmul(w[1], w[2]);
    w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]);
    w[9] = cmul(w[8], w[1]); w[10] = cmul(w[8], w[2]); w[11] = cmul(w[8], w[3]);
    w[12] = cmul(w[8], w[4]); w[13] = cmul(w[8], w[5]); w[14] = cmul(w[8], w[6]); w[15] = cmul(w[8], w[7]);
}

// 4096-point transform of the 16 values of butterfly t (element t + 256 r in v[r]), in place
// through the team's buffer; on return v[r] = X[t + 256 r].
template <bool INV>
__device__ __forceinline__ void ru_fft4096(float2 (&v)[16], float2 *buf, const float2 *tw2, const float2 *tw3, int t,
                                           int team)
{
    // pass 1 (Ns = 1): no twiddles, out[t*16 + r]
    fft16<INV>(v);
    ru_bar(team);                             // the previous user of the buffer is done reading
#pragma unroll
    for (int r = 0; r < 16; r++) buf[spad(t * 16 + r)] = v[r];
    ru_bar(team);
    // pass 2 (Ns = 16)
    {
        const int k = t & 15;
#pragma unroll
        for (int r = 0; r < 16; r++) v[r] = buf[spad(t + 256 * r)];
        float2 w[16];
        ru_powers(tw2, 16, k, w);
#pragma unroll
        for (int r = 1; r < 16; r++) v[r] = cmul(v[r], tw_dir<INV>(w[r]));
        fft16<INV>(v);
        ru_bar(team);
        const int j0 = (t - k) * 16 + k;
#pragma unroll
        for (int r = 0; r < 16; r++) buf[spad(j0 + 16 * r)] = v[r];
        ru_bar(team);
    }
    // pass 3 (Ns = 256): k = t, output X[t + 256 r] stays in registers
    {
#pragma unroll
        for (int r = 0; r < 16; r++) v[r] = buf[spad(t + 256 * r)];
        float2 w[16];
        ru_powers(tw3, 256, t, w);
#pragma unroll
        for (int r = 1; r < 16; r++) v[r] = cmul(v[r], tw_dir<INV>(w[r]));
        fft16<INV>(v);
    }
}

template <bool POST>
__global__ void __launch_bounds__(RU_THREADS, 1) k_resample_up(const __grid_constant__ RuParams pu)
{
    const ResParams &p = pu.r;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RuSmem &sm = *reinterpret_cast<RuSmem *>(smem_raw);
    const int tid = threadIdx.x, team = tid / RU_TEAM, t = tid - team * RU_TEAM;
    const int L = pu.L, no = p.no;
    constexpr int hi = RU_HI;

    // twiddle tables from the Ni-point root table (tw_in[k] = e^{+j 2 pi k / 4096})
    for (int i = tid; i < 4 * 16; i += RU_THREADS) sm.tw2[i] = __ldg(p.tw_in + ((i & 15) << (i >> 4)) * 16);
    for (int i = tid; i < 4 * 256; i += RU_THREADS) sm.tw3[i] = __ldg(p.tw_in + (((i & 255) << (i >> 8)) & (RU_NI - 1)));
    __syncthreads();

    float2 *buf = sm.buf[team];
    unsigned clip = 0;
    const long long team0 = (long long)blockIdx.x * RU_TEAMS + team;
    const long long n_teams = (long long)gridDim.x * RU_TEAMS;
    // every team runs the same number of rounds (a team without a hop idles through the barriers)
    const long long rounds = (p.total_hops + n_teams - 1) / n_teams;
    for (long long rd = 0; rd < rounds; rd++) {
        const long long hop = team0 + rd * n_teams;
        const bool live = hop < p.total_hops;
        const long long base = (live ? hop : 0) * hi;

        // ---- block c_b straight into the registers of the first pass: element t + 256 r ----
        float2 v[16];
        float2 c_lo[8];                       // c_b[t + 256 r], r < 8, kept for the rho = 0 outputs
        if (p.dbg & 1) {
#pragma unroll
            for (int r = 0; r < 16; r++) v[r] = make_float2((float)(t + r), (float)(t - r));
#pragma unroll
            for (int r = 0; r < 8; r++) c_lo[r] = v[r];
        }
        else if (base >= 2 * hi) {
            // (all hops but the first two of a launch: the three input halves are in `in`)
            const float2 *src = p.in + base + t;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int m = t + 256 * r;
                const float w0 = __ldg(p.win + m), w1 = __ldg(p.win + hi + m);
                const float2 a = __ldg(src + 256 * r - hi);        // H_{b-1}[m]
                const float2 b = __ldg(src + 256 * r);             // H_b[m]
                const float2 c = __ldg(src + 256 * r - 2 * hi);    // H_{b-2}[m]
                v[r] = cfma(w0, a, cscale(a, w1));
                v[r + 8] = cfma(w1, b, cscale(c, w0));
                c_lo[r] = v[r];
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
                v[r] = cfma(w0, a, cscale(a, w1));
                v[r + 8] = cfma(w1, b, cscale(c, w0));
                c_lo[r] = v[r];
            }
        }
        if (!(p.dbg & 8)) ru_fft4096<false>(v, buf, sm.tw2, sm.tw3, t, team);
        // v[r] = F[t + 256 r].  Keep it; publish the Nyquist bin (t = 0, r = 8).
        float2 F[16];
#pragma unroll
        for (int r = 0; r < 16; r++) F[r] = v[r];
        if (t == 0) sm.nyq[team] = F[8];
        ru_bar(team);
        {
            // rho = 0: the block itself plus the second copy of the Nyquist bin
            const float2 ny = sm.nyq[team];
            const float sgn = (t & 1) ? -1.0f : 1.0f;     // (-1)^m, m = t + 256 r has the parity of t
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const float2 y = cfma((float)RU_NI, c_lo[r], cscale(ny, sgn));
                sm.stage[team][0][t + 256 * r] = cscale(y, p.factor);
            }
        }
        float2 wt = __ldg(p.tw_out + t);      // e^{j 2 pi t rho / No} of the next phase, fetched one phase ahead
        for (int rho = 1; rho < ((p.dbg & 4) ? 1 : L); rho++) {
            // G_rho[k] = F[k] e^{j 2 pi k' rho / No}, tw_out[i] = e^{+j 2 pi i / No}.  The thread's bins are
            // k' = t + 256 r (r < 8) and t + 256 (r - 16) (r >= 8): a geometric sequence in r with ratio
            // c = e^{j 2 pi 256 rho / No} -- one table value per phase, the rest by multiplication.
            const float2 c = __ldg(p.tw_out + 256 * rho), cc = make_float2(c.x, -c.y);
            float2 root = wt;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                v[r] = cmul(F[r], root);
                root = cmul(root, c);
            }
            root = cmul(wt, cc);
#pragma unroll
            for (int r = 15; r > 8; r--) {
                v[r] = cmul(F[r], root);
                root = cmul(root, cc);
            }
            // r = 8: k' = t - 2048; bin Ni/2 (t = 0) sits on both sides: e^{+j a} + e^{-j a} = 2 cos a
            if (t == 0) root = make_float2(root.x + root.x, 0.f);
            v[8] = cmul(F[8], root);
            wt = __ldg(p.tw_out + t * (rho + 1));
            ru_fft4096<true>(v, buf, sm.tw2, sm.tw3, t, team);
#pragma unroll
            for (int r = 0; r < 8; r++)
                sm.stage[team][rho][t + 256 * r] = cscale(v[r], p.factor);
        }
        ru_bar(team);
        // ---- interleave the L phases and store: out[(hop*hi + m) * L + rho] ----
        if (live && !(p.dbg & 2)) {
            const size_t obase = (size_t)hop * hi * L;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int m = t + 256 * r;
                if (L == 4) {
                    // the four phases of input sample m are four consecutive output samples: 32 bytes per lane
                    store_run4<POST>(p.out, obase + (size_t)m * 4, sm.stage[team][0][m], sm.stage[team][1][m],
                                     sm.stage[team][2][m], sm.stage[team][3][m], p.post, clip);
                }
                else {
                    for (int rho = 0; rho < L; rho++)
                        store_sample<POST>(p.out, obase + (size_t)m * L + rho, sm.stage[team][rho][m], p.post, clip);
                }
            }
        }
        // (the next round's first write to `stage` comes after several team barriers)
    }
    if (POST && p.post.format != 0) flush_clip(p.post, clip);
}

// ---------------------------------------------------------------------------------------------------------
// k_resample_up3: the same transform with THREE teams per SM.
//
// Measured on k_resample_up (two teams of 128-register threads, 99 KB of shared memory each; tools/res_dbg.py,
// 256 TFs): transforms only 0.92 ms, input loads + epilogue + stores only 0.59 ms, together 1.53 ms -- the
// sum, not the maximum.  While a team waits for its 24 input loads or drains its staging through the epilogue,
// the other team's eight warps cannot fill the FP32 pipe alone (48 % busy over the launch, 80 % when both
// compute).  A third team needs the registers and the shared memory of the first two to shrink:
//   * the spectrum F lives in shared memory (each thread parks its own 16 bins and re-reads them per phase:
//     private slots, no barrier) instead of 32 registers;
//   * no output staging: the thread that ends phase rho's last pass with X_rho[t + 256 r] is the thread that
//     owns output samples L m + rho of m = t + 256 r, so phases leave in pairs (rho 0 | 1, 2 | 3) as 16-byte
//     stores straight from registers, the even phase of a pair parked in 16 registers for one transform;
//   * pass-2 twiddles come from a 15 x 16 table instead of 11 complex products per pass.
// 80 registers x 768 threads, 67 KB per team.
// ---------------------------------------------------------------------------------------------------------
constexpr int RU3_TEAMS = 3;
constexpr int RU3_THREADS = RU_TEAM * RU3_TEAMS;

struct Ru3Smem {
    float2 tw2[15 * 16];                      // pass 2 (Ns = 16):  tw2[(r-1)*16 + k] = e^{+j 2 pi k r / 256}
    float2 tw3[4 * 256];                      // pass 3 (Ns = 256): tw3[i*256 + k] = e^{+j 2 pi k 2^i / 4096}
    float2 nyq[RU3_TEAMS];
    float2 buf[RU3_TEAMS][RU_BUF];
    float2 spec[RU3_TEAMS][RU_NI];            // F[t + 256 r] at [t + 256 r]: thread-private slots
};

template <bool INV>
__device__ __forceinline__ void ru3_fft4096(float2 (&v)[16], float2 *buf, const float2 *tw2, const float2 *tw3, int t,
                                            int team)
{
    fft16<INV>(v);
    ru_bar(team);                             // the previous user of the buffer is done reading
#pragma unroll
    for (int r = 0; r < 16; r++) buf[spad(t * 16 + r)] = v[r];
    ru_bar(team);
    {
        const int k = t & 15;
#pragma unroll
        for (int r = 0; r < 16; r++) v[r] = buf[spad(t + 256 * r)];
#pragma unroll
        for (int r = 1; r < 16; r++) v[r] = cmul(v[r], tw_dir<INV>(tw2[(r - 1) * 16 + k]));
        fft16<INV>(v);
        ru_bar(team);
        const int j0 = (t - k) * 16 + k;
#pragma unroll
        for (int r = 0; r < 16; r++) buf[spad(j0 + 16 * r)] = v[r];
        ru_bar(team);
    }
    {
#pragma unroll
        for (int r = 0; r < 16; r++) v[r] = buf[spad(t + 256 * r)];
        // powers of the thread's root as they are needed (few live at a time)
        const float2 w1 = tw_dir<INV>(tw3[t]), w2 = tw_dir<INV>(tw3[256 + t]);
        const float2 w4 = tw_dir<INV>(tw3[512 + t]), w8 = tw_dir<INV>(tw3[768 + t]);
        v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[4] = cmul(v[4], w4); v[8] = cmul(v[8], w8);
        const float2 w3 = cmul(w1, w2);
        v[3] = cmul(v[3], w3);
        v[12] = cmul(v[12], cmul(w8, w4));
        const float2 w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
        v[5] = cmul(v[5], w5); v[6] = cmul(v[6], w6); v[7] = cmul(v[7], w7);
        v[9] = cmul(v[9], cmul(w8, w1)); v[10] = cmul(v[10], cmul(w8, w2)); v[11] = cmul(v[11], cmul(w8, w3));
        v[13] = cmul(v[13], cmul(w8, w5)); v[14] = cmul(v[14], cmul(w8, w6)); v[15] = cmul(v[15], cmul(w8, w7));
        fft16<INV>(v);
    }
}

template <bool POST, int L>
__global__ void __launch_bounds__(RU3_THREADS, 1) k_resample_up3(const __grid_constant__ RuParams pu)
{
    const ResParams &p = pu.r;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ru3Smem &sm = *reinterpret_cast<Ru3Smem *>(smem_raw);
    const int tid = threadIdx.x, team = tid / RU_TEAM, t = tid - team * RU_TEAM;
    constexpr int hi = RU_HI;

    for (int i = tid; i < 15 * 16; i += RU3_THREADS) sm.tw2[i] = __ldg(p.tw_in + (((i & 15) * (i / 16 + 1)) & 255) * 16);
    for (int i = tid; i < 4 * 256; i += RU3_THREADS) sm.tw3[i] = __ldg(p.tw_in + (((i & 255) << (i >> 8)) & (RU_NI - 1)));
    __syncthreads();

    float2 *buf = sm.buf[team];
    float2 *spec = sm.spec[team] + t;
    unsigned clip = 0;
    const long long team0 = (long long)blockIdx.x * RU3_TEAMS + team;
    const long long n_teams = (long long)gridDim.x * RU3_TEAMS;
    const long long rounds = (p.total_hops + n_teams - 1) / n_teams;   // idle teams still walk the barriers
    for (long long rd = 0; rd < rounds; rd++) {
        const long long hop = team0 + rd * n_teams;
        const bool live = hop < p.total_hops;
        const long long base = (live ? hop : 0) * hi;

        float2 v[16];
        float2 y0[8];                         // c_b[t + 256 r], then the rho = 0 outputs
        if (base >= 2 * hi) {
            const float2 *src = p.in + base + t;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int m = t + 256 * r;
                const float w0 = __ldg(p.win + m), w1 = __ldg(p.win + hi + m);
                const float2 a = __ldg(src + 256 * r - hi);        // H_{b-1}[m]
                const float2 b = __ldg(src + 256 * r);             // H_b[m]
                const float2 c = __ldg(src + 256 * r - 2 * hi);    // H_{b-2}[m]
                v[r] = cfma(w0, a, cscale(a, w1));
                v[r + 8] = cfma(w1, b, cscale(c, w0));
                y0[r] = v[r];
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
                v[r] = cfma(w0, a, cscale(a, w1));
                v[r + 8] = cfma(w1, b, cscale(c, w0));
                y0[r] = v[r];
            }
        }
        {
            // the team's next hop: its three input halves (48 KB, contiguous) towards L2 while this one computes
            const long long nb = (hop + n_teams) * hi;
            if (hop + n_teams < p.total_hops && t < 192) {
                const char *q = reinterpret_cast<const char *>(p.in + nb - 2 * hi) + t * 256;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 128));
            }
        }
        ru3_fft4096<false>(v, buf, sm.tw2, sm.tw3, t, team);
        // v[r] = F[t + 256 r]: parked in the thread's own slots; the Nyquist bin (t = 0, r = 8) goes to everyone
#pragma unroll
        for (int r = 0; r < 16; r++) spec[256 * r] = v[r];
        if (t == 0) sm.nyq[team] = v[8];
        ru_bar(team);
        {
            // rho = 0: the block itself plus the second copy of the Nyquist bin
            const float2 ny = sm.nyq[team];
            const float sgn = (t & 1) ? -1.0f : 1.0f;     // (-1)^m, m = t + 256 r has the parity of t
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const float2 y = cfma((float)RU_NI, y0[r], cscale(ny, sgn));
                y0[r] = cscale(y, p.factor);
            }
        }
        const size_t obase = (size_t)hop * hi * L + (size_t)t * L;     // sample L m of m = t: + 256 r L per r
        const bool pairs = (L & 1) == 0;
        if (!pairs && live && !(p.dbg & 2)) {
#pragma unroll
            for (int r = 0; r < 8; r++) store_sample<POST>(p.out, obase + (size_t)256 * r * L, y0[r], p.post, clip);
        }
        float2 wt = __ldg(p.tw_out + t);      // e^{j 2 pi t rho / No} of the next phase, fetched one phase ahead
#pragma unroll 1
        for (int rho = 1; rho < ((p.dbg & 4) ? 1 : L); rho++) {
            // G_rho[k] = F[k] e^{j 2 pi k' rho / No}: see k_resample_up
            const float2 c = __ldg(p.tw_out + 256 * rho), cc = make_float2(c.x, -c.y);
            float2 root = wt;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                v[r] = cmul(spec[256 * r], root);
                root = cmul(root, c);
            }
            root = cmul(wt, cc);
#pragma unroll
            for (int r = 15; r > 8; r--) {
                v[r] = cmul(spec[256 * r], root);
                root = cmul(root, cc);
            }
            if (t == 0) root = make_float2(root.x + root.x, 0.f);
            v[8] = cmul(spec[256 * 8], root);
            wt = __ldg(p.tw_out + t * (rho + 1));
            ru3_fft4096<true>(v, buf, sm.tw2, sm.tw3, t, team);
#pragma unroll
            for (int r = 0; r < 8; r++) v[r] = cscale(v[r], p.factor);
            if (pairs) {
                if (rho & 1) {
                    // y0 holds phase rho - 1: two consecutive output samples per m
                    if (live && !(p.dbg & 2)) {
#pragma unroll
                        for (int r = 0; r < 8; r++)
                            store_run2<POST>(p.out, obase + (size_t)256 * r * L + (rho - 1), y0[r], v[r], p.post, clip);
                    }
                }
                else {
#pragma unroll
                    for (int r = 0; r < 8; r++) y0[r] = v[r];
                }
            }
            else if (live && !(p.dbg & 2)) {
#pragma unroll
                for (int r = 0; r < 8; r++)
                    store_sample<POST>(p.out, obase + (size_t)256 * r * L + rho, v[r], p.post, clip);
            }
        }
    }
    if (POST && p.post.format != 0) flush_clip(p.post, clip);
}

} // namespace dabmod
