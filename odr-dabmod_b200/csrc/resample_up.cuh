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
        if (base >= 2 * hi) {
            // (all hops but the first two of a launch: the three input halves are in `in`)
            const float2 *src = p.in + base + t;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int m = t + 256 * r;
                const float w0 = __ldg(p.win + m), w1 = __ldg(p.win + hi + m);
                const float2 a = __ldg(src + 256 * r - hi);        // H_{b-1}[m]
                const float2 b = __ldg(src + 256 * r);             // H_b[m]
                const float2 c = __ldg(src + 256 * r - 2 * hi);    // H_{b-2}[m]
                v[r] = make_float2(fmaf(w0, a.x, w1 * a.x), fmaf(w0, a.y, w1 * a.y));
                v[r + 8] = make_float2(fmaf(w1, b.x, w0 * c.x), fmaf(w1, b.y, w0 * c.y));
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
                v[r] = make_float2(fmaf(w0, a.x, w1 * a.x), fmaf(w0, a.y, w1 * a.y));
                v[r + 8] = make_float2(fmaf(w1, b.x, w0 * c.x), fmaf(w1, b.y, w0 * c.y));
                c_lo[r] = v[r];
            }
        }
        ru_fft4096<false>(v, buf, sm.tw2, sm.tw3, t, team);
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
                const float2 y = make_float2(fmaf((float)RU_NI, c_lo[r].x, sgn * ny.x),
                                             fmaf((float)RU_NI, c_lo[r].y, sgn * ny.y));
                sm.stage[team][0][t + 256 * r] = make_float2(y.x * p.factor, y.y * p.factor);
            }
        }
        float2 wt = __ldg(p.tw_out + t);      // e^{j 2 pi t rho / No} of the next phase, fetched one phase ahead
        for (int rho = 1; rho < L; rho++) {
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
                sm.stage[team][rho][t + 256 * r] = make_float2(v[r].x * p.factor, v[r].y * p.factor);
        }
        ru_bar(team);
        // ---- interleave the L phases and store: out[(hop*hi + m) * L + rho] ----
        if (live) {
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

} // namespace dabmod
