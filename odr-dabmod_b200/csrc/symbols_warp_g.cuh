// k_symbols_wg: the warp-per-symbol structure of k_symbols_w (symbols_warp.cuh) for TM IV (N = 1024) and TM II
// (N = 512) as "64 / R1 symbols per warp", R1 = N / 32.  (TM III, N = 256, 8 symbols per warp with an odd sample
// count per symbol -- so plain stores instead of bulk copies -- was built and measured twice: 0.66 ms per 4096 TFs with
// the warps in step, 0.70 ms with them in three groups; k_symbols runs it in 0.65 ms and keeps it.)
//
// Same chain and the same reference lines as k_symbols_w (QpskSymbolMapper.cpp:105-156,
// FrequencyInterleaver.cpp:103-126, DifferentialModulator.cpp:45-76, SignalMultiplexer.cpp:45-71,
// OfdmGenerator.cpp:157-308, GainControl.cpp:82-340, GuardIntervalInserter.cpp:301-319), plain configuration
// (no CicEq, CFR, windowing or TII frame: those run in k_symbols).
//
// A warp takes G = 64 / R1 CONSECUTIVE symbols of a transmission frame per iteration (2 / 4 for TM IV / II), i.e.
// 2048 points like one TM I symbol, so the machinery of k_symbols_w carries over unchanged:
//   * every lane holds 64 points in registers: for each of the G symbols the bins lane + 32 r, r < R1; pass 1 is
//     G register transforms of R1 points (fft32 / fft16) instead of one of 64;
//   * one exchange through shared memory with the same addressing (lane stride 65);
//   * pass 2: two radix-32 butterflies per lane (64 per warp = R1 per symbol), twiddles W_N^(l q);
//   * the only synchronisation is __syncwarp; the scaled symbols are staged in natural order and leave through
//     cp.async.bulk (symbol starts are 16-byte aligned in both modes).
// K = 3 N / 4 in every mode, so G symbols have 1536 carriers like one TM I symbol: R1 / 2 lanes own one symbol's
// carriers, 48 each, and a lane's share of a bit row is the same 6 + 6 bytes as in TM I.
//
// Differential modulation across the G symbols of a group: every lane computes the phase increments of ITS OWN
// symbol's bit row (one row fetch per lane and iteration, fetched one iteration ahead, like TM I) and an inclusive
// scan over the lanes that own the same carriers in the G symbols (warp shuffles, nibble-wise addition mod 8)
// turns them into the G phases; the last one is the base of the next group.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft_reg.cuh"
#include "kernels.cuh"
#include "symbols_warp.cuh"

namespace dabmod {

template <int R1>
struct SymWgSmem {
    float2 tw[31 * R1];                     // second pass: tw[(l-1)*R1 + q] = e^{+j 2 pi l q / N}
    uint32_t bin_t[SW_CPL / 2 * 32];        // code index of the lane's source carriers within its symbol, two per word
    uint32_t spread[256];
    uint32_t ph0[6 * 32];                   // phase reference of the lane's carriers, nibble packed: [word][lane]
    float2 c8[16];
    float2 x[SW_WARPS][32 * SW_XPAD];       // exchange buffer, then the staging area of the G scaled symbols
    uint8_t code[SW_WARPS][SW_CODES];       // phase codes of the G symbols being assembled: [g * K + index]
};

struct SymWgParams {
    SymParams s;
    const float2 *twiddle;                  // N entries e^{+j 2 pi k / N}
    int n_tf;
    int n_groups;                           // groups per TF = ceil(L / G)
};

template <int R1>
__device__ __forceinline__ void wg_fft_r1(float2 (&v)[R1])
{
    static_assert(R1 == 32 || R1 == 16, "TM IV and TM II");
    if (R1 == 32) fft32<true>(reinterpret_cast<float2 (&)[32]>(v));
    else fft16<true>(&v[0]);
}

template <int R1, bool POST>
__global__ void __launch_bounds__(SW_THREADS, 1) k_symbols_wg(const __grid_constant__ SymWgParams pw)
{
    const SymParams &p = pw.s;
    constexpr int N = 32 * R1, K = 24 * R1, G = 64 / R1;
    constexpr int LPS = 32 / G;             // lanes that own one symbol's carriers
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SymWgSmem<R1> &sm = *reinterpret_cast<SymWgSmem<R1> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gs = lane / LPS, ls = lane - gs * LPS;     // the lane's symbol within a group, its place among the owners

    // ---- per-CTA tables ----
    for (int i = tid; i < 31 * R1; i += SW_THREADS) {
        const int l = i / R1 + 1, q = i - (l - 1) * R1;
        sm.tw[i] = __ldg(pw.twiddle + ((l * q) & (N - 1)));
    }
    for (int i = tid; i < SW_CPL / 2 * 32; i += SW_THREADS) {
        const int c = i >> 5, l = (i & 31) % LPS;        // lanes of different symbols share the table of their place
        const uint32_t two = __ldg(reinterpret_cast<const uint32_t *>(p.bin_of_src) + l * (SW_CPL / 2) + c);
        const uint32_t b0 = two & 0xffffu, b1 = two >> 16;
        sm.bin_t[i] = (b0 < N / 2 ? b0 - 1 : b0 - (N - K)) | ((b1 < N / 2 ? b1 - 1 : b1 - (N - K)) << 16);
    }
    for (int b = tid; b < 256; b += SW_THREADS) {
        uint32_t s = 0;
#pragma unroll
        for (int n = 0; n < 8; n++) s |= ((b >> (7 - n)) & 1u) << (4 * n);
        sm.spread[b] = s;
    }
    for (int i = tid; i < 6 * 32; i += SW_THREADS) {
        const int w = i >> 5, l = (i & 31) % LPS;
        uint32_t v = 0;
#pragma unroll
        for (int n = 0; n < 8; n++) v |= (uint32_t)__ldg(p.phase0 + SW_CPL * l + 8 * w + n) << (4 * n);
        sm.ph0[i] = v;
    }
    if (tid < 16) {
        // the ideal 8-PSK points (see k_symbols_w: an approximation of the reference's float32 product chain
        // within the parity tolerance)
        const float v = 0.70710678118654752440f;
        const float c[8] = {1.f, v, 0.f, -v, -1.f, -v, 0.f, v};
        sm.c8[tid] = tid < 8 ? make_float2(c[tid], c[(tid + 6) & 7]) : make_float2(0.f, 0.f);
    }
    __syncthreads();

    float2 *xb = sm.x[warp];
    uint8_t *code = sm.code[warp];
    unsigned clip = 0;
    const int L = p.L, NG = pw.n_groups;
    // every warp of the grid takes one contiguous range of the batch's n_tf * NG symbol groups
    const long long n_grp = (long long)pw.n_tf * NG;
    const long long n_warps = (long long)gridDim.x * SW_WARPS;
    const int per_warp = (int)((n_grp + n_warps - 1) / n_warps);
    const long long g0 = ((long long)blockIdx.x * SW_WARPS + warp) * per_warp;
    const long long g1 = g0 + per_warp < n_grp ? g0 + per_warp : n_grp;

    // the lane's share of a bit row: bytes 6 ls .. 6 ls + 5 of the I half and of the Q half (K / 8 bytes each)
    auto fetch_row = [&](const uint8_t *row) {
        const uintptr_t ai = reinterpret_cast<uintptr_t>(row + 6 * ls);
        const uint32_t *wi = reinterpret_cast<const uint32_t *>(ai & ~(uintptr_t)3);
        const uint32_t *wq = wi + K / 32;
        RowRaw r;
        r.i0 = __ldg(wi); r.i1 = __ldg(wi + 1);
        r.q0 = __ldg(wq); r.q1 = __ldg(wq + 1);
        return r;
    };
    auto unpack_row = [&](const RowRaw &r) {
        const unsigned sh = (unsigned)((6 * ls) & 2) * 8;
        RowBits b;
        b.i_lo = __funnelshift_r(r.i0, r.i1, sh);
        b.i_hi = (r.i1 >> sh) & 0xffffu;
        b.q_lo = __funnelshift_r(r.q0, r.q1, sh);
        b.q_hi = (r.q1 >> sh) & 0xffffu;
        return b;
    };

    // base[w]: phase of the symbol BEFORE the lane group's first symbol of the current group, for the lane's carriers
    uint32_t base[6] = {0, 0, 0, 0, 0, 0};
    RowRaw nextrow = {0, 0, 0, 0};
    if (g0 < n_grp) {
        const int tf = (int)(g0 / NG);
        const int gi = (int)(g0 - (long long)tf * NG);
        const int s0 = 1 + gi * G;                           // first symbol of the group
        const uint8_t *bits = p.bits + (size_t)tf * p.tf_in_bytes;
#pragma unroll
        for (int w = 0; w < 6; w++) base[w] = sm.ph0[w * 32 + lane];
        // phase of symbol s0 - 1 = reference + the rows of the symbols 2 .. s0 - 1 (symbol s carries row s - 2),
        // summed bit-sliced as in k_symbols_w
        const int nd = max(0, s0 - 2);
        uint32_t c0l = 0, c0h = 0, c1l = 0, c1h = 0, pql = 0, pqh = 0;
        const uint8_t *row = bits;
#pragma unroll 8
        for (int d = 0; d < nd; d++, row += K / 4) {
            const RowBits b = unpack_row(fetch_row(row));
            const uint32_t xl = b.i_lo ^ b.q_lo, xh = b.i_hi ^ b.q_hi;
            c1l ^= c0l & xl; c1h ^= c0h & xh;
            c0l ^= xl; c0h ^= xh;
            pql ^= b.q_lo; pqh ^= b.q_hi;
        }
        const uint32_t nb = (uint32_t)(nd & 7) * 0x11111111u;
        const uint32_t m2l = c1l ^ pql, m2h = c1h ^ pqh;
#pragma unroll
        for (int w = 0; w < 6; w++) {
            const uint32_t b0 = ((w < 4 ? c0l >> (8 * w) : c0h >> (8 * (w - 4)))) & 0xffu;
            const uint32_t b1 = ((w < 4 ? m2l >> (8 * w) : m2h >> (8 * (w - 4)))) & 0xffu;
            const uint32_t t = (nb + 2u * sm.spread[b0] + 4u * sm.spread[b1]) & 0x77777777u;
            base[w] = (base[w] + t) & 0x77777777u;
        }
        const int s = s0 + gs;
        if (s >= 2 && s <= L) nextrow = fetch_row(bits + (size_t)(s - 2) * (K / 4));
    }

    // in step within a group of warps that share their sub-partitions, the groups apart (see k_symbols_w)
#ifndef SWG_GROUPS
#define SWG_GROUPS (-1)
#endif
    constexpr int SW_GROUPS = sw_n_groups(SWG_GROUPS);
    const int grp = sw_group_of(warp, SWG_GROUPS);
    for (int it = 0; it < per_warp; it++) {
        sw_bar_sync(1 + grp, SW_THREADS / SW_GROUPS);        // the warps walk through the loop body in step (instruction cache)
        if (SW_GROUPS > 1 && it == 0 && grp > 0) __nanosleep(18000u / SW_GROUPS * grp);
        const long long g = g0 + it;
        const bool live_group = g < g1;
        const int tf = live_group ? (int)(g / NG) : 0;
        const int gi = live_group ? (int)(g - (long long)tf * NG) : 0;
        const int s0 = 1 + gi * G;
        const uint8_t *bits = p.bits + (size_t)tf * p.tf_in_bytes;
        const size_t out_base = (size_t)tf * p.tf_samples;
        if (!live_group) continue;
        if (gi == 0) {
            // start of a TF: the null symbol (no TII) is all zeros whatever gain it borrows; the differential chain
            // restarts from the phase reference
            for (int i = lane; i < p.null_size; i += 32)
                store_sample<POST>(p.out, out_base + i, make_float2(0.f, 0.f), p.post, clip);
#pragma unroll
            for (int w = 0; w < 6; w++) base[w] = sm.ph0[w * 32 + lane];
        }
        // ---- 1. differential phases of the group's symbols ----
        const int s = s0 + gs;                               // the lane's own symbol (may lie beyond L in the last group)
        uint32_t ph[6];
        {
            const RowBits b = unpack_row(nextrow);
            // the row of the lane's symbol in the NEXT group, one iteration ahead
            if (g + 1 < g1) {
                const long long gn = g + 1;
                const int tfn = (int)(gn / NG);
                const int sn = 1 + (int)(gn - (long long)tfn * NG) * G + gs;
                if (sn >= 2 && sn <= L)
                    nextrow = fetch_row(p.bits + (size_t)tfn * p.tf_in_bytes + (size_t)(sn - 2) * (K / 4));
            }
            const bool has_row = s >= 2 && s <= L;
#pragma unroll
            for (int w = 0; w < 6; w++) {
                const uint32_t ib = ((w < 4 ? b.i_lo >> (8 * w) : b.i_hi >> (8 * (w - 4)))) & 0xffu;
                const uint32_t qb = ((w < 4 ? b.q_lo >> (8 * w) : b.q_hi >> (8 * (w - 4)))) & 0xffu;
                uint32_t inc = has_row ? phase_step(sm.spread, ib, qb) & 0x77777777u : 0u;
                // inclusive scan over the G lanes that own these carriers (lane stride LPS)
#pragma unroll
                for (int d = LPS; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc = (inc + o) & 0x77777777u;
                }
                ph[w] = (base[w] + inc) & 0x77777777u;
                // the last symbol of the group is the base of the next one
                base[w] = __shfl_sync(0xffffffffu, ph[w], (G - 1) * LPS + ls);
            }
        }
        // ---- scatter the codes of the lane's 48 carriers ----
        {
            uint32_t bins[SW_CPL / 2];
#pragma unroll
            for (int i = 0; i < SW_CPL / 2; i++) bins[i] = sm.bin_t[i * 32 + lane];
            uint8_t *cs = code + gs * K;
#pragma unroll
            for (int i = 0; i < SW_CPL; i++) {
                const uint32_t c = (ph[i >> 3] >> (4 * (i & 7))) & 7u;
                const uint32_t bin = (i & 1) ? bins[i >> 1] >> 16 : bins[i >> 1] & 0xffffu;
                cs[bin] = (uint8_t)c;
            }
        }
        __syncwarp();

        // ---- 2. pass 1: for every symbol of the group the bins lane + 32 r, r < R1 (register transform) ----
        // occupied: bins 1 .. K/2 (r < 3 R1 / 8, and r == 3 R1 / 8 for lane 0) and N - K/2 .. N - 1 (r >= 5 R1 / 8)
#pragma unroll
        for (int gg = 0; gg < G; gg++) {
            float2 v[R1];
            const uint8_t *cs = code + gg * K;
#pragma unroll
            for (int r = 0; r < R1; r++) {
                if (r > 3 * R1 / 8 && r < 5 * R1 / 8) {
                    v[r] = make_float2(0.f, 0.f);
                }
                else {
                    uint32_t c = cs[r == 0 ? max(lane - 1, 0) : r <= 3 * R1 / 8 ? lane + 32 * r - 1 : lane + 32 * (r - R1 / 4)];
                    if (r == 0 && lane == 0) c = 8;
                    if (r == 3 * R1 / 8 && lane != 0) c = 8;
                    v[r] = sm.c8[c];
                }
            }
            wg_fft_r1<R1>(v);
            if (gg == 0) {
                // the previous group's bulk copies must have read the staging area by now
                if (!POST) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                __syncwarp();
            }
#pragma unroll
            for (int r = 0; r < R1; r++) xb[lane * SW_XPAD + gg * R1 + r] = v[r];
        }
        __syncwarp();

        // ---- pass 2: butterflies bi = lane and lane + 32 (radix 32): symbol bi / R1, samples q + R1 p, q = bi % R1 ----
        float2 y[2][32];
        {
            // R1 <= 32: both butterflies of a lane have the same q = lane % R1, hence the same twiddles: one table
            // read serves both
            const int q = lane % R1;
#pragma unroll
            for (int l = 0; l < 32; l++) {
                y[0][l] = xb[l * SW_XPAD + lane];
                y[1][l] = xb[l * SW_XPAD + lane + 32];
            }
#pragma unroll
            for (int l = 1; l < 32; l++) {
                const float2 w = sm.tw[(l - 1) * R1 + q];
                y[0][l] = cmul(y[0][l], w);
                y[1][l] = cmul(y[1][l], w);
            }
            fft32<true>(y[0]);
            fft32<true>(y[1]);
        }
        __syncwarp();                        // exchange buffer read by every lane: free for the staging

        // ---- 3. gain per symbol (GainControl.cpp:196-340): statistics over the N samples = R1 lanes x 32 ----
        float g_sym[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float2 (&u)[32] = y[h];
            float gv;
            if (p.gain_mode == 0) {
                gv = 512.0f;
            }
            else if (p.gain_mode == 1) {
                float mn = u[0].x, mx = u[0].x;
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    mn = fminf(mn, fminf(u[i].x, u[i].y));
                    mx = fmaxf(mx, fmaxf(u[i].x, u[i].y));
                }
#pragma unroll
                for (int o = R1 / 2; o > 0; o >>= 1) {
                    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                const float m = fmaxf(-mn, mx);
                gv = ((int)m != 0) ? 32767.0f / m : 1.0f;
            }
            else {
                float2 sum = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 32; i++) sum = cadd(sum, u[i]);
#pragma unroll
                for (int o = R1 / 2; o > 0; o >>= 1) {
                    sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
                    sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
                }
                const float2 mean = make_float2(sum.x * (1.0f / N), sum.y * (1.0f / N));
                float2 var = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const float2 d = csub(u[i], mean);
                    var = __ffma2_rn(d, d, var);
                }
                float vr = var.x, vi = var.y;
#pragma unroll
                for (int o = R1 / 2; o > 0; o >>= 1) {
                    vr += __shfl_xor_sync(0xffffffffu, vr, o);
                    vi += __shfl_xor_sync(0xffffffffu, vi, o);
                }
                const float sdr = p.var_factor * sqrtf(vr * (1.0f / N));
                const float sdi = p.var_factor * sqrtf(vi * (1.0f / N));
                gv = ((int)sdr != 0) ? 32767.0f / fmaxf(sdr, sdi) : 1.0f;       // NULL detection: real part only
            }
            g_sym[h] = gv * p.gain_const;
        }

        // ---- 4. guard interval + store (GuardIntervalInserter.cpp:301-319) ----
        const int pre = p.sym_size - N;
        if (POST) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int bi = lane + 32 * h, sg = bi / R1, q = bi % R1;
                const int ss = s0 + sg;
                if (ss > L) continue;
                const size_t pos = out_base + sym_pos(p, ss);
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const int n = q + R1 * i;
                    const float2 o = cscale(y[h][i], g_sym[h]);
                    store_sample<POST>(p.out, pos + pre + n, o, p.post, clip);
                    if (n >= N - pre) store_sample<POST>(p.out, pos + n - (N - pre), o, p.post, clip);
                }
            }
        }
        else {
            // the scaled symbols in natural order, symbol sg at [sg * N, sg * N + N)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int bi = lane + 32 * h, sg = bi / R1, q = bi % R1;
#pragma unroll
                for (int i = 0; i < 32; i++) xb[sg * N + q + R1 * i] = cscale(y[h][i], g_sym[h]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            // One lane issues every copy: the bulk-copy instruction takes uniform operands, issued from several
            // lanes it becomes a serialising loop over them.  (Each copy costs ~0.01 ms per launch in issue latency,
            // profiles/r02_ncu_symbols_wg.txt: 2 G copies per group are why TM II trails TM IV trails TM I.)
            if (lane == 0) {
                float2 *gout = reinterpret_cast<float2 *>(p.out) + out_base + sym_pos(p, s0);
#pragma unroll
                for (int sg = 0; sg < G; sg++) {
                    if (s0 + sg <= L) {
                        sw_bulk_store(gout + (size_t)sg * p.sym_size + pre, xb + sg * N, N * (int)sizeof(float2));
                        sw_bulk_store(gout + (size_t)sg * p.sym_size, xb + sg * N + (N - pre), pre * (int)sizeof(float2));
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (!POST) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    if (POST && p.post.format != 0) flush_clip(p.post, clip);
}

} // namespace dabmod
