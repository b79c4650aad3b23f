// k_symbols_w: the TM I (N = 2048) symbol kernel, one WARP per OFDM symbol.
//
// Same chain as k_symbols (kernels.cuh) for the plain configuration -- no CicEq,
// no CFR, no windowing, no TII frame -- i.e. reference QpskSymbolMapper.cpp:105-156,
// FrequencyInterleaver.cpp:103-126, DifferentialModulator.cpp:45-76,
// SignalMultiplexer.cpp:45-71 (NullSymbol), OfdmGenerator.cpp:157-308,
// GainControl.cpp:82-340, GuardIntervalInserter.cpp:301-319.
//
// Why a second kernel: k_symbols is bound by the shared-memory/LSU data pipe
// (128 B/clk/SM): a 2048-point transform done as 16 x 16 x 8 by 128 threads crosses
// shared memory twice and synchronises the CTA six times per symbol.  Here a warp
// owns the whole symbol: every lane holds 64 points in registers, the transform is
// 64 x 32 (fft_reg.cuh), so the data crosses shared memory ONCE, and the only
// synchronisation is __syncwarp.  Per symbol the LSU moves ~0.9 k wavefronts
// instead of ~1.9 k, which is what the HBM write stream (910 clk/symbol/SM) needs.
//
// Work item of a warp = (TF, chunk of consecutive symbols), as in k_symbols; the
// differential phases at the chunk start are obtained by bit-sliced counting over
// the preceding bit rows (2-bit counter of i^q and parity of q per carrier).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft_reg.cuh"
#include "kernels.cuh"

namespace dabmod {

constexpr int SW_WARPS = 12;                // warps (= symbols in flight) per CTA
constexpr int SW_THREADS = SW_WARPS * 32;
// Grouping of the twelve warps (see the kernel): -1 = three groups ACROSS the sub-partitions (every sub-partition hosts
// one warp of each group), 1 = all warps in step, 2 / 4 = groups on DISJOINT sub-partitions.
#ifndef SW_GROUPS_PLAIN
#define SW_GROUPS_PLAIN (-1)
#endif
#ifndef SW_GROUPS_FUSE
#define SW_GROUPS_FUSE (-1)
#endif
__device__ __forceinline__ constexpr int sw_n_groups(int mode) { return mode < 0 ? 3 : mode; }
// a warp's scheduler (SM sub-partition) is warp % 4
__device__ __forceinline__ int sw_group_of(int warp, int mode)
{
    return mode < 0 ? (warp >> 2) : mode == 4 ? (warp & 3) : mode == 2 ? ((warp >> 1) & 1) : 0;
}
constexpr int SW_XPAD = 65;                 // lane stride (complex) of the exchange buffer
constexpr int SW_N = 2048, SW_K = 1536;
constexpr int SW_CPL = SW_K / 32;           // 48 source carriers per lane

constexpr int SW_CODES = SW_K;              // one phase code per occupied bin (the 512 guard-band bins are skipped)

struct SymWSmem {
    float2 tw[31 * 32];                     // second pass twiddles: tw[(r-1)*32 + j] = e^{+j 2 pi j r / 2048}, j < 32
                                            // (j + 32: times the constants W64^r, fft_reg.cuh mul_w64_ramp)
    uint32_t bin_t[SW_CPL / 2 * 32];        // code index of the lane's source carriers, two per word: [i/2][lane]
                                            // holds carriers 48*lane + i (low half) and i + 1 (high half);
                                            // code index = bin - 1 (bins 1..768), bin - 512 (bins 1280..2047)
    uint32_t spread[256];
    uint32_t ph0[6 * 32];                   // phase reference of the lane's carriers, nibble packed: [word][lane]
    float2 c8[16];                          // value of phase code 0..7 (units of pi/4); code 8 = empty bin
    float2 x[SW_WARPS][32 * SW_XPAD];       // per-warp exchange buffer between the two FFT passes, then the
                                            // staging area the symbol is bulk-copied to HBM from
    uint8_t code[SW_WARPS][SW_CODES];       // per-warp phase codes of the symbol being assembled
};
static_assert(sizeof(SymWSmem) <= 227 * 1024, "k_symbols_w shared memory");

struct SymWParams {
    SymParams s;                            // shared with k_symbols
    const float2 *twiddle_w;                // 31 * 64 entries
    int n_tf;
    float2 taps[45];                        // FUSE: the FIR taps as (tap, tap) pairs
    const float2 *tii_tail;                 // FUSE with TII: the last 44 samples of the stream's TII null symbol (else nullptr)
    int compact;                            // 1: write only the N samples of every data symbol, back to back
                                            // ([tf][s-1][N], no null symbol, no cyclic prefix): the layout
                                            // k_fir_sym reads (it rebuilds the guard interval itself)
};

// I/Q bit bytes of the lane's 48 carriers in one bit row: 6 bytes each, as (4 bytes, 2 bytes)
struct RowBits { uint32_t i_lo, i_hi, q_lo, q_hi; };

// The lane's 6 bytes of a bit row start at byte 6*lane (2-byte aligned): two aligned 32-bit
// loads cover them.  Loading (RowRaw) and unpacking (RowBits) are separate so that a row
// can be fetched one symbol ahead without anything waiting on the load.
struct RowRaw { uint32_t i0, i1, q0, q1; };

__device__ __forceinline__ RowRaw sw_fetch_row(const uint8_t *row, int lane)
{
    const uintptr_t ai = reinterpret_cast<uintptr_t>(row + 6 * lane);
    const uint32_t *wi = reinterpret_cast<const uint32_t *>(ai & ~(uintptr_t)3);
    const uint32_t *wq = wi + SW_K / 32;           // the Q half starts K/8 = 192 bytes later
    RowRaw r;
    r.i0 = __ldg(wi); r.i1 = __ldg(wi + 1);
    r.q0 = __ldg(wq); r.q1 = __ldg(wq + 1);
    return r;
}

__device__ __forceinline__ RowBits sw_unpack_row(const RowRaw &r, int lane)
{
    const unsigned sh = (unsigned)((6 * lane) & 2) * 8;    // rows are 4-byte aligned
    RowBits b;
    b.i_lo = __funnelshift_r(r.i0, r.i1, sh);
    b.i_hi = (r.i1 >> sh) & 0xffffu;
    b.q_lo = __funnelshift_r(r.q0, r.q1, sh);
    b.q_hi = (r.q1 >> sh) & 0xffffu;
    return b;
}

__device__ __forceinline__ RowBits sw_load_row(const uint8_t *row, int lane)
{
    return sw_unpack_row(sw_fetch_row(row, lane), lane);
}

__device__ __forceinline__ void sw_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sw_bar_arrive(int id, int nthreads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// shared -> global bulk async copy (cp.async.bulk, SASS UBLKCP); addresses and size multiples of 16
__device__ __forceinline__ void sw_bulk_store(void *gdst, const void *ssrc, int bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}

template <bool POST, bool FUSE = false>
__global__ void __launch_bounds__(SW_THREADS, 1) k_symbols_w(const __grid_constant__ SymWParams pw)
{
    const SymParams &p = pw.s;
    constexpr int N = SW_N, K = SW_K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SymWSmem &sm = *reinterpret_cast<SymWSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- per-CTA tables ----
    for (int i = tid; i < 31 * 32; i += SW_THREADS) {
        const int r1 = i >> 5, j = i & 31;
        sm.tw[i] = __ldg(pw.twiddle_w + r1 * 64 + j);
    }
    for (int i = tid; i < K / 2; i += SW_THREADS) {
        const int l = i / (SW_CPL / 2), c = i - l * (SW_CPL / 2);
        const uint32_t two = __ldg(reinterpret_cast<const uint32_t *>(p.bin_of_src) + i);
        const uint32_t b0 = two & 0xffffu, b1 = two >> 16;
        sm.bin_t[c * 32 + l] = (b0 < 1024 ? b0 - 1 : b0 - 512) | ((b1 < 1024 ? b1 - 1 : b1 - 512) << 16);
    }
    for (int b = tid; b < 256; b += SW_THREADS) {
        uint32_t s = 0;
#pragma unroll
        for (int n = 0; n < 8; n++) s |= ((b >> (7 - n)) & 1u) << (4 * n);
        sm.spread[b] = s;
    }
    for (int i = tid; i < 6 * 32; i += SW_THREADS) {
        const int w = i >> 5, l = i & 31;
        uint32_t v = 0;
#pragma unroll
        for (int n = 0; n < 8; n++) v |= (uint32_t)__ldg(p.phase0 + SW_CPL * l + 8 * w + n) << (4 * n);
        sm.ph0[i] = v;
    }
    if (tid < 16) {
        // The ideal 8-PSK points {1, v, 0, -v, -1}, v = (float)M_SQRT1_2.  An approximation of the reference,
        // not a restatement: its std::complex<float> product chain drifts (fl(v*v) = 0.49999997, about 3e-8 per
        // symbol, 2e-6 relative at worst over 75 symbols); the parity tolerance (2e-6 relative RMS) covers it.
        // (The fixed-point chain IS closed: 11585^2 rounds to 8192, symbols_fixed.cuh.)
        const float v = 0.70710678118654752440f;
        const float c[8] = {1.f, v, 0.f, -v, -1.f, -v, 0.f, v};
        sm.c8[tid] = tid < 8 ? make_float2(c[tid], c[(tid + 6) & 7]) : make_float2(0.f, 0.f);
    }
    __syncthreads();

    float2 *xb = sm.x[warp];
    uint8_t *code = sm.code[warp];
    // The loop body is ~50 KB of straight-line code, far beyond the 32 KB instruction cache: warps at different
    // places in it each stream it from L2 (measured: 1.03 ms free-running, 0.62 ms with all twelve in step on one named
    // barrier per symbol).  In step, however, everybody scatters and exchanges at the same time (LSU busy, FP32 idle)
    // and then everybody runs butterflies.  Groups of warps that are in step among themselves and apart from each other
    // overlap the phases at the price of one instruction stream per group; WHERE the groups sit matters (a warp's
    // scheduler is warp % 4), measured per 1024 TFs, compact layout:
    //                                                                    plain     with the FIR inside (FUSE)
    //   one group                                                       0.456 ms   0.996 ms
    //   two groups, warps 0-5 / 6-11 (both streams on every sub-partition, round 2 first attempt)  slower than one
    //   two groups on disjoint sub-partitions (bit 1 of the warp index)  0.417 ms   0.897 ms
    //   four groups, one per sub-partition                              0.408 ms   0.882 ms
    //   three groups ACROSS the sub-partitions (warp / 4: every sub-partition hosts one warp of each group, so its
    //   warps are in three different phases)                             0.383 ms   0.876-0.884 ms   <- both kernels
    // (profiles/r02_sw_groups_after_modifiers.txt.  Before the complex arithmetic moved to the packed instructions'
    // operand modifiers -- 15 % more instructions -- the fused kernel preferred two disjoint groups: 0.923 against
    // 0.945 ms, profiles/r02_sw_groups.txt.)
    constexpr int GMODE = FUSE ? SW_GROUPS_FUSE : SW_GROUPS_PLAIN;
    constexpr int SW_GROUPS = sw_n_groups(GMODE);
    constexpr int GRP_THREADS = SW_THREADS / SW_GROUPS;
    const int grp = sw_group_of(warp, GMODE);
    bool first_iter = true;

    // Work split: the batch is one sequence of n_tf * L transformed symbols (symbol s = 1..L of
    // every TF; the all-zero null symbol s = 0 is written by whoever owns s = 1).  Every warp
    // of the grid takes one contiguous range of it, so there is one phase prefix per warp
    // and the ranges differ by at most one symbol.
    unsigned clip = 0;
    const int L = p.L;
    const long long n_sym = (long long)pw.n_tf * L;
    const long long n_warps = (long long)gridDim.x * SW_WARPS;
    const int per_warp = (int)((n_sym + n_warps - 1) / n_warps);
    const long long g0 = ((long long)blockIdx.x * SW_WARPS + warp) * per_warp;
    const long long g1 = g0 + per_warp < n_sym ? g0 + per_warp : n_sym;

    uint32_t ph[6] = {0, 0, 0, 0, 0, 0};
    RowRaw nextrow = {0, 0, 0, 0};
    if (g0 < n_sym) {
        const int tf = (int)(g0 / L);
        const int s_first = 1 + (int)(g0 - (long long)tf * L);
        const uint8_t *bits = p.bits + (size_t)tf * p.tf_in_bytes;
        // ---- running phase of the lane's 48 source carriers, 8 nibbles per word ----
#pragma unroll
        for (int w = 0; w < 6; w++) ph[w] = sm.ph0[w * 32 + lane];
        // Phase prefix: symbol s >= 2 carries data row d = s - 2.  Rows before the range are
        // summed bit-sliced: increment = 1 + 2 (i ^ q) + 4 q (units of pi/4), so
        // sum = nd + 2 (c0 + 2 c1) + 4 pq with (c1 c0) a 2-bit counter of i^q, pq the parity of q.
        const int nd = max(0, s_first - 2);
        uint32_t c0l = 0, c0h = 0, c1l = 0, c1h = 0, pql = 0, pqh = 0;
        const uint8_t *row = bits;
#pragma unroll 16
        for (int d = 0; d < nd; d++, row += K / 4) {
            const RowBits b = sw_load_row(row, lane);
            const uint32_t xl = b.i_lo ^ b.q_lo, xh = b.i_hi ^ b.q_hi;
            c1l ^= c0l & xl; c1h ^= c0h & xh;
            c0l ^= xl; c0h ^= xh;
            pql ^= b.q_lo; pqh ^= b.q_hi;
        }
        const uint32_t base = (uint32_t)(nd & 7) * 0x11111111u;
        const uint32_t m2l = c1l ^ pql, m2h = c1h ^ pqh;
#pragma unroll
        for (int w = 0; w < 6; w++) {
            const uint32_t b0 = ((w < 4 ? c0l >> (8 * w) : c0h >> (8 * (w - 4)))) & 0xffu;
            const uint32_t b1 = ((w < 4 ? m2l >> (8 * w) : m2h >> (8 * (w - 4)))) & 0xffu;
            const uint32_t t = (base + 2u * sm.spread[b0] + 4u * sm.spread[b1]) & 0x77777777u;
            ph[w] = (ph[w] + t) & 0x77777777u;
        }
        // bit row of the first data symbol of the range; afterwards always one symbol ahead
        if (s_first >= 2) nextrow = sw_fetch_row(bits + (size_t)(s_first - 2) * (K / 4), lane);
    }
    // FUSE: the 45-tap FIR (FIRFilter.cpp:168-191: out[n] = sum_j taps[j] in[n + j]) runs on the staged symbol before
    // it leaves.  The last 44 outputs of a symbol's body reach into the next symbol's cyclic prefix: they are
    // computed one iteration later, from the 44 samples carried in registers and the prefix head of the symbol then
    // staged.  A warp whose range ends inside a transmission frame therefore assembles ONE symbol more (the first of
    // its neighbour's range, "ghost": nothing of it is stored) to finish its own last symbol.
    float2 carry0 = make_float2(0.f, 0.f), carry1 = carry0;     // lane < 22: samples 2004 + 2 lane (+ 1) of the previous symbol
    bool have_carry = false;
    const bool need_ghost = FUSE && g1 > g0 && g1 < n_sym && (g1 % L) != 0;
    {
        for (int it = 0; it < per_warp + (FUSE ? 1 : 0); it++) {
            sw_bar_sync(1 + grp, GRP_THREADS);
#ifndef SW_STAGGER_PLAIN_NS
#define SW_STAGGER_PLAIN_NS 6000u
#endif
            if (SW_GROUPS > 1 && first_iter && grp > 0) __nanosleep((FUSE ? 5000u : SW_STAGGER_PLAIN_NS) * grp);   // start the groups apart
            const long long g = g0 + it;
            const bool ghost = FUSE && need_ghost && g == g1;
            const bool fft_symbol = g < g1 || ghost;
            const int tf = fft_symbol ? (int)(g / L) : 0;
            const int s = 1 + (int)(g - (long long)tf * L);
            const uint8_t *bits = p.bits + (size_t)tf * p.tf_in_bytes;
            const size_t out_base = (size_t)tf * p.tf_samples;
            if (fft_symbol && s == 1) {
                // start of a TF.  Null symbol without TII: all-zero carriers -> all-zero samples,
                // whatever gain it borrows from symbol 1 (GainControl.cpp:139-144).  The
                // differential chain restarts from the phase reference (DifferentialModulator.cpp:65).
                // FUSE with TII: every second frame of the stream (TII.cpp:225-242) starts with the TII symbol, one constant
                // vector: k_tii_fill writes its filtered samples, and its last 44 samples are what symbol 1's prefix follows
                const bool tii_frame = FUSE && pw.tii_tail != nullptr && ((p.tf_offset + (unsigned long long)tf) & 1ull) == 0;
                if (!pw.compact && !tii_frame) {
                    // (FUSE: the last 44 samples of the filtered null symbol see the first prefix, see below)
                    for (int i = lane; i < p.null_size - (FUSE ? 44 : 0); i += 32)
                        store_sample<POST>(p.out, out_base + i, make_float2(0.f, 0.f), p.post, clip);
                }
                if (FUSE) {
                    carry0 = carry1 = make_float2(0.f, 0.f);        // the null symbol ends in zeros
                    if (tii_frame && lane < 22) {
                        carry0 = __ldg(pw.tii_tail + 2 * lane);
                        carry1 = __ldg(pw.tii_tail + 2 * lane + 1);
                    }
                    have_carry = true;
                }
#pragma unroll
                for (int w = 0; w < 6; w++) ph[w] = sm.ph0[w * 32 + lane];
            }
            if (fft_symbol) {
                // ---- 1. differential phase of this symbol, scattered by FFT bin as byte codes ----
                const RowBits b = sw_unpack_row(nextrow, lane);
                if (s + 1 <= L && (g + 1 < g1 || (need_ghost && g + 1 == g1)))
                    nextrow = sw_fetch_row(bits + (size_t)(s - 1) * (K / 4), lane);
                if (s >= 2) {
#pragma unroll
                    for (int w = 0; w < 6; w++) {
                        const uint32_t ib = ((w < 4 ? b.i_lo >> (8 * w) : b.i_hi >> (8 * (w - 4)))) & 0xffu;
                        const uint32_t qb = ((w < 4 ? b.q_lo >> (8 * w) : b.q_hi >> (8 * (w - 4)))) & 0xffu;
                        ph[w] = (ph[w] + phase_step(sm.spread, ib, qb)) & 0x77777777u;
                    }
                }
                // (all table loads first: the compiler cannot tell that the byte stores below
                // never hit the table, and would otherwise order every load behind a store)
                uint32_t bins[SW_CPL / 2];
#pragma unroll
                for (int i = 0; i < SW_CPL / 2; i++) bins[i] = sm.bin_t[i * 32 + lane];
#pragma unroll
                for (int i = 0; i < SW_CPL; i++) {
                    const uint32_t c = (ph[i >> 3] >> (4 * (i & 7))) & 7u;
                    const uint32_t bin = (i & 1) ? bins[i >> 1] >> 16 : bins[i >> 1] & 0xffffu;
                    code[bin] = (uint8_t)c;
                }
                __syncwarp();

                // ---- 2. inverse FFT, pass 1: lane owns bins lane + 32 r, r < 64 (radix 64) ----
                // Bins 769..1279 and bin 0 are empty (OfdmGenerator.cpp:207-220): r in 25..39 for
                // every lane, r = 24 except lane 0, r = 0 for lane 0.
                float2 v[64];
#pragma unroll
                for (int r = 0; r < 64; r++) {
                    if (r >= 25 && r <= 39) {
                        v[r] = make_float2(0.f, 0.f);
                    }
                    else {
                        uint32_t c = code[r == 0 ? max(lane - 1, 0) : r < 25 ? lane + 32 * r - 1 : lane + 32 * (r - 16)];
                        if (r == 0 && lane == 0) c = 8;
                        if (r == 24 && lane != 0) c = 8;
                        v[r] = sm.c8[c];
                    }
                }
                fft64<true>(v);
                // the previous symbol's bulk copies must have read the staging area by now
                if (!POST) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                }
#pragma unroll
                for (int r = 0; r < 64; r++) xb[lane * SW_XPAD + r] = v[r];
                __syncwarp();
            }
            first_iter = false;
            if (fft_symbol) {
                // ---- pass 2: butterflies j = lane and lane + 32 (radix 32), sample n = j + 64 r ----
                // y[i] = sample lane + 32 i
                float2 y[64];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int j = lane + 32 * h;
                    float2 u[32];
#pragma unroll
                    for (int r = 0; r < 32; r++) u[r] = xb[r * SW_XPAD + j];
#pragma unroll
                    for (int r = 1; r < 32; r++) u[r] = cmul(u[r], sm.tw[(r - 1) * 32 + lane]);
                    if (h == 1) mul_w64_ramp<true>(u);
                    fft32<true>(u);
#pragma unroll
                    for (int r = 0; r < 32; r++) y[2 * r + h] = u[r];
                }
                __syncwarp();                        // exchange buffer read by every lane: free for the staging

                // ---- 3. gain (GainControl.cpp:196-340), statistics over the N samples ----
                float g_sym;
                if (p.gain_mode == 0) {
                    g_sym = 512.0f;
                }
                else if (p.gain_mode == 1) {
                    float mn = y[0].x, mx = y[0].x;
#pragma unroll
                    for (int i = 0; i < 64; i++) {
                        mn = fminf(mn, fminf(y[i].x, y[i].y));
                        mx = fmaxf(mx, fmaxf(y[i].x, y[i].y));
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    }
                    const float m = fmaxf(-mn, mx);
                    g_sym = ((int)m != 0) ? 32767.0f / m : 1.0f;
                }
                else {
                    // two-pass mean / variance of re and im separately (packed: .x = re, .y = im)
                    float2 sum = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 64; i++) sum = cadd(sum, y[i]);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
                        sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
                    }
                    const float2 mean = make_float2(sum.x * (1.0f / N), sum.y * (1.0f / N));
                    float2 var = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 64; i++) {
                        const float2 d = csub(y[i], mean);
                        var = __ffma2_rn(d, d, var);
                    }
                    float vr = var.x, vi = var.y;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        vr += __shfl_xor_sync(0xffffffffu, vr, o);
                        vi += __shfl_xor_sync(0xffffffffu, vi, o);
                    }
                    const float sdr = p.var_factor * sqrtf(vr * (1.0f / N));
                    const float sdi = p.var_factor * sqrtf(vi * (1.0f / N));
                    // NULL detection looks at the real part only (GainControl.cpp:331)
                    g_sym = ((int)sdr != 0) ? 32767.0f / fmaxf(sdr, sdi) : 1.0f;
                }
                g_sym *= p.gain_const;

                // ---- 4. guard interval + store (GuardIntervalInserter.cpp:301-319) ----
                const int pre = p.sym_size - N;
                const size_t pos = out_base + sym_pos(p, s);
                if (POST) {
#pragma unroll
                    for (int i = 0; i < 64; i++) {
                        const int n = lane + 32 * i;
                        const float2 o = cscale(y[i], g_sym);
                        store_sample<POST>(p.out, pos + pre + n, o, p.post, clip);
                        if (n >= N - pre) store_sample<POST>(p.out, pos + n - (N - pre), o, p.post, clip);
                    }
                }
                else {
                    // complexf output: the scaled symbol goes to shared memory in natural order and
                    // two bulk async copies (TMA engine) land it in HBM, body and cyclic prefix;
                    // the LSU sees 64 conflict-free shared stores instead of 79 global ones and
                    // nobody waits for the memory system.
#pragma unroll
                    for (int i = 0; i < 64; i++) xb[lane + 32 * i] = cscale(y[i], g_sym);
                    if (FUSE) {
                        __syncwarp();
                        float2 *const outp = reinterpret_cast<float2 *>(p.out);
                        // The phase codes are dead until the next symbol: their buffer holds the seam windows.
                        //   strip A = prefix end ++ body start = samples 2004..2047 ++ 0..43 of this symbol
                        //             -> the 44 outputs across the prefix/body seam, prefix positions 460..503
                        //   strip B = the previous symbol's samples 2004..2047 (carried; zeros after the null symbol)
                        //             ++ this symbol's prefix head = its samples 1544..1587
                        //             -> the last 44 body outputs of the previous symbol, resp. the end of the null symbol
                        float2 *const strip = reinterpret_cast<float2 *>(code);
                        if (!ghost) {
                            for (int i = lane; i < 88; i += 32) strip[i] = i < 44 ? xb[2004 + i] : xb[i - 44];
                        }
                        if (have_carry) {
                            if (lane < 22) { strip[88 + 2 * lane] = carry0; strip[88 + 2 * lane + 1] = carry1; }
                            for (int i = lane; i < 44; i += 32) strip[132 + i] = xb[1544 + i];
                        }
                        const bool tf_end = s == L && !ghost;
                        const int cl = lane < 22 ? lane : 21;
                        const float2 nc0 = xb[2004 + 2 * cl], nc1 = xb[2005 + 2 * cl];
                        __syncwarp();
                        // The body outputs whose window stays inside the symbol, 0..2003, are filtered in place: a lane
                        // owns 17 consecutive outputs per pass (odd lane stride: conflict free); everybody reads before
                        // anybody writes, and a pass only writes its own range.  118 of the 128 lane slots carry body
                        // outputs; slots 118..123 (lanes 22..27 of the last pass) take the two strips, three lanes each,
                        // and store their results themselves -- one code path, different base pointers (like k_fir_tma).
                        // At the end of a TF a fifth pass turns strip A into "samples 2004..2047 ++ zeros": the symbol's
                        // own last 44 outputs, whose window runs into zeros (FIRFilter.cpp:186-191).
#pragma unroll 1
                        for (int pass = ghost ? 3 : 0; pass < (tf_end ? 5 : 4); pass++) {
                            const bool extra = pass == 4;
                            if (extra) {
                                __syncwarp();
                                for (int i = lane; i < 44; i += 32) strip[44 + i] = make_float2(0.f, 0.f);
                                __syncwarp();
                            }
                            const int slot = 32 * (extra ? 3 : pass) + lane;
                            const int sl = slot - 118;                  // >= 0: a strip lane
                            const int which = sl / 3, k0 = sl >= 0 ? 17 * (sl - 3 * which) : 17 * slot;
                            const bool on = sl < 0 ? (!ghost && !extra)
                                                   : extra ? which == 0 : which == 0 ? !ghost : which == 1 ? have_carry : false;
                            float2 acc[17];
#pragma unroll
                            for (int m = 0; m < 17; m++) acc[m] = make_float2(0.f, 0.f);
                            if (on) {
                                const float2 *x = (sl < 0 ? xb : strip + 88 * which) + k0;
#pragma unroll
                                for (int i = 0; i < 17 + 45 - 1; i++) {
                                    const float2 v = x[i];
#pragma unroll
                                    for (int m = 0; m < 17; m++) {
                                        const int j = i - m;
                                        if (j >= 0 && j < 45) acc[m] = __ffma2_rn(v, pw.taps[j], acc[m]);
                                    }
                                }
                            }
                            __syncwarp();
                            if (on) {
                                float2 *dst;
                                int lim;
                                if (sl < 0) { dst = xb + k0; lim = 2004 - k0; }
                                else {
                                    lim = 44 - k0;
                                    if (extra) dst = outp + pos + pre + 2004 + k0;                          // end of the TF
                                    else if (which == 0) dst = outp + pos + (pre - 44) + k0;                // prefix/body seam
                                    else if (s == 1) dst = outp + out_base + (p.null_size - 44) + k0;      // end of the null symbol
                                    else dst = outp + (pos - p.sym_size) + pre + 2004 + k0;                 // the previous symbol's tail
                                }
#pragma unroll
                                for (int m = 0; m < 17; m++)
                                    if (m < lim) dst[m] = acc[m];
                            }
                        }
                        carry0 = nc0; carry1 = nc1;
                        have_carry = s != L;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && !ghost) {
                        if (FUSE) {
                            // body outputs 0..2003, and the prefix: its outputs 0..459 ARE the body outputs 1544..2003
                            // (same operands in the same order)
                            float2 *gout = reinterpret_cast<float2 *>(p.out) + pos;
                            sw_bulk_store(gout + pre, xb, 2004 * (int)sizeof(float2));
                            sw_bulk_store(gout, xb + 1544, (pre - 44) * (int)sizeof(float2));
                        }
                        else if (pw.compact) {
                            float2 *gout = reinterpret_cast<float2 *>(p.out) + ((size_t)tf * L + (s - 1)) * N;
                            sw_bulk_store(gout, xb, N * (int)sizeof(float2));
                        }
                        else {
                            float2 *gout = reinterpret_cast<float2 *>(p.out) + pos;
                            sw_bulk_store(gout + pre, xb, N * (int)sizeof(float2));
                            sw_bulk_store(gout, xb + (N - pre), pre * (int)sizeof(float2));
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
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
