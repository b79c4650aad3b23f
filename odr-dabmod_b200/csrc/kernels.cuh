// sm_100a kernels of the COFDM hot path.  See DESIGN.md for the data layout
// and the roofline each kernel is measured against.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft.cuh"

namespace dabmod {

constexpr int SYM_THREADS = 128;     // one symbol group = 2048 complex points = 16 per thread
constexpr int SYM_POINTS = 2048;     // G * N for every transmission mode
constexpr int SYM_BUF = SYM_POINTS + SYM_POINTS / 16;  // padded (spad)
constexpr int MAX_TII = 64;          // 32 carrier pairs
constexpr int MAX_FIR_TAPS = 128;
constexpr int MAX_WINDOW = 1024;     // 2 * windowOverlap entries kept on chip (W <= 512)

// ---------------------------------------------------------------------------
// Epilogue: [MemlessPoly] -> [FormatConverter] -> store.
// Reference: MemlessPoly.cpp:237-309, FormatConverter.cpp:112-165.
// ---------------------------------------------------------------------------
struct PostParams {
    int dpd_mode;        // 0 none, 1 odd polynomial, 2 LUT
    int format;          // 0 complexf, 1 s16, 2 u8, 3 s8
    float am[5];
    float pm[5];
    float lut_scale;
    const float *lut;    // 32 floats (device)
    unsigned long long *clipped; // device counter
};

__device__ __forceinline__ float2 dpd_apply(const PostParams &pp, float2 x)
{
    if (pp.dpd_mode == 1) {
        // Horner steps as fused multiply-adds (the reference is built with GCC's default
        // -ffp-contract=fast on an FMA machine, so neither rounding is canonical)
        const float mag = fmaf(x.x, x.x, x.y * x.y);
        float amp = pp.am[4];
        float ph = pp.pm[4];
#pragma unroll
        for (int i = 3; i >= 0; i--) {
            amp = fmaf(mag, amp, pp.am[i]);
            ph = fmaf(mag, ph, pp.pm[i]);
        }
        ph = -ph;
        const float p2 = ph * ph;
        const float re = fmaf(-p2, fmaf(p2, fmaf(p2, -0.00138888f, 0.486666f), -0.5f), 1.0f);
        const float im = ph * fmaf(p2, fmaf(p2, 0.00833333f, 0.166666f), 1.0f);
        const float ar = x.x * amp, ai = x.y * amp;
        return make_float2(fmaf(ar, re, -ai * im), fmaf(ar, im, ai * re));
    }
    if (pp.dpd_mode == 2) {
        const float mag = hypotf(x.x, x.y);
        const unsigned scaled = (unsigned)__float2ll_rn(__fmul_rn(mag, pp.lut_scale));
        const float g = __ldg(pp.lut + (scaled >> 27));
        return make_float2(x.x * g, x.y * g);
    }
    return x;
}

// The odd-polynomial predistorter on two samples at once: the Horner chains run on packed FP32 (FFMA2: same rounding
// as two fmaf, half the issue slots).  Other modes fall back to dpd_apply.
__device__ __forceinline__ void dpd_apply2(const PostParams &pp, float2 &a, float2 &b)
{
    if (pp.dpd_mode != 1) {
        a = dpd_apply(pp, a);
        b = dpd_apply(pp, b);
        return;
    }
    const float2 mag = make_float2(fmaf(a.x, a.x, a.y * a.y), fmaf(b.x, b.x, b.y * b.y));
    float2 amp = make_float2(pp.am[4], pp.am[4]), ph = make_float2(pp.pm[4], pp.pm[4]);
#pragma unroll
    for (int i = 3; i >= 0; i--) {
        amp = __ffma2_rn(mag, amp, make_float2(pp.am[i], pp.am[i]));
        ph = __ffma2_rn(mag, ph, make_float2(pp.pm[i], pp.pm[i]));
    }
    ph = make_float2(-ph.x, -ph.y);
    const float2 p2 = __fmul2_rn(ph, ph);
    const float2 np2 = make_float2(-p2.x, -p2.y);
    float2 re = __ffma2_rn(p2, make_float2(-0.00138888f, -0.00138888f), make_float2(0.486666f, 0.486666f));
    re = __ffma2_rn(p2, re, make_float2(-0.5f, -0.5f));
    re = __ffma2_rn(np2, re, make_float2(1.0f, 1.0f));
    float2 im = __ffma2_rn(p2, make_float2(0.00833333f, 0.00833333f), make_float2(0.166666f, 0.166666f));
    im = __ffma2_rn(p2, im, make_float2(1.0f, 1.0f));
    im = __fmul2_rn(ph, im);
    const float ar = a.x * amp.x, ai = a.y * amp.x, br = b.x * amp.y, bi = b.y * amp.y;
    a = make_float2(fmaf(ar, re.x, -ai * im.x), fmaf(ar, im.x, ai * re.x));
    b = make_float2(fmaf(br, re.y, -bi * im.y), fmaf(br, im.y, bi * re.y));
}

// saturating conversion with C truncation, counts clipped components
// (branch free: float -> int conversion saturates, the range test only feeds the counter)
__device__ __forceinline__ int fmt_s16(float v, unsigned &clip)
{
    clip += (v < -32768.0f) | (v > 32767.0f);
    return max(-32768, min(32767, __float2int_rz(v)));
}
__device__ __forceinline__ int fmt_u8(float v, unsigned &clip)
{
    const float s = v + 128.0f;
    clip += (s < 0.0f) | (s > 255.0f);
    return max(0, min(255, __float2int_rz(s)));
}
__device__ __forceinline__ int fmt_s8(float v, unsigned &clip)
{
    clip += (v < -128.0f) | (v > 127.0f);
    return max(-128, min(127, __float2int_rz(v)));
}

// Stores complex sample number `idx` of the output stream.
template <bool POST>
__device__ __forceinline__ void store_sample(void *out, size_t idx, float2 v, const PostParams &pp,
                                             unsigned &clip)
{
    if (!POST) {
        reinterpret_cast<float2 *>(out)[idx] = v;
        return;
    }
    v = dpd_apply(pp, v);
    if (pp.format == 0) {
        reinterpret_cast<float2 *>(out)[idx] = v;
    }
    else if (pp.format == 1) {
        short2 s;
        s.x = (short)fmt_s16(v.x, clip);
        s.y = (short)fmt_s16(v.y, clip);
        reinterpret_cast<short2 *>(out)[idx] = s;
    }
    else if (pp.format == 2) {
        uchar2 s;
        s.x = (unsigned char)fmt_u8(v.x, clip);
        s.y = (unsigned char)fmt_u8(v.y, clip);
        reinterpret_cast<uchar2 *>(out)[idx] = s;
    }
    else {
        char2 s;
        s.x = (signed char)fmt_s8(v.x, clip);
        s.y = (signed char)fmt_s8(v.y, clip);
        reinterpret_cast<char2 *>(out)[idx] = s;
    }
}

// Two consecutive samples starting at an even `idx`: one store of twice the width.
template <bool POST>
__device__ __forceinline__ void store_run2(void *out, size_t idx, float2 a, float2 b, const PostParams &pp, unsigned &clip)
{
    if (POST) dpd_apply2(pp, a, b);
    if (!POST || pp.format == 0) {
        *reinterpret_cast<float4 *>(reinterpret_cast<float2 *>(out) + idx) = make_float4(a.x, a.y, b.x, b.y);
    }
    else if (pp.format == 1) {
        const unsigned w0 = (unsigned)(fmt_s16(a.x, clip) & 0xffff) | ((unsigned)(fmt_s16(a.y, clip) & 0xffff) << 16);
        const unsigned w1 = (unsigned)(fmt_s16(b.x, clip) & 0xffff) | ((unsigned)(fmt_s16(b.y, clip) & 0xffff) << 16);
        *reinterpret_cast<uint2 *>(reinterpret_cast<short2 *>(out) + idx) = make_uint2(w0, w1);
    }
    else {
        const int ax = pp.format == 2 ? fmt_u8(a.x, clip) : fmt_s8(a.x, clip), ay = pp.format == 2 ? fmt_u8(a.y, clip) : fmt_s8(a.y, clip);
        const int bx = pp.format == 2 ? fmt_u8(b.x, clip) : fmt_s8(b.x, clip), by = pp.format == 2 ? fmt_u8(b.y, clip) : fmt_s8(b.y, clip);
        *reinterpret_cast<unsigned *>(reinterpret_cast<uchar2 *>(out) + idx) =
            (unsigned)(ax & 0xff) | ((unsigned)(ay & 0xff) << 8) | ((unsigned)(bx & 0xff) << 16) | ((unsigned)(by & 0xff) << 24);
    }
}

// Four consecutive samples starting at `idx` (idx a multiple of 4): same arithmetic as store_sample, one or two
// 16-byte stores (one 8-byte store for the one-byte formats) instead of four narrow ones.
template <bool POST>
__device__ __forceinline__ void store_run4(void *out, size_t idx, float2 a, float2 b, float2 c, float2 d,
                                           const PostParams &pp, unsigned &clip)
{
    if (POST) { dpd_apply2(pp, a, b); dpd_apply2(pp, c, d); }
    if (!POST || pp.format == 0) {
        float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float2 *>(out) + idx);
        o[0] = make_float4(a.x, a.y, b.x, b.y);
        o[1] = make_float4(c.x, c.y, d.x, d.y);
    }
    else if (pp.format == 1) {
        const float2 v[4] = {a, b, c, d};
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
            w[i] = (unsigned)(fmt_s16(v[i].x, clip) & 0xffff) | ((unsigned)(fmt_s16(v[i].y, clip) & 0xffff) << 16);
        *reinterpret_cast<uint4 *>(reinterpret_cast<short2 *>(out) + idx) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    else {
        const float2 v[4] = {a, b, c, d};
        unsigned w[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int x = pp.format == 2 ? fmt_u8(v[i].x, clip) : fmt_s8(v[i].x, clip);
            const int y = pp.format == 2 ? fmt_u8(v[i].y, clip) : fmt_s8(v[i].y, clip);
            w[i >> 1] |= ((unsigned)(x & 0xff) | ((unsigned)(y & 0xff) << 8)) << (16 * (i & 1));
        }
        *reinterpret_cast<uint2 *>(reinterpret_cast<uchar2 *>(out) + idx) = make_uint2(w[0], w[1]);
    }
}

__device__ __forceinline__ void flush_clip(const PostParams &pp, unsigned clip)
{
    // one atomic per warp that saw clipping
    const unsigned total = __reduce_add_sync(0xffffffffu, clip);
    if (total && (threadIdx.x & 31) == 0) atomicAdd(pp.clipped, (unsigned long long)total);
}

// ---------------------------------------------------------------------------
// k_symbols: bits -> QPSK -> frequency interleave -> differential modulation
//   -> null/TII multiplex -> [CicEq] -> carrier placement -> IFFT -> [CFR]
//   -> gain -> guard interval [-> window] -> store
// Reference: QpskSymbolMapper.cpp:105-156, FrequencyInterleaver.cpp:103-126,
// DifferentialModulator.cpp:45-76, SignalMultiplexer.cpp:45-71, TII.cpp:172-245,
// CicEqualizer.cpp:66-91, OfdmGenerator.cpp:157-373, GainControl.cpp:82-340,
// GuardIntervalInserter.cpp:115-323.
//
// One CTA = one (TF, chunk of consecutive symbol groups).  A symbol group is
// G = 2048/N consecutive OFDM symbols transformed side by side in one 2048
// point shared buffer, so every mode keeps all 128 threads busy with 16
// points each.
// ---------------------------------------------------------------------------
// Per-symbol statistics of the CFR block (OfdmGenerator.cpp:228-275): what the reference feeds to its
// PAPRStats / clip ratio / MER read-outs ("papr", "clip_stats" remote-control parameters).
struct CfrSymStat {
    float peak_before, sum_before;   // max and sum of |x|^2 over the N samples after the first IFFT
    float peak_after, sum_after;     // the same after the CFR iteration
    float sum_ref, sum_delta;        // sum |X_ref|^2 and sum |X_out - X_ref|^2 over the bins (MER, by Parseval)
    unsigned clip, errclip;          // clipped samples / clipped error bins
};

struct SymParams {
    // mode
    int L, K, N, null_size, sym_size, tf_in_bytes, tf_samples;
    int G;                  // symbols per group = 2048 / N
    int n_groups;           // ceil((L+1)/G)
    int groups_per_chunk;
    int n_chunks;
    // per-source-carrier tables (device)
    const uint16_t *bin_of_src;   // K: FFT bin of the carrier that source j is interleaved to
    const uint8_t *phase0;        // K: phase reference of that carrier, units of pi/4 (even)
    const float *cic;             // K or nullptr: CicEqualizer gain of that carrier
    const float2 *twiddle;        // per-pass tables of the mode's IFFT (tables.h: symbol_fft_twiddles)
    int n_twiddle;
    // null symbol / TII
    int tii_count;                // carriers set in the TII symbol (0 = plain null symbol)
    int tii_parity;               // TII is inserted on TFs where ((tf + tii_parity) & 1) == 0
    const uint16_t *tii_bin;      // tii_count FFT bins
    const float2 *tii_val;        // tii_count values (phase reference, CicEq applied)
    // CFR
    int cfr;
    float cfr_clip, cfr_errclip;
    CfrSymStat *cfr_stats;        // (L+1) records per TF of this launch, or nullptr
    // gain
    int gain_mode;
    float gain_const;             // normalise * digital_gain
    float var_factor;
    // guard interval windowing
    int window;                   // windowOverlap W
    const float *window_tab;      // 2W floats
    // I/O
    const uint8_t *bits;          // n_tf * tf_in_bytes
    void *out;                    // n_tf * tf_samples samples
    unsigned long long tf_offset; // index of the first TF of this launch within the stream
    PostParams post;
};

struct SymSmem {
    float2 buf[SYM_BUF];
    float2 tw[SYM_POINTS];
    uint32_t spread[256];
    float2 c8[8];
    float red[4][8][4];    // [warp][symbol in group][re, im, re2, im2] / [min,max]
    float gain[8];
};

// extra shared memory of the OPT variant (CFR and/or OFDM windowing)
struct SymSmemOpt {
    float2 ref[SYM_BUF];            // frequency-domain symbols as fed to the IFFT (CFR reference)
    float2 tail[2][MAX_WINDOW];     // windowed falling edge of the previous symbol, double buffered
    float win[MAX_WINDOW];          // rising edge, 2W entries
    float stat[3][4][8][4];         // CFR statistics: [phase][warp][symbol in group][field] partial results
};

// Reduce per-thread partials of the G symbols of a group over the warp and park them in shared memory
// (every lane of a warp looks at the same symbol for a given element index, see the CFR loops).
template <int G>
__device__ __forceinline__ void cfr_stat_warp(float (*dst)[8][4], int tid, const float (&mx)[G], const float (&sum1)[G],
                                              const float (&sum2)[G], const unsigned (&cnt)[G])
{
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int g = 0; g < G; g++) {
        const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(mx[g]));   // non-negative floats order like uints
        const unsigned c = __reduce_add_sync(0xffffffffu, cnt[g]);
        float a = sum1[g], b = sum2[g];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
            dst[warp][g][0] = __uint_as_float(m);
            dst[warp][g][1] = a;
            dst[warp][g][2] = b;
            dst[warp][g][3] = __uint_as_float(c);
        }
    }
}

// out-position of symbol s inside the TF
__device__ __forceinline__ int sym_pos(const SymParams &p, int s)
{
    return s == 0 ? 0 : p.null_size + (s - 1) * p.sym_size;
}

// D nibbles (phase increments in units of pi/4) of 8 carriers from their I and Q bytes
__device__ __forceinline__ uint32_t phase_step(const uint32_t *spread, unsigned ib, unsigned qb)
{
    return 0x11111111u + 2u * spread[ib ^ qb] + 4u * spread[qb];
}

// One in-place Stockham pass over the 2048-point buffer (G transforms of size N).
template <int R, bool INV, int PER, int N>
__device__ __forceinline__ void sym_pass(float2 *buf, int tid, int Ns, const float2 *twp)
{
    StockhamPass<R, INV, PER> ps;
    ps.load(buf, tid, SYM_THREADS, N, Ns, twp);
    __syncthreads();
    ps.store(buf, tid, SYM_THREADS, N, Ns);
    __syncthreads();
}

// All passes of the mode's transform, in place, natural order in and out.
// Radix plans: N=2048: 16,16,8; 1024: 16,8,8; 512: 8,8,8; 256: 16,16 (tables.h: symbol_fft_twiddles).
template <int N, bool INV>
__device__ __forceinline__ void sym_fft(float2 *buf, const float2 *tw, int tid)
{
    if (N == 2048) {
        sym_pass<16, INV, 1, N>(buf, tid, 1, tw);
        sym_pass<16, INV, 1, N>(buf, tid, 16, tw);
        sym_pass<8, INV, 2, N>(buf, tid, 256, tw + 15 * 16);
    }
    else if (N == 1024) {
        sym_pass<16, INV, 1, N>(buf, tid, 1, tw);
        sym_pass<8, INV, 2, N>(buf, tid, 16, tw);
        sym_pass<8, INV, 2, N>(buf, tid, 128, tw + 7 * 16);
    }
    else if (N == 512) {
        sym_pass<8, INV, 2, N>(buf, tid, 1, tw);
        sym_pass<8, INV, 2, N>(buf, tid, 8, tw);
        sym_pass<8, INV, 2, N>(buf, tid, 64, tw + 7 * 8);
    }
    else {
        sym_pass<16, INV, 1, N>(buf, tid, 1, tw);
        sym_pass<16, INV, 1, N>(buf, tid, 16, tw);
    }
}

// OPT = true adds the optional features of the same reference blocks: crest factor
// reduction (OfdmGenerator.cpp:310-373) and OFDM windowing (GuardIntervalInserter.cpp:149-300).
// They cost shared memory and registers, so the plain configurations get their own instantiation.
template <int N, bool POST, bool OPT>
__global__ void __launch_bounds__(SYM_THREADS, OPT ? 3 : 5) k_symbols(const __grid_constant__ SymParams p)
{
    constexpr int G = SYM_POINTS / N;         // symbols per group
    constexpr int TG = SYM_THREADS / G;       // threads per symbol in the emit phase
    constexpr int K16 = (N * 3 / 4) / 16;     // 16-carrier work items per symbol (K = 3N/4)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SymSmem &sm = *reinterpret_cast<SymSmem *>(smem_raw);
    SymSmemOpt &so = *reinterpret_cast<SymSmemOpt *>(smem_raw + sizeof(SymSmem));   // OPT only

    const int tid = threadIdx.x;
    const int tf = blockIdx.x / p.n_chunks;
    const int chunk = blockIdx.x - tf * p.n_chunks;
    const int grp0 = chunk * p.groups_per_chunk;
    const int grp1 = min(grp0 + p.groups_per_chunk, p.n_groups);
    constexpr int K = N * 3 / 4;              // carriers of the mode (p.K), a compile-time fact here: divisions by constants
    const uint8_t *bits = p.bits + (size_t)tf * p.tf_in_bytes;
    const size_t out_base = (size_t)tf * p.tf_samples;
    const bool tii_on = p.tii_count > 0 && (((p.tf_offset + tf + p.tii_parity) & 1) == 0);
    const int W = OPT ? p.window : 0;
    const bool cfr = OPT && p.cfr != 0;

    // ---- per-CTA tables ----
    for (int i = tid; i < p.n_twiddle; i += SYM_THREADS) sm.tw[i] = __ldg(p.twiddle + i);
    for (int b = tid; b < 256; b += SYM_THREADS) {
        uint32_t s = 0;
#pragma unroll
        for (int n = 0; n < 8; n++) s |= ((b >> (7 - n)) & 1u) << (4 * n);
        sm.spread[b] = s;
    }
    if (tid < 8) {
        // The ideal 8-PSK points {1, v, 0, -v, -1}, v = (float)M_SQRT1_2.  An approximation of the reference,
        // not a restatement: its std::complex<float> product chain drifts (fl(v*v) = 0.49999997, about 3e-8 per
        // symbol, 2e-6 relative at worst over 75 symbols); the parity tolerance (2e-6 relative RMS) covers it.
        // (The fixed-point chain IS closed: 11585^2 rounds to 8192, symbols_fixed.cuh.)
        const float v = 0.70710678118654752440f;
        const float c[8] = {1.f, v, 0.f, -v, -1.f, -v, 0.f, v};
        sm.c8[tid] = make_float2(c[tid], c[(tid + 6) & 7]);
    }
    if (OPT) {
        for (int i = tid; i < 2 * W; i += SYM_THREADS) {
            so.win[i] = __ldg(p.window_tab + i);
            so.tail[0][i] = make_float2(0.f, 0.f);
            so.tail[1][i] = make_float2(0.f, 0.f);
        }
    }

    // ---- iteration schedule over symbol groups ----
    //   plain:  the chunk's groups in order, except that TM I (G == 1) runs group 1 before
    //           group 0: the null symbol takes the gain of symbol 1 (GainControl.cpp:139-144)
    //   window: groups in natural order (each symbol needs the falling edge of its
    //           predecessor), preceded by one extra iteration: the group before the chunk
    //           (tail only), or for TM I chunk 0 group 1 (gain only).
    enum { EMIT = 0, TAIL_ONLY = 1, GAIN_ONLY = 2 };
    const bool win_pre = W > 0 && (grp0 > 0 || G == 1);
    const int n_iter = (grp1 - grp0) + (win_pre ? 1 : 0);
    const int g_first = (W > 0 && grp0 > 0) ? grp0 - 1 : grp0;   // first group whose data bits are consumed

    // ---- carrier work item of this thread: 16 consecutive source carriers ----
    // thread (g, jj): symbol g of the group, carriers 16*jj .. 16*jj+15
    const int cg = tid / K16, jj = tid - cg * K16;
    const bool carrier_thread = tid < G * K16;
    uint32_t ph_lo = 0, ph_hi = 0;      // nibble-packed running phase, carriers 0-7 / 8-15
    if (carrier_thread) {
#pragma unroll
        for (int n = 0; n < 8; n++) {
            ph_lo |= (uint32_t)__ldg(p.phase0 + 16 * jj + n) << (4 * n);
            ph_hi |= (uint32_t)__ldg(p.phase0 + 16 * jj + 8 + n) << (4 * n);
        }
    }
    __syncthreads();

    // Phase prefix: consume the data symbols that precede the first group.
    // Symbol s >= 2 carries data symbol d = s - 2.
    {
        const int nd = max(0, g_first * G - 2);
        if (carrier_thread) {
            const uint8_t *row = bits + 2 * jj;
            for (int d = 0; d < nd; d++, row += K / 4) {
                const unsigned iw = __ldg(reinterpret_cast<const unsigned short *>(row));
                const unsigned qw = __ldg(reinterpret_cast<const unsigned short *>(row + K / 8));
                ph_lo = (ph_lo + phase_step(sm.spread, iw & 0xff, qw & 0xff)) & 0x77777777u;
                ph_hi = (ph_hi + phase_step(sm.spread, iw >> 8, qw >> 8)) & 0x77777777u;
            }
        }
    }

    unsigned clip = 0;
    float gain_sym1 = 1.0f;
    int tail_par = 0;
    // group handled by iteration `it` of the schedule above
    auto group_of = [&](int it) {
        if (W > 0) {
            if (grp0 > 0) return grp0 - 1 + it;
            if (G == 1) return it == 0 ? 1 : it - 1;
            return it;
        }
        if (G == 1 && grp0 == 0) return it == 0 ? 1 : it == 1 ? 0 : it;   // neither consumes data bits
        return grp0 + it;
    };
    // The bit rows of a group's symbols (I bits | Q bits << 16 of this thread's 16 carriers) are fetched one
    // iteration ahead, and the carriers' FFT bins once per CTA: the carrier threads were spending 40 % of the
    // kernel's stall samples behind these dependent loads (profiles, per-instruction samples).
    auto fetch_rows = [&](int grp, unsigned short (&ri)[G], unsigned short (&rq)[G]) {
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int s = grp * G + g;
            ri[g] = 0; rq[g] = 0;
            if (s >= 2 && s <= p.L) {               // (the values are not touched before the next iteration)
                const uint8_t *row = bits + (size_t)(s - 2) * (K / 4) + 2 * jj;
                ri[g] = __ldg(reinterpret_cast<const unsigned short *>(row));
                rq[g] = __ldg(reinterpret_cast<const unsigned short *>(row + K / 8));
            }
        }
    };
    unsigned short rows_i[G], rows_q[G];
    uint32_t binp[8];
#pragma unroll
    for (int g = 0; g < G; g++) { rows_i[g] = 0; rows_q[g] = 0; }
#pragma unroll
    for (int n = 0; n < 8; n++) binp[n] = 0;
    if (carrier_thread) {
        fetch_rows(group_of(0), rows_i, rows_q);
#pragma unroll
        for (int n = 0; n < 8; n++) binp[n] = __ldg(reinterpret_cast<const uint32_t *>(p.bin_of_src + 16 * jj) + n);
    }
    for (int it = 0; it < n_iter; it++) {
        const int grp = group_of(it);
        int what = EMIT;
        if (W > 0) {
            if (grp0 > 0) what = it == 0 ? TAIL_ONLY : EMIT;
            else if (G == 1) what = it == 0 ? GAIN_ONLY : EMIT;
        }
        unsigned cur_i[G], cur_q[G];
#pragma unroll
        for (int g = 0; g < G; g++) { cur_i[g] = rows_i[g]; cur_q[g] = rows_q[g]; }
        if (carrier_thread && it + 1 < n_iter) fetch_rows(group_of(it + 1), rows_i, rows_q);
        if (!OPT && G == 1 && grp == 0 && !tii_on) {
            // plain null symbol: all-zero carriers -> all-zero samples
            const size_t pos = out_base;
            for (int i = tid; i < p.null_size; i += SYM_THREADS)
                store_sample<POST>(p.out, pos + i, make_float2(0.f, 0.f), p.post, clip);
            continue;
        }
        const int s0 = grp * G;   // first symbol of the group
        const bool stats_on = OPT && cfr && what == EMIT && p.cfr_stats != nullptr;
        // ---- 1. frequency-domain symbols into the shared buffer ----
        // zero bins: DC and the guard band (OfdmGenerator.cpp:207-220)
        {
            const int nz = N - K;            // bins K/2+1 .. N-K/2-1 plus bin 0
            for (int i = tid; i < G * nz; i += SYM_THREADS) {
                const int g = i / nz, z = i - g * nz;
                const int bin = z == 0 ? 0 : K / 2 + z;
                sm.buf[spad(g * N + bin)] = make_float2(0.f, 0.f);
            }
        }
        if (carrier_thread) {
            // advance the phase through the symbols of the group; emit our own
            uint32_t my_lo = 0, my_hi = 0;
            int my_kind = 0; // 0 = null (zeros/TII), 1 = phase (ref or data)
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int s = s0 + g;
                if (s >= 2 && s <= p.L) {
                    const unsigned iw = cur_i[g], qw = cur_q[g];
                    ph_lo = (ph_lo + phase_step(sm.spread, iw & 0xff, qw & 0xff)) & 0x77777777u;
                    ph_hi = (ph_hi + phase_step(sm.spread, iw >> 8, qw >> 8)) & 0x77777777u;
                }
                if (g == cg) {
                    my_lo = ph_lo; my_hi = ph_hi;
                    my_kind = (s >= 1 && s <= p.L) ? 1 : 0;
                }
            }
            float2 *dst = sm.buf + 0;
            const int fbase = cg * N;
#pragma unroll
            for (int n = 0; n < 16; n++) {
                const int j = 16 * jj + n;
                const unsigned ph = ((n < 8 ? my_lo >> (4 * n) : my_hi >> (4 * (n - 8)))) & 7u;
                float2 v = my_kind ? sm.c8[ph] : make_float2(0.f, 0.f);
                if (p.cic) {
                    const float f = __ldg(p.cic + j);
                    v.x *= f; v.y *= f;
                }
                dst[spad(fbase + (int)((n & 1) ? binp[n >> 1] >> 16 : binp[n >> 1] & 0xffffu))] = v;
            }
        }
        __syncthreads();
        // TII carriers overwrite the zeroed null symbol (symbol 0 is in group 0, slot 0)
        if (grp == 0 && tii_on && tid < p.tii_count)
            sm.buf[spad(__ldg(p.tii_bin + tid))] = __ldg(p.tii_val + tid);
        if (grp == 0 && tii_on) __syncthreads();

        // ---- 2. inverse FFT of the G symbols, in place ----
        // thread (eg, tt): symbol eg of the group, samples tt + TG*i
        const int eg = tid / TG, tt = tid - eg * TG;
        float2 x[16];
        if (OPT) {
            if (cfr) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int e = spad(tid + SYM_THREADS * i);
                    so.ref[e] = sm.buf[e];
                }
            }
            sym_fft<N, true>(sm.buf, sm.tw, tid);
            if (cfr) {
                // ---- 2'. one CFR iteration (OfdmGenerator.cpp:310-373) ----
                // Element tid + 128 i belongs to symbol i / (16 / G) of the group for every thread, so the
                // statistics are G compile-time indexed partials per thread.
                const float clip_sq = p.cfr_clip * p.cfr_clip;
                const float err_sq = p.cfr_errclip * p.cfr_errclip;
                float st_mx[G], st_a[G], st_b[G];
                unsigned st_n[G];
#pragma unroll
                for (int g = 0; g < G; g++) { st_mx[g] = 0.f; st_a[g] = 0.f; st_b[g] = 0.f; st_n[g] = 0; }
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int e = spad(tid + SYM_THREADS * i);
                    float2 v = sm.buf[e];
                    const float mag = v.x * v.x + v.y * v.y;
                    st_mx[i / (16 / G)] = fmaxf(st_mx[i / (16 / G)], mag);
                    st_a[i / (16 / G)] += mag;
                    if (mag > clip_sq) {
                        const float f = sqrtf(clip_sq / mag);
                        v.x *= f; v.y *= f;
                        sm.buf[e] = v;
                        st_n[i / (16 / G)]++;
                    }
                }
                if (stats_on) cfr_stat_warp<G>(so.stat[0], tid, st_mx, st_a, st_b, st_n);
                __syncthreads();
                if (stats_on && tid < G && s0 + tid <= p.L) {
                    CfrSymStat &o = p.cfr_stats[(size_t)tf * (p.L + 1) + s0 + tid];
                    float mx = 0.f, a = 0.f;
                    unsigned c = 0;
                    for (int w = 0; w < 4; w++) {
                        mx = fmaxf(mx, so.stat[0][w][tid][0]);
                        a += so.stat[0][w][tid][1];
                        c += __float_as_uint(so.stat[0][w][tid][3]);
                    }
                    o.peak_before = mx; o.sum_before = a; o.clip = c;
                }
                sym_fft<N, false>(sm.buf, sm.tw, tid);
#pragma unroll
                for (int g = 0; g < G; g++) { st_a[g] = 0.f; st_b[g] = 0.f; st_n[g] = 0; }
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int e = spad(tid + SYM_THREADS * i);
                    const float2 f = sm.buf[e];
                    const float2 pt = make_float2(f.x * (1.0f / N), f.y * (1.0f / N));
                    const float2 r = so.ref[e];
                    float2 err = make_float2(r.x - pt.x, r.y - pt.y);
                    const float mag = err.x * err.x + err.y * err.y;
                    st_a[i / (16 / G)] = fmaf(r.x, r.x, fmaf(r.y, r.y, st_a[i / (16 / G)]));
                    if (mag > err_sq) {
                        const float s = sqrtf(err_sq / mag);
                        err.x *= s; err.y *= s;
                        // the bin moves by (1 - s) |error| away from the reference: the MER numerator
                        st_b[i / (16 / G)] = fmaf(mag * (1.0f - s), 1.0f - s, st_b[i / (16 / G)]);
                        st_n[i / (16 / G)]++;
                    }
                    sm.buf[e] = make_float2(pt.x + err.x, pt.y + err.y);
                }
                if (stats_on) cfr_stat_warp<G>(so.stat[1], tid, st_mx, st_a, st_b, st_n);
                __syncthreads();
                if (stats_on && tid < G && s0 + tid <= p.L) {
                    CfrSymStat &o = p.cfr_stats[(size_t)tf * (p.L + 1) + s0 + tid];
                    float a = 0.f, b = 0.f;
                    unsigned c = 0;
                    for (int w = 0; w < 4; w++) {
                        a += so.stat[1][w][tid][1];
                        b += so.stat[1][w][tid][2];
                        c += __float_as_uint(so.stat[1][w][tid][3]);
                    }
                    o.sum_ref = a; o.sum_delta = b; o.errclip = c;
                }
                sym_fft<N, true>(sm.buf, sm.tw, tid);
            }
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = sm.buf[spad(eg * N + tt + TG * i)];
            if (stats_on) {
                // PAPR after CFR: peak and mean power of the symbol this thread emits (TG threads per symbol)
                constexpr int SL = TG < 32 ? TG : 32;
                float mx = 0.f, a = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float mag = x[i].x * x[i].x + x[i].y * x[i].y;
                    mx = fmaxf(mx, mag);
                    a += mag;
                }
#pragma unroll
                for (int o = SL / 2; o > 0; o >>= 1) {
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                }
                const int lane = tid & 31;
                if ((lane & (SL - 1)) == 0) {
                    so.stat[2][tid >> 5][lane / SL][0] = mx;
                    so.stat[2][tid >> 5][lane / SL][1] = a;
                }
                __syncthreads();
                if (tid < G && s0 + tid <= p.L) {
                    constexpr int WPS_ = TG < 32 ? 1 : TG / 32, SPW_ = TG < 32 ? 32 / TG : 1;
                    const int ww = TG < 32 ? tid / SPW_ : tid * WPS_, ss = TG < 32 ? tid % SPW_ : 0;
                    float m2 = 0.f, a2 = 0.f;
                    for (int w = 0; w < WPS_; w++) {
                        m2 = fmaxf(m2, so.stat[2][ww + w][ss][0]);
                        a2 += so.stat[2][ww + w][ss][1];
                    }
                    CfrSymStat &o = p.cfr_stats[(size_t)tf * (p.L + 1) + s0 + tid];
                    o.peak_after = m2; o.sum_after = a2;
                }
            }
        }
        else if (N == 2048) {
            // Per-pass twiddle tables (see StockhamPass): pass 2 at sm.tw, pass 3 behind it.
            sym_pass<16, true, 1, N>(sm.buf, tid, 1, sm.tw);
            sym_pass<16, true, 1, N>(sm.buf, tid, 16, sm.tw);
            // last pass stays in registers: butterfly p of this thread yields samples
            // tid + 128*p + 256*r, i.e. x[2r + p] of the emit layout (TG = 128)
            StockhamPass<8, true, 2> ps;
            ps.load(sm.buf, tid, SYM_THREADS, N, 256, sm.tw + 15 * 16);
#pragma unroll
            for (int r = 0; r < 8; r++) { x[2 * r] = ps.v[0][r]; x[2 * r + 1] = ps.v[1][r]; }
        }
        else {
            sym_fft<N, true>(sm.buf, sm.tw, tid);
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = sm.buf[spad(eg * N + tt + TG * i)];
        }
        // ---- 3. gain: statistics per symbol over its N samples ----
        float g_sym;
        if (p.gain_mode == 0) {
            g_sym = 512.0f;
        }
        else {
            // A symbol is owned by TG threads: LANES lanes of WPS warps, or (TG < 32)
            // one LANES-wide segment ("slot") of a warp that holds SPW symbols.
            constexpr int LANES = TG < 32 ? TG : 32;
            constexpr int WPS = TG < 32 ? 1 : TG / 32;
            constexpr int SPW = TG < 32 ? 32 / TG : 1;
            const int warp = tid >> 5, lane = tid & 31;
            const int slot = lane / LANES;
            const int w0 = TG < 32 ? warp : eg * WPS;   // first warp of my symbol
            const bool seg_leader = (lane & (LANES - 1)) == 0;
            if (p.gain_mode == 1) {
                float mn = x[0].x, mx = x[0].x;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    mn = fminf(mn, fminf(x[i].x, x[i].y));
                    mx = fmaxf(mx, fmaxf(x[i].x, x[i].y));
                }
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) {
                    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                if (seg_leader) { sm.red[warp][slot][0] = mn; sm.red[warp][slot][1] = mx; }
                __syncthreads();
                if (tid < G) {
                    const int ww = TG < 32 ? tid / SPW : tid * WPS, ss = TG < 32 ? tid % SPW : 0;
                    float a = sm.red[ww][ss][0], b = sm.red[ww][ss][1];
                    for (int w = 1; w < WPS; w++) {
                        a = fminf(a, sm.red[ww + w][ss][0]);
                        b = fmaxf(b, sm.red[ww + w][ss][1]);
                    }
                    const float m = fmaxf(-a, b);
                    sm.gain[tid] = ((int)m != 0) ? 32767.0f / m : 1.0f;
                }
            }
            else {
                // two-pass mean / variance of re and im (GainControl.cpp:251-340)
                float sr = 0.f, si = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i++) { sr += x[i].x; si += x[i].y; }
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) {
                    sr += __shfl_xor_sync(0xffffffffu, sr, o);
                    si += __shfl_xor_sync(0xffffffffu, si, o);
                }
                if (seg_leader) { sm.red[warp][slot][0] = sr; sm.red[warp][slot][1] = si; }
                __syncthreads();
                float mr = 0.f, mi = 0.f;
                for (int w = 0; w < WPS; w++) { mr += sm.red[w0 + w][slot][0]; mi += sm.red[w0 + w][slot][1]; }
                mr *= 1.0f / N; mi *= 1.0f / N;
                float vr = 0.f, vi = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float dr = x[i].x - mr, di = x[i].y - mi;
                    vr = fmaf(dr, dr, vr); vi = fmaf(di, di, vi);
                }
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) {
                    vr += __shfl_xor_sync(0xffffffffu, vr, o);
                    vi += __shfl_xor_sync(0xffffffffu, vi, o);
                }
                if (seg_leader) { sm.red[warp][slot][2] = vr; sm.red[warp][slot][3] = vi; }
                __syncthreads();
                if (tid < G) {
                    const int ww = TG < 32 ? tid / SPW : tid * WPS, ss = TG < 32 ? tid % SPW : 0;
                    float a = 0.f, b = 0.f;
                    for (int w = 0; w < WPS; w++) { a += sm.red[ww + w][ss][2]; b += sm.red[ww + w][ss][3]; }
                    const float sdr = p.var_factor * sqrtf(a * (1.0f / N));
                    const float sdi = p.var_factor * sqrtf(b * (1.0f / N));
                    // NULL detection looks at the real part only (GainControl.cpp:331)
                    sm.gain[tid] = ((int)sdr != 0) ? 32767.0f / fmaxf(sdr, sdi) : 1.0f;
                }
            }
            __syncthreads();
            // the null symbol borrows the gain of symbol 1 (GainControl.cpp:139-144)
            if (G > 1) g_sym = (s0 + eg == 0) ? sm.gain[1] : sm.gain[eg];
            else g_sym = sm.gain[0];
        }
        if (G == 1) {
            if (grp == 1) gain_sym1 = g_sym;
            if (grp == 0) g_sym = gain_sym1;
        }
        g_sym *= p.gain_const;

        // ---- 4. guard interval + store ----
        if (W == 0) {
            const int s = s0 + eg;
            if (s <= p.L) {
                const int size = s == 0 ? p.null_size : p.sym_size;
                const int pre = size - N;
                const size_t pos = out_base + sym_pos(p, s);
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int n = tt + TG * i;
                    const float2 v = make_float2(x[i].x * g_sym, x[i].y * g_sym);
                    store_sample<POST>(p.out, pos + pre + n, v, p.post, clip);
                    if (n >= N - pre) store_sample<POST>(p.out, pos + n - (N - pre), v, p.post, clip);
                }
            }
            __syncthreads();
        }
        else if (OPT) {
            // Windowed guard interval.  With pos = start of symbol s in the TF and `size` its
            // length, the symbol's cyclic extension ext(o) = x[(o - pre) mod N] covers
            // o in [-W, size + W); it rises with win[o + W] over [-W, W) and falls with
            // win[2W-1-(o-(size-W))] over [size-W, size+W); overlapping edges of neighbours
            // add.  The owner of symbol s writes [pos - W, pos + size - W): both
            // contributions of its leading overlap, none of the trailing one.  The null
            // symbol has no rising edge, the last symbol no falling edge.
            // (each thread overwrites exactly the elements it loaded into x[])
            if (what != GAIN_ONLY) {
#pragma unroll
                for (int i = 0; i < 16; i++)
                    sm.buf[spad(eg * N + tt + TG * i)] = make_float2(x[i].x * g_sym, x[i].y * g_sym);
            }
            __syncthreads();
            if (what != GAIN_ONLY) {
                for (int g = 0; g < G; g++) {
                    const int s = s0 + g;
                    if (s > p.L) break;
                    const int size = s == 0 ? p.null_size : p.sym_size;
                    const int pre = size - N;
                    const bool first = s == 0, last = s == p.L;
                    const float2 *cur = so.tail[tail_par];
                    float2 *nxt = so.tail[tail_par ^ 1];
                    const float2 *xs = sm.buf;
                    if (what == EMIT) {
                        const long long pos = (long long)out_base + sym_pos(p, s);
                        for (int o = (first ? 0 : -W) + tid; o < (last ? size : size - W); o += SYM_THREADS) {
                            int idx = o - pre;
                            if (idx < 0) idx += N;
                            float2 v = xs[spad(g * N + idx)];
                            if (!first && o < W) {
                                const float w = so.win[o + W];
                                const float2 t = cur[o + W];
                                v = make_float2(fmaf(v.x, w, t.x), fmaf(v.y, w, t.y));
                            }
                            store_sample<POST>(p.out, (size_t)(pos + o), v, p.post, clip);
                        }
                    }
                    if (!last) {
                        for (int i = tid; i < 2 * W; i += SYM_THREADS) {
                            int idx = N - W + i;          // ext(size - W + i)
                            if (idx >= N) idx -= N;
                            const float2 v = xs[spad(g * N + idx)];
                            const float w = so.win[2 * W - 1 - i];
                            nxt[i] = make_float2(v.x * w, v.y * w);
                        }
                    }
                    tail_par ^= 1;
                    __syncthreads();
                }
            }
        }
    }
    if (POST && p.post.format != 0) flush_clip(p.post, clip);
}

} // namespace dabmod

namespace dabmod {

// ---------------------------------------------------------------------------
// k_fir: real-tap FIR on interleaved I/Q, anti-causal, zero-extended at the end
// of each TF, no state across TFs:  out[n] = sum_j taps[j] * in[n + j]
// Reference: FIRFilter.cpp:144-192 (taps: :59-71, :95-141).
//
// One CTA = one tile of FIR_TILE consecutive samples of one TF.  The tile and
// its NT-1 sample forward halo are staged in shared memory; each thread keeps
// FIR_M consecutive outputs in registers (packed I/Q pairs) and slides over
// FIR_M+NT-1 inputs, so every input is read from shared memory once per thread
// ((FIR_M+NT-1)/FIR_M loads per output) and every (tap, tap) pair is a uniform-
// register operand of one FFMA2.  Accumulation runs in ascending
// tap order like the reference's inner loop.
// ---------------------------------------------------------------------------
constexpr int FIR_THREADS = 128;
constexpr int FIR_M = 16;
constexpr int FIR_TILE = FIR_THREADS * FIR_M;

struct FirParams {
    const float2 *in;     // n_tf * tf_samples
    void *out;
    int tf_samples;
    int tiles_per_tf;
    int ntaps;
    float2 taps[MAX_FIR_TAPS];  // (tap, tap) pairs, zero padded
    PostParams post;
};

// Blackwell packed FP32 (sm_100 __ffma2_rn -> SASS FFMA2): one instruction does the
// I and the Q multiply-accumulate of a tap (same rounding as two fmaf), halving
// the issue slots of the FIR inner loop.
// shared-memory index padding for 8-byte elements read with a lane stride of FIR_M:
// thread t starts at 17*t, an odd stride, so a half-warp covers all 16 bank pairs
static_assert(FIR_M == 16, "fpad assumes FIR_M == 16");
__device__ __forceinline__ constexpr int fpad(int i) { return i + (i >> 4); }

// NT > 0: tap count known at compile time, tap loop fully unrolled (each input is
// loaded once per thread).  NT == 0: any tap count up to MAX_FIR_TAPS, taps consumed
// in chunks of FIR_CHUNK with the chunk body unrolled.
constexpr int FIR_CHUNK = 16;

template <int NT, bool POST>
__global__ void __launch_bounds__(FIR_THREADS) k_fir(const __grid_constant__ FirParams p)
{
    constexpr int NTMAX = NT > 0 ? NT : MAX_FIR_TAPS;
    constexpr int SPAN = FIR_TILE + NTMAX;      // samples staged (>= tile + halo, kept even)
    constexpr int XS_IN = SPAN + SPAN / 16 + 8;            // padded input window
    constexpr int XS_OUT = FIR_THREADS * (FIR_M + 2);      // padded output staging (float4 slots 9*t)
    __shared__ __align__(16) float2 xs[XS_IN > XS_OUT ? XS_IN : XS_OUT];
    const int tid = threadIdx.x;
    const int tf = blockIdx.x / p.tiles_per_tf;
    const int tile = blockIdx.x - tf * p.tiles_per_tf;
    const int n0 = tile * FIR_TILE;
    const float2 *in = p.in + (size_t)tf * p.tf_samples;
    const int span = NT > 0 ? SPAN : FIR_TILE + ((p.ntaps + FIR_CHUNK - 1) / FIR_CHUNK) * FIR_CHUNK;

    // stage: two samples (16 B) per thread per step, zero beyond the TF end
    for (int i = 2 * tid; i < span; i += 2 * FIR_THREADS) {
        const int n = n0 + i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n + 1 < p.tf_samples) {
            v = __ldg(reinterpret_cast<const float4 *>(in + n));
        }
        else if (n < p.tf_samples) {
            const float2 a = __ldg(in + n);
            v.x = a.x; v.y = a.y;
        }
        xs[fpad(i)] = make_float2(v.x, v.y);
        xs[fpad(i + 1)] = make_float2(v.z, v.w);
    }
    __syncthreads();

    float2 acc[FIR_M];
#pragma unroll
    for (int m = 0; m < FIR_M; m++) acc[m] = make_float2(0.f, 0.f);
    if (NT > 0) {
        const float2 *x = xs + tid * (FIR_M + 1);    // fpad(tid * FIR_M)
#pragma unroll
        for (int i = 0; i < FIR_M + NTMAX - 1; i++) {
            const float2 v = x[i + (i >> 4)];
#pragma unroll
            for (int m = 0; m < FIR_M; m++) {
                const int j = i - m;
                if (j >= 0 && j < NTMAX) acc[m] = __ffma2_rn(v, p.taps[j], acc[m]);
            }
        }
    }
    else {
        // ascending tap order is kept: chunk c covers taps [16c, 16c+16)
        for (int c = 0; c * FIR_CHUNK < p.ntaps; c++) {
            const int i0 = tid * FIR_M + c * FIR_CHUNK;
#pragma unroll
            for (int i = 0; i < FIR_M + FIR_CHUNK - 1; i++) {
                const float2 v = xs[fpad(i0 + i)];
#pragma unroll
                for (int m = 0; m < FIR_M; m++) {
                    const int j = i - m;
                    if (j >= 0 && j < FIR_CHUNK) acc[m] = __ffma2_rn(v, p.taps[c * FIR_CHUNK + j], acc[m]);
                }
            }
        }
    }
    // Transpose through shared memory so that the global stores are coalesced:
    // thread t parks its FIR_M outputs at 16-byte slots 9*t + m/2 (odd slot stride:
    // conflict-free float4 writes), then the CTA streams the tile out in order.
    __syncthreads();
    float4 *ys = reinterpret_cast<float4 *>(xs);
#pragma unroll
    for (int m = 0; m < FIR_M; m += 2)
        ys[tid * (FIR_M / 2 + 1) + m / 2] = make_float4(acc[m].x, acc[m].y, acc[m + 1].x, acc[m + 1].y);
    __syncthreads();

    unsigned clip = 0;
    const size_t obase = (size_t)tf * p.tf_samples + n0;
#pragma unroll
    for (int k = 0; k < FIR_M / 2; k++) {
        const int q = k * FIR_THREADS + tid;          // pair index within the tile
        const int n = 2 * q;                          // sample index within the tile
        const float4 v = ys[q + (q >> 3)];
        if (!POST && n0 + n + 1 < p.tf_samples) {
            reinterpret_cast<float4 *>(reinterpret_cast<float2 *>(p.out) + obase)[q] = v;
        }
        else {
            if (n0 + n < p.tf_samples) store_sample<POST>(p.out, obase + n, make_float2(v.x, v.y), p.post, clip);
            if (n0 + n + 1 < p.tf_samples) store_sample<POST>(p.out, obase + n + 1, make_float2(v.z, v.w), p.post, clip);
        }
    }
    if (POST && p.post.format != 0) flush_clip(p.post, clip);
}

// ---------------------------------------------------------------------------
// k_tii_fill: the TII null symbol for the warp-per-symbol kernels.  The TII symbol (TII.cpp:213-245) carries the
// phase reference on a fixed set of carriers and borrows the gain of symbol 1 (GainControl.cpp:139-144), which is
// the phase reference symbol itself: neither depends on the frame's bits, so the null symbol of every TII frame of
// a stream is ONE constant vector of null_size samples.  It is produced once per call by the general kernel
// (k_symbols on a single frame) and this kernel copies it into every frame where the TII toggles on
// (TII.cpp:225-242: every second frame), through the configured epilogue.
// ---------------------------------------------------------------------------
template <bool POST>
__global__ void __launch_bounds__(256) k_tii_fill(const float2 *null_sym, void *out, int null_size, int tf_samples, int n_tf,
                                                  unsigned long long tf_offset, PostParams post)
{
    unsigned clip = 0;
    const int per_tf = (null_size + 255) / 256;
    for (long long b = blockIdx.x; b < (long long)n_tf * per_tf; b += gridDim.x) {
        const int tf = (int)(b / per_tf), i = (int)(b - (long long)tf * per_tf) * 256 + threadIdx.x;
        if (((tf_offset + tf) & 1) != 0 || i >= null_size) continue;
        store_sample<POST>(out, (size_t)tf * tf_samples + i, __ldg(null_sym + i), post, clip);
    }
    if (POST && post.format != 0) flush_clip(post, clip);
}

// ---------------------------------------------------------------------------
// k_fir_long: tap counts above MAX_FIR_TAPS (FIRFilter::load_filter_taps takes any count,
// src/FIRFilter.cpp:95-141).  Same tile, same ascending tap order and the same packed multiply-adds as
// k_fir<0>; the taps come from global memory (uniform loads, a chunk of 16 at a time) and the input window
// of FIR_TILE + ntaps samples lives in dynamic shared memory.
// ---------------------------------------------------------------------------
constexpr int MAX_FIR_TAPS_LONG = 16384;

struct FirLongParams {
    const float2 *in;
    void *out;
    int tf_samples;
    int tiles_per_tf;
    int ntaps;
    const float *taps;    // ntaps floats, zero padded to a multiple of FIR_CHUNK (device)
    PostParams post;
};

inline size_t fir_long_smem(int ntaps)
{
    const int span = FIR_TILE + ((ntaps + FIR_CHUNK - 1) / FIR_CHUNK) * FIR_CHUNK;
    return sizeof(float2) * (size_t)(span + span / 16 + 8);
}

template <bool POST>
__global__ void __launch_bounds__(FIR_THREADS) k_fir_long(const __grid_constant__ FirLongParams p)
{
    extern __shared__ __align__(16) unsigned char fir_long_raw[];
    float2 *xs = reinterpret_cast<float2 *>(fir_long_raw);
    const int tid = threadIdx.x;
    const int tf = blockIdx.x / p.tiles_per_tf;
    const int tile = blockIdx.x - tf * p.tiles_per_tf;
    const int n0 = tile * FIR_TILE;
    const float2 *in = p.in + (size_t)tf * p.tf_samples;
    const int span = FIR_TILE + ((p.ntaps + FIR_CHUNK - 1) / FIR_CHUNK) * FIR_CHUNK;

    for (int i = 2 * tid; i < span; i += 2 * FIR_THREADS) {
        const int n = n0 + i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n + 1 < p.tf_samples) v = __ldg(reinterpret_cast<const float4 *>(in + n));
        else if (n < p.tf_samples) {
            const float2 a = __ldg(in + n);
            v.x = a.x; v.y = a.y;
        }
        xs[fpad(i)] = make_float2(v.x, v.y);
        xs[fpad(i + 1)] = make_float2(v.z, v.w);
    }
    __syncthreads();

    float2 acc[FIR_M];
#pragma unroll
    for (int m = 0; m < FIR_M; m++) acc[m] = make_float2(0.f, 0.f);
    for (int c = 0; c * FIR_CHUNK < p.ntaps; c++) {
        float tp[FIR_CHUNK];
#pragma unroll
        for (int j = 0; j < FIR_CHUNK; j++) tp[j] = __ldg(p.taps + c * FIR_CHUNK + j);
        const int i0 = tid * FIR_M + c * FIR_CHUNK;
#pragma unroll
        for (int i = 0; i < FIR_M + FIR_CHUNK - 1; i++) {
            const float2 v = xs[fpad(i0 + i)];
#pragma unroll
            for (int m = 0; m < FIR_M; m++) {
                const int j = i - m;
                if (j >= 0 && j < FIR_CHUNK) acc[m] = __ffma2_rn(v, make_float2(tp[j], tp[j]), acc[m]);
            }
        }
    }
    __syncthreads();
    float4 *ys = reinterpret_cast<float4 *>(xs);
#pragma unroll
    for (int m = 0; m < FIR_M; m += 2)
        ys[tid * (FIR_M / 2 + 1) + m / 2] = make_float4(acc[m].x, acc[m].y, acc[m + 1].x, acc[m + 1].y);
    __syncthreads();

    unsigned clip = 0;
    const size_t obase = (size_t)tf * p.tf_samples + n0;
#pragma unroll
    for (int k = 0; k < FIR_M / 2; k++) {
        const int q = k * FIR_THREADS + tid;
        const int n = 2 * q;
        const float4 v = ys[q + (q >> 3)];
        if (!POST && n0 + n + 1 < p.tf_samples) {
            reinterpret_cast<float4 *>(reinterpret_cast<float2 *>(p.out) + obase)[q] = v;
        }
        else {
            if (n0 + n < p.tf_samples) store_sample<POST>(p.out, obase + n, make_float2(v.x, v.y), p.post, clip);
            if (n0 + n + 1 < p.tf_samples) store_sample<POST>(p.out, obase + n + 1, make_float2(v.z, v.w), p.post, clip);
        }
    }
    if (POST && p.post.format != 0) flush_clip(p.post, clip);
}

// ---------------------------------------------------------------------------
// k_fir_sym: the same filter for TM I fed by k_symbols_w in its compact layout
// ([tf][symbol 1..L][N] samples, no null symbol, no cyclic prefix).  One CTA = one
// OFDM symbol.  Two facts about the guard-interval stream make this cheaper than
// filtering it sample by sample, with bit-identical results:
//   * the cyclic prefix is a copy of the symbol's tail and the filter is shift
//     invariant, so the outputs over the prefix, except the last NT-1, ARE the
//     outputs over the tail (same operands, same order): 460 of a symbol's 2552
//     outputs are copies, not computed (-17 % multiply-adds);
//   * the prefix need not exist in memory: the filter input is N samples per
//     symbol instead of sym_size (-21 % reads), and the symbol kernel writes 21 % less.
// Per symbol: N outputs over the body (window reaching into the next symbol's prefix
// = that symbol's samples N-pre ...), NT-1 outputs across the prefix/body seam
// (a circular window), and for symbol 1 the NT-1 outputs at the end of the null
// symbol.  Zeros (null symbol, beyond the TF end) enter as 0 * tap like in k_fir.
// ---------------------------------------------------------------------------
struct FirSymParams {
    const float2 *in;     // n_tf * L * N samples, compact
    void *out;            // n_tf * tf_samples
    int L, null_size, sym_size, tf_samples;
    float2 taps[MAX_FIR_TAPS];
    PostParams post;
};

constexpr int FIRS_THREADS = FIR_THREADS + 32;   // four warps on the body, one on the seams
constexpr int FIRS_SM = 4;                        // seam outputs per lane of the fifth warp

template <int NT, bool POST>
__global__ void __launch_bounds__(FIRS_THREADS, 8) k_fir_sym(const __grid_constant__ FirSymParams p)
{
    constexpr int N = FIR_TILE;                 // 2048 = TM I spacing
    constexpr int H = NT - 1;                   // halo
    constexpr int XS_IN = N + NT + (N + NT) / 16 + 8;
    constexpr int XS_OUT = FIR_THREADS * (FIR_M + 2);        // padded output staging (float4 slots 9*t)
    constexpr int SEAM_LANES = (H + FIRS_SM - 1) / FIRS_SM;  // 11 lanes x 4 outputs per seam
    static_assert(2 * SEAM_LANES <= 32 && SEAM_LANES <= 16, "the seams are one warp's work");
    // body + next symbol's prefix head (fpad layout); later the transposed staging of the body outputs
    __shared__ __align__(16) float2 xs[XS_IN > XS_OUT ? XS_IN : XS_OUT];
    __shared__ __align__(16) float2 seam[2][SEAM_LANES * FIRS_SM];   // outputs of the two seams
    const int tid = threadIdx.x;
    const int tf = blockIdx.y;                               // grid: (L, n_tf)
    const int s = 1 + blockIdx.x;                            // symbol 1..L
    const int pre = p.sym_size - N;                          // cyclic prefix length
    const float2 *body = p.in + ((size_t)tf * p.L + (s - 1)) * N;
    const bool has_next = s < p.L;

    // stage the head of the next symbol's prefix (zeros beyond the TF; fetched first so that its latency
    // overlaps the body's) and the body
    float2 halo = make_float2(0.f, 0.f);
    if (has_next && tid < H) halo = __ldg(body + N + (N - pre) + tid);
    {
        // all loads in flight before the first shared store (a rolled loop would serialise the round trips)
        constexpr int NLD = (N / 2 + FIRS_THREADS - 1) / FIRS_THREADS;
        float4 v[NLD];
#pragma unroll
        for (int k = 0; k < NLD; k++) {
            const int i = 2 * (tid + k * FIRS_THREADS);
            v[k] = i < N ? __ldg(reinterpret_cast<const float4 *>(body + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < NLD; k++) {
            const int i = 2 * (tid + k * FIRS_THREADS);
            if (i < N) {
                xs[fpad(i)] = make_float2(v[k].x, v[k].y);
                xs[fpad(i + 1)] = make_float2(v[k].z, v[k].w);
            }
        }
    }
    if (tid < NT) xs[fpad(N + tid)] = halo;
    __syncthreads();

    float2 acc[FIR_M];
#pragma unroll
    for (int m = 0; m < FIR_M; m++) acc[m] = make_float2(0.f, 0.f);
    if (tid < FIR_THREADS) {
        const float2 *x = xs + tid * (FIR_M + 1);            // fpad(tid * FIR_M)
#pragma unroll
        for (int i = 0; i < FIR_M + NT - 1; i++) {
            const float2 v = x[i + (i >> 4)];
#pragma unroll
            for (int m = 0; m < FIR_M; m++) {
                const int j = i - m;
                if (j >= 0 && j < NT) acc[m] = __ffma2_rn(v, p.taps[j], acc[m]);
            }
        }
    }
    else {
        // The seams, ascending tap order like everything else, zeros entering as 0 * tap like in k_fir:
        //   lanes 0-10:  the last NT-1 prefix positions -- window over the prefix end (= body end), then the body start;
        //   lanes 16-26 (symbol 1): the last NT-1 positions of the null symbol -- zeros, then the prefix start.
        const int lane = tid - FIR_THREADS, which = lane >> 4, l = lane & 15;
        if (l < SEAM_LANES && (which == 0 || s == 1)) {
            const int i0 = FIRS_SM * l;
#pragma unroll
            for (int i = 0; i < FIRS_SM + NT - 1; i++) {
                const int t = i0 + i;                        // position in the seam's 2H-sample input
                float2 v = make_float2(0.f, 0.f);
                if (which == 0) v = xs[fpad(t < H ? N - H + t : t - H)];
                else if (t >= H) v = xs[fpad(N - pre + t - H)];
#pragma unroll
                for (int m = 0; m < FIRS_SM; m++) {
                    const int j = i - m;
                    if (j >= 0 && j < NT) acc[m] = __ffma2_rn(v, p.taps[j], acc[m]);
                }
            }
#pragma unroll
            for (int m = 0; m < FIRS_SM; m++) seam[which][i0 + m] = acc[m];
        }
    }
    __syncthreads();                                         // everybody is done with xs
    float4 *ys = reinterpret_cast<float4 *>(xs);
    if (tid < FIR_THREADS) {
#pragma unroll
        for (int m = 0; m < FIR_M; m += 2)
            ys[tid * (FIR_M / 2 + 1) + m / 2] = make_float4(acc[m].x, acc[m].y, acc[m + 1].x, acc[m + 1].y);
    }
    __syncthreads();

    // ---- store: [null symbol (s == 1)] | prefix | body, coalesced pairs ----
    unsigned clip = 0;
    const size_t tf_base = (size_t)tf * p.tf_samples;
    const size_t pos = tf_base + p.null_size + (size_t)(s - 1) * p.sym_size;
    auto emit2 = [&](size_t at, float4 v) {                  // two samples at an even position
        if (!POST) *reinterpret_cast<float4 *>(reinterpret_cast<float2 *>(p.out) + at) = v;
        else {
            store_sample<POST>(p.out, at, make_float2(v.x, v.y), p.post, clip);
            store_sample<POST>(p.out, at + 1, make_float2(v.z, v.w), p.post, clip);
        }
    };
    for (int q = tid; q < N / 2; q += FIRS_THREADS) emit2(pos + pre + 2 * q, ys[q + (q >> 3)]);      // body
    // prefix: copies of the outputs over the tail (pairs stay aligned: N - pre, pre and H are even), then the seam
    const float4 *seam4 = reinterpret_cast<const float4 *>(seam[0]);
    for (int q = tid; q < pre / 2; q += FIRS_THREADS) {
        const int qb = (N - pre) / 2 + q;
        emit2(pos + 2 * q, q < (pre - H) / 2 ? ys[qb + (qb >> 3)] : seam4[q - (pre - H) / 2]);
    }
    if (s == 1) {                                            // null symbol: zeros, then its seam with symbol 1
        const float4 *seam4n = reinterpret_cast<const float4 *>(seam[1]);
        for (int q = tid; q < p.null_size / 2; q += FIRS_THREADS)
            emit2(tf_base + 2 * q, q < (p.null_size - H) / 2 ? make_float4(0.f, 0.f, 0.f, 0.f)
                                                             : seam4n[q - (p.null_size - H) / 2]);
    }
    if (POST && p.post.format != 0) flush_clip(p.post, clip);
}

// ---------------------------------------------------------------------------
// k_fir_tma: k_fir_sym's work (TM I, compact input, complexf output) as a persistent, TMA-fed pipeline.
//
// Both FIR kernels above stop at ~0.63 ms with neither HBM, the FMA pipe nor issue saturated: inside a CTA the
// chain load -> filter -> transpose -> store is serial and the padded shared layout keeps every byte on the LSU.
// Here a lane owns 17 consecutive outputs instead of 16: the lane stride is odd, so the window needs NO padding,
// is conflict free as it lies in memory, and can therefore be moved by the bulk-copy engine in both directions:
//   * input: cp.async.bulk global -> shared of the symbol's 2048 samples, the 44-sample head of the next symbol's
//     prefix and the two seam strips, completing on an mbarrier; two stages, fetched two symbols ahead;
//   * output: every thread parks its 17 results in natural order (stride 17 again), then three bulk copies
//     shared -> global land body, prefix (= the outputs over the tail, straight from the same staging) and seam.
// No global load or store instruction is left in the steady state; a CTA walks a contiguous range of symbols.
// Threads 0..120 take the 2048 body outputs, 121..123 the 44 outputs of the prefix/body seam, 124..126 (symbol 1)
// the 44 at the end of the null symbol -- one code path, different base pointers.  Same operands in the same
// order as k_fir: bit-identical.
// ---------------------------------------------------------------------------
constexpr int FIRT_THREADS = 128;
constexpr int FIRT_M = 17;
constexpr int FIRT_IN = 2048 + 44 + 12;           // body | next prefix head | slack for the last body thread's window
constexpr int FIRT_STRIP = 96;                    // 2 * 44 seam inputs + slack
constexpr int FIRT_OUT = 2048 + 16;               // body outputs + slack for the last body thread
constexpr int FIRT_SEAM = 64;
constexpr int FIRT_CTAS_PER_SM = 3;

struct FirTmaStage {
    float2 in[FIRT_IN];
    float2 strip[2][FIRT_STRIP];                  // [0]: body end ++ body start; [1]: zeros ++ prefix start (symbol 1)
    float2 out[FIRT_OUT];
    float2 seam[2][FIRT_SEAM];
};
struct FirTmaSmem {
    FirTmaStage st[2];
    unsigned long long full[2];                   // mbarriers: stage's input has landed
};

__device__ __forceinline__ uint32_t firt_saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void firt_load(void *sdst, const void *gsrc, int bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(firt_saddr(sdst)), "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(firt_saddr(bar)) : "memory");
}
__device__ __forceinline__ void firt_store(void *gdst, const void *ssrc, int bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(__cvta_generic_to_global(gdst)), "r"(firt_saddr(ssrc)), "r"(bytes) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(FIRT_THREADS, FIRT_CTAS_PER_SM) k_fir_tma(const __grid_constant__ FirSymParams p, long long n_items)
{
    constexpr int N = 2048, H = NT - 1;
    static_assert(H == 44, "buffer geometry is laid out for the 45 default taps");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FirTmaSmem &sm = *reinterpret_cast<FirTmaSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const int pre = p.sym_size - N;
    const long long per = (n_items + gridDim.x - 1) / gridDim.x;
    const long long it0 = (long long)blockIdx.x * per;
    const long long it1 = it0 + per < n_items ? it0 + per : n_items;
    if (it0 >= it1) return;

    // item = tf * L + (s - 1): everything the symbol's filter reads, by bulk copies onto one mbarrier
    auto fetch = [&](long long item, int stage) {
        const int s = 1 + (int)(item % p.L);
        const float2 *body = p.in + (size_t)item * N;
        FirTmaStage &st = sm.st[stage];
        unsigned bytes = N * 8 + 2 * H * 8;
        if (s < p.L) bytes += H * 8;
        if (s == 1) bytes += H * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(firt_saddr(&sm.full[stage])), "r"(bytes) : "memory");
        firt_load(st.in, body, N * 8, &sm.full[stage]);
        if (s < p.L) firt_load(st.in + N, body + N + (N - pre), H * 8, &sm.full[stage]);      // next symbol's prefix head
        firt_load(st.strip[0], body + N - H, H * 8, &sm.full[stage]);                          // prefix end = body end
        firt_load(st.strip[0] + H, body, H * 8, &sm.full[stage]);                              // ... then the body start
        if (s == 1) firt_load(st.strip[1] + H, body + (N - pre), H * 8, &sm.full[stage]);      // null symbol: zeros, then the prefix
    };

    for (int i = tid; i < 2 * FIRT_STRIP; i += FIRT_THREADS) {
        sm.st[0].strip[i / FIRT_STRIP][i % FIRT_STRIP] = make_float2(0.f, 0.f);
        sm.st[1].strip[i / FIRT_STRIP][i % FIRT_STRIP] = make_float2(0.f, 0.f);
    }
    for (int i = tid; i < FIRT_IN - N; i += FIRT_THREADS) {
        sm.st[0].in[N + i] = make_float2(0.f, 0.f);
        sm.st[1].in[N + i] = make_float2(0.f, 0.f);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(firt_saddr(&sm.full[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(firt_saddr(&sm.full[1])) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // zeros and barrier inits, before any bulk copy
    __syncthreads();
    if (tid == 0) {
        fetch(it0, 0);
        if (it0 + 1 < it1) fetch(it0 + 1, 1);
    }

    // this thread's share: 17 consecutive outputs of the body, or of a seam
    const int cls = tid < 121 ? 0 : tid < 124 ? 1 : tid < 127 ? 2 : 3;
    const int k0 = cls == 0 ? FIRT_M * tid : cls == 1 ? FIRT_M * (tid - 121) : FIRT_M * (tid - 124);

    for (long long item = it0; item < it1; item++) {
        const int it = (int)(item - it0), stage = it & 1;
        const unsigned parity = (unsigned)(it >> 1) & 1u;
        FirTmaStage &st = sm.st[stage];
        const int tf = (int)(item / p.L);
        const int s = 1 + (int)(item - (long long)tf * p.L);
        // A. the stage's input has landed
        {
            unsigned done = 0;
            const uint32_t bar = firt_saddr(&sm.full[stage]);
            while (!done)
                asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        if (s == p.L) {                                      // TF end: the window runs into zeros (FIRFilter.cpp:186-191)
            if (tid < H) st.in[N + tid] = make_float2(0.f, 0.f);
            __syncthreads();
        }
        // B. filter
        float2 acc[FIRT_M];
#pragma unroll
        for (int m = 0; m < FIRT_M; m++) acc[m] = make_float2(0.f, 0.f);
        if (cls == 0 || cls == 1 || (cls == 2 && s == 1)) {
            const float2 *x = (cls == 0 ? st.in : st.strip[cls - 1]) + k0;
#pragma unroll
            for (int i = 0; i < FIRT_M + NT - 1; i++) {
                const float2 v = x[i];
#pragma unroll
                for (int m = 0; m < FIRT_M; m++) {
                    const int j = i - m;
                    if (j >= 0 && j < NT) acc[m] = __ffma2_rn(v, p.taps[j], acc[m]);
                }
            }
        }
        // C. park the results in natural order (lane stride 17: conflict free)
        {
            float2 *y = cls == 0 ? st.out + k0 : st.seam[cls == 2 ? 1 : 0] + k0;
            if (cls != 3) {
#pragma unroll
                for (int m = 0; m < FIRT_M; m++) y[m] = acc[m];
            }
        }
        // null symbol of this TF: zeros up to the seam (plain stores, one symbol in 76)
        const size_t tf_base = (size_t)tf * p.tf_samples;
        if (s == 1) {
            float4 *z = reinterpret_cast<float4 *>(reinterpret_cast<float2 *>(p.out) + tf_base);
            for (int q = tid; q < (p.null_size - H) / 2; q += FIRT_THREADS) z[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // D. the previous symbol's bulk stores have read their staging (so the NEXT symbol may overwrite it), this
        //    symbol's staging is complete and visible to the copy engine, nobody reads the input stage any more
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        // E. results out, input for the symbol after next in
        if (tid == 0) {
            float2 *g = reinterpret_cast<float2 *>(p.out) + tf_base + p.null_size + (size_t)(s - 1) * p.sym_size;
            firt_store(g + pre, st.out, N * 8);                                   // body
            firt_store(g, st.out + (N - pre), (pre - H) * 8);                     // prefix = outputs over the tail
            firt_store(g + (pre - H), st.seam[0], H * 8);                         // prefix/body seam
            if (s == 1)
                firt_store(reinterpret_cast<float2 *>(p.out) + tf_base + (p.null_size - H), st.seam[1], H * 8);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (item + 2 < it1) fetch(item + 2, stage);
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

} // namespace dabmod
