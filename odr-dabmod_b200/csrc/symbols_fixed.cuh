// k_symbols_fix: the symbol kernel of the FIXED-POINT engine: the chain
// DabModulator wires for FFTEngine::KISS (src/DabModulator.cpp:144-224) --
//   QpskSymbolMapper / PhaseReference / FrequencyInterleaver / DifferentialModulator on complexfix
//   = std::complex<fpm::fixed<int16, int32, 14>> (src/Buffer.h:42-43), NullSymbol / TII,
//   OfdmGeneratorFixed (src/OfdmGenerator.cpp:467-579: the vendored KISS FFT with FIXED_POINT=16),
//   no GainControl, GuardIntervalInserter do_process<complexfix> (src/GuardIntervalInserter.cpp:115-323)
// -- bits in, int16 I/Q out, every bit equal to the reference's.
//
// * Carriers: the fixed-point product chain of the differential modulator only visits
//   {0, +-11585, +-16384} (11585 = fixed(1/sqrt 2); fpm's rounding multiply gives 11585^2 -> 8192,
//   16384 * 11585 -> 11585), the images of the float chain's {0, +-1/sqrt2, +-1}: the same integer
//   phase arithmetic as k_symbols (units of pi/4) with an 8-entry int16 value table.
// * IFFT: KISS is a decimation-in-time transform (kiss_fft.c:236-291): the input is read in
//   mixed-radix digit-reversed order into the output array, then one pass of radix-2 (N = 2048, 512)
//   and passes of radix-4 butterflies run in place from the smallest sub-transform up, each dividing
//   its inputs by the radix (DIVSCALAR: x * (32767 / p), rounded, >> 15) and rounding every twiddle
//   product once (C_MUL / sround).  Integer rounding is not associative, so the kernel performs exactly
//   these operations in exactly this order (kf_bfly2 / kf_bfly4 below), on int16 pairs in shared
//   memory; the carriers are scattered straight to their digit-reversed positions.
// * A CTA works on one (TF, chunk of symbol groups); a group is 2048 / N symbols side by side in one
//   2048-point buffer, so every mode keeps all threads busy (512 radix-4 butterflies per pass).
// * Windowing: symbol l's rising edge (2W samples around its start) is ADDED to the falling edge of
//   symbol l-1, both multiplied by the fixed(raised cosine) window with fpm's rounding multiply;
//   the falling edge comes from the previous symbol of the group or from a shared tail kept across
//   groups (one chunk per TF when W > 0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dabmod {

constexpr int FX_THREADS = 128;
constexpr int FX_POINTS = 2048;
constexpr int FX_MAX_STAGES = 8;
constexpr int FX_MAX_WINDOW = 1024;           // 2 W <= 2 * (sym_size - N) of TM I = 1008
// Shared-memory index padding: 8 words after every 128.  A fused pass walks the buffer at strides of
// 8 / 16 / 128 points; without the skew the 128-point blocks of one warp's tasks would alias on the banks.
__device__ __host__ __forceinline__ constexpr int fxpad(int i) { return i + ((i >> 7) << 3); }
constexpr int FX_BUF = FX_POINTS + FX_POINTS / 16;

struct FixParams {
    int L, K, N, null_size, sym_size, tf_in_bytes, tf_samples;
    int G, n_groups, groups_per_chunk, n_chunks;
    const uint16_t *pos_of_src;   // K: digit-reversed position (within its symbol) of the bin source carrier j goes to
    const uint8_t *phase0;        // K: phase reference, units of pi/4
    const short2 *tw;             // N: kiss twiddles, kf_cexp(+2 pi i / N)
    int n_stages;                 // kf_factor's list, executed from the last entry to the first
    int stage_p[FX_MAX_STAGES], stage_m[FX_MAX_STAGES];
    int tii_count, tii_parity;
    const uint16_t *tii_pos;      // digit-reversed positions of the TII carriers
    const short2 *tii_val;
    int window;                   // windowOverlap W
    const short *window_tab;      // 2W fixed-point window values (rising edge)
    const uint8_t *bits;
    short2 *out;
    unsigned long long tf_offset;
};

struct FixSmem {
    uint32_t buf[FX_BUF];         // int16 pairs, fxpad layout
    uint32_t tw[FX_POINTS];       // kiss twiddles, int16 pairs, one compact table per radix-4 stage (fx_tw)
    uint32_t spread[256];
    short2 c8[8];
    short2 tail[FX_MAX_WINDOW];   // falling edge of the last symbol of the previous group
    short win[FX_MAX_WINDOW];
};

// fpm::fixed<int16, int32, 14>::operator*= with rounding (fpm/fixed.hpp:156-169)
__device__ __forceinline__ short fx_mul(short a, short b)
{
    const int v = ((int)a * (int)b) / 8192;       // truncating division, like the reference
    return (short)(v / 2 + v % 2);
}

// KISS arithmetic on int pairs.  The reference keeps every intermediate in an int16 (kiss_fft_cpx), i.e.
// reduces it mod 2^16; additions and subtractions commute with that reduction, so the butterflies below work
// on 32-bit values and reduce only where it matters: when a value is loaded (sign extension of the stored
// int16) before it enters a multiplication, and when it is stored (the low 16 bits).  DIVSCALAR and C_MUL
// results of in-range operands fit an int16 by construction (|x| <= 32767 -> |x * 8191 >> 15| <= 8192,
// twiddles <= 32767), so they need no reduction either.
struct fxc { int x, y; };
__device__ __forceinline__ fxc fx_unpack(uint32_t w)
{
    int lo;     // prmt in its default mode: selector nibble 9 = sign of byte 1 replicated (the low half, sign extended)
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(lo) : "r"(w));
    return fxc{lo, (int)w >> 16};
}
__device__ __forceinline__ uint32_t fx_pack(fxc c) { return __byte_perm((uint32_t)c.x, (uint32_t)c.y, 0x5410); }
__device__ __forceinline__ int fx_sround(int x) { return (x + (1 << 14)) >> 15; }
__device__ __forceinline__ fxc fx_fixdiv(fxc c, int k)              // C_FIXDIV: k = 32767 / radix
{
    return fxc{fx_sround(c.x * k), fx_sround(c.y * k)};
}
__device__ __forceinline__ fxc fx_cmul(fxc a, fxc b)               // C_MUL
{
    return fxc{fx_sround(a.x * b.x - a.y * b.y), fx_sround(a.x * b.y + a.y * b.x)};
}
__device__ __forceinline__ fxc fx_add(fxc a, fxc b) { return fxc{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ fxc fx_sub(fxc a, fxc b) { return fxc{a.x - b.x, a.y - b.y}; }

// kf_bfly4 with st->inverse (kiss_fft.c:66-113) on f[0..3] = Fout[0], Fout[m], Fout[2m], Fout[3m]
__device__ __forceinline__ void fx_bfly4(fxc (&f)[4], fxc w1, fxc w2, fxc w3)
{
#pragma unroll
    for (int q = 0; q < 4; q++) f[q] = fx_fixdiv(f[q], 32767 / 4);
    const fxc a0 = fx_cmul(f[1], w1), a1 = fx_cmul(f[2], w2), a2 = fx_cmul(f[3], w3);
    const fxc a5 = fx_sub(f[0], a1);
    const fxc t0 = fx_add(f[0], a1);
    const fxc a3 = fx_add(a0, a2), a4 = fx_sub(a0, a2);
    f[2] = fx_sub(t0, a3);
    f[0] = fx_add(t0, a3);
    f[1] = fxc{a5.x - a4.y, a5.y + a4.x};
    f[3] = fxc{a5.x + a4.y, a5.y - a4.x};
}
// kf_bfly2 (kiss_fft.c:40-64)
__device__ __forceinline__ void fx_bfly2(fxc &f0, fxc &f1, fxc w)
{
    f0 = fx_fixdiv(f0, 32767 / 2);
    f1 = fx_fixdiv(f1, 32767 / 2);
    const fxc t = fx_cmul(f1, w);
    f1 = fx_sub(f0, t);
    f0 = fx_add(f0, t);
}

// Twiddles of the radix-4 stage with sub-transform size M: KISS reads tw[fs k], tw[2 fs k], tw[3 fs k] with
// fs = N / (4 M) (kiss_fft.c kf_bfly4) -- for the early stages a stride of 64 ... 256 words, i.e. every lane of a
// warp on the same bank (ncu, round 1: 8-way conflicts on these loads, 47 % of the kernel's shared wavefronts were
// excess).  The same values are therefore kept stage by stage, contiguous in k: entry (q, k) of stage M at
// fx_tw_off(M) + (q - 1) M + k.  Stage sizes are 1, 4, 16, ... or 2, 8, 32, ...; the tables of all stages below M
// take M - 1 resp. M - 2 words, the whole set fewer than N.
__device__ __host__ constexpr int fx_tw_off(int M) { return M - (((M & 0x55555555) != 0) ? 1 : 2); }
__device__ __forceinline__ uint32_t fx_tw(const uint32_t *tw, int M, int q, int k) { return tw[fx_tw_off(M) + (q - 1) * M + k]; }

// ---- fused passes: two KISS stages per trip through shared memory --------------------------------
// Stages run smallest sub-transform first (kf_work's recursion unwinds that way); a butterfly of the
// second stage needs four outputs of the first, so a thread that owns 16 points k + MA j (j < 16) of one
// 16 MA block can do four butterflies of the stage with m = MA and then four of the stage with m = 4 MA.
// Every butterfly is exactly kf_bfly4 / kf_bfly2: fusing only changes where the values wait in between.

// radix 2 (m = 1) then radix 4 (m = 2): 8 contiguous points
template <int N>
__device__ __forceinline__ void fx_pass_24(uint32_t *buf, const uint32_t *tw, int task)
{
    uint4 *v = reinterpret_cast<uint4 *>(buf + fxpad(8 * task));
    const uint4 lo = v[0], hi = v[1];
    fxc f[8] = {fx_unpack(lo.x), fx_unpack(lo.y), fx_unpack(lo.z), fx_unpack(lo.w),
                fx_unpack(hi.x), fx_unpack(hi.y), fx_unpack(hi.z), fx_unpack(hi.w)};
    const fxc w0 = fx_unpack(0x00007fffu);        // kiss twiddle 0: (32767, 0)
#pragma unroll
    for (int i = 0; i < 4; i++) fx_bfly2(f[2 * i], f[2 * i + 1], w0);
#pragma unroll
    for (int kb = 0; kb < 2; kb++) {
        fxc g[4] = {f[kb], f[kb + 2], f[kb + 4], f[kb + 6]};
        fx_bfly4(g, fx_unpack(fx_tw(tw, 2, 1, kb)), fx_unpack(fx_tw(tw, 2, 2, kb)), fx_unpack(fx_tw(tw, 2, 3, kb)));
        f[kb] = g[0]; f[kb + 2] = g[1]; f[kb + 4] = g[2]; f[kb + 6] = g[3];
    }
    v[0] = make_uint4(fx_pack(f[0]), fx_pack(f[1]), fx_pack(f[2]), fx_pack(f[3]));
    v[1] = make_uint4(fx_pack(f[4]), fx_pack(f[5]), fx_pack(f[6]), fx_pack(f[7]));
}

// radix 4 (m = MA) then radix 4 (m = 4 MA): points k + MA j, j < 16, of one 16 MA block
template <int N, int MA>
__device__ __forceinline__ void fx_pass_44(uint32_t *buf, const uint32_t *tw, int task)
{
    const int blk = task / MA, k = task % MA;
    const int base = blk * 16 * MA + k;
    fxc f[16];
#pragma unroll
    for (int j = 0; j < 16; j++) f[j] = fx_unpack(buf[fxpad(base + j * MA)]);
    {
        const fxc w1 = fx_unpack(fx_tw(tw, MA, 1, k)), w2 = fx_unpack(fx_tw(tw, MA, 2, k)), w3 = fx_unpack(fx_tw(tw, MA, 3, k));
#pragma unroll
        for (int i = 0; i < 4; i++) {
            fxc g[4] = {f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]};
            fx_bfly4(g, w1, w2, w3);
            f[4 * i] = g[0]; f[4 * i + 1] = g[1]; f[4 * i + 2] = g[2]; f[4 * i + 3] = g[3];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int kb = k + i * MA;
        fxc g[4] = {f[i], f[i + 4], f[i + 8], f[i + 12]};
        fx_bfly4(g, fx_unpack(fx_tw(tw, 4 * MA, 1, kb)), fx_unpack(fx_tw(tw, 4 * MA, 2, kb)), fx_unpack(fx_tw(tw, 4 * MA, 3, kb)));
        f[i] = g[0]; f[i + 4] = g[1]; f[i + 8] = g[2]; f[i + 12] = g[3];
    }
#pragma unroll
    for (int j = 0; j < 16; j++) buf[fxpad(base + j * MA)] = fx_pack(f[j]);
}

// one radix 4 stage (m = M)
template <int N, int M>
__device__ __forceinline__ void fx_pass_4(uint32_t *buf, const uint32_t *tw, int b)
{
    const int blk = b / M, k = b % M;
    const int base = blk * 4 * M + k;
    fxc g[4];
#pragma unroll
    for (int q = 0; q < 4; q++) g[q] = fx_unpack(buf[fxpad(base + q * M)]);
    fx_bfly4(g, fx_unpack(fx_tw(tw, M, 1, k)), fx_unpack(fx_tw(tw, M, 2, k)), fx_unpack(fx_tw(tw, M, 3, k)));
#pragma unroll
    for (int q = 0; q < 4; q++) buf[fxpad(base + q * M)] = fx_pack(g[q]);
}

// The mode's transform(s): G = 2048 / N symbols side by side; kf_factor gives
//   2048: 4 4 4 4 4 2    1024: 4 4 4 4 4    512: 4 4 4 4 2    256: 4 4 4 4     (executed right to left)
template <int N>
__device__ __forceinline__ void fx_ifft(uint32_t *buf, const uint32_t *tw, int tid)
{
    if (N == 2048 || N == 512) {
#pragma unroll
        for (int i = 0; i < FX_POINTS / 8 / FX_THREADS; i++) fx_pass_24<N>(buf, tw, tid + FX_THREADS * i);
        __syncthreads();
        fx_pass_44<N, 8>(buf, tw, tid);
        __syncthreads();
        if (N == 2048) fx_pass_44<N, 128>(buf, tw, tid);
        else {
#pragma unroll
            for (int i = 0; i < FX_POINTS / 4 / FX_THREADS; i++) fx_pass_4<N, 128>(buf, tw, tid + FX_THREADS * i);
        }
    }
    else {
        fx_pass_44<N, 1>(buf, tw, tid);
        __syncthreads();
        fx_pass_44<N, 16>(buf, tw, tid);
        if (N == 1024) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < FX_POINTS / 4 / FX_THREADS; i++) fx_pass_4<N, 256>(buf, tw, tid + FX_THREADS * i);
        }
    }
    __syncthreads();
}

// out-position of symbol s inside the TF and its length
__device__ __forceinline__ int fx_pos(const FixParams &p, int s) { return s == 0 ? 0 : p.null_size + (s - 1) * p.sym_size; }
__device__ __forceinline__ int fx_size(const FixParams &p, int s) { return s == 0 ? p.null_size : p.sym_size; }

// sample at offset o (may be negative or beyond the symbol: cyclic extension) of a symbol whose
// N samples start at x and whose cyclic prefix is `pre` long
__device__ __forceinline__ short2 fx_cyclic(const uint32_t *buf, int x0, int N, int pre, int o)
{
    const uint32_t w = buf[fxpad(x0 + ((o - pre) & (N - 1)))];    // N is a power of two
    return make_short2((short)(w & 0xffffu), (short)(w >> 16));
}

template <int N>
__global__ void __launch_bounds__(FX_THREADS) k_symbols_fix(const __grid_constant__ FixParams p)
{
    __shared__ __align__(16) FixSmem sm;
    const int tid = threadIdx.x;
    constexpr int K = N * 3 / 4, G = FX_POINTS / N;
    constexpr int K16 = K / 16;                    // 16-carrier work items per symbol
    const int tf = blockIdx.x / p.n_chunks;
    const int chunk = blockIdx.x - tf * p.n_chunks;
    const int grp0 = chunk * p.groups_per_chunk;
    const int grp1 = min(grp0 + p.groups_per_chunk, p.n_groups);
    const uint8_t *bits = p.bits + (size_t)tf * p.tf_in_bytes;
    short2 *out = p.out + (size_t)tf * p.tf_samples;
    const bool tii_on = p.tii_count > 0 && (((p.tf_offset + tf + p.tii_parity) & 1) == 0);
    const int W = p.window;

    // ---- per-CTA tables ----
    for (int M = (N == 2048 || N == 512) ? 2 : 1; 4 * M <= N; M *= 4)          // one table per radix-4 stage (fx_tw)
        for (int i = tid; i < 3 * M; i += FX_THREADS)
            sm.tw[fx_tw_off(M) + i] = __ldg(reinterpret_cast<const uint32_t *>(p.tw) + (N / (4 * M)) * (i / M + 1) * (i % M));
    for (int b = tid; b < 256; b += FX_THREADS) {
        uint32_t s = 0;
#pragma unroll
        for (int n = 0; n < 8; n++) s |= ((b >> (7 - n)) & 1u) << (4 * n);
        sm.spread[b] = s;
    }
    if (tid < 8) {
        // fixed(1) = 16384, fixed(M_SQRT1_2) = 11585 (QpskSymbolMapper.cpp:60, PhaseReference.cpp:139-150)
        const short c[8] = {16384, 11585, 0, -11585, -16384, -11585, 0, 11585};
        sm.c8[tid] = make_short2(c[tid], c[(tid + 6) & 7]);
    }
    for (int i = tid; i < 2 * W; i += FX_THREADS) {
        sm.win[i] = __ldg(p.window_tab + i);
        sm.tail[i] = make_short2(0, 0);
    }

    // ---- carrier work item of this thread: symbol cg of the group, carriers 16*jj .. 16*jj+15 ----
    const int cg = tid / K16, jj = tid - cg * K16;
    const bool carrier_thread = tid < G * K16;
    uint32_t ph_lo = 0, ph_hi = 0;             // nibble-packed running phase, carriers 0-7 / 8-15
    if (carrier_thread) {
#pragma unroll
        for (int n = 0; n < 8; n++) {
            ph_lo |= (uint32_t)__ldg(p.phase0 + 16 * jj + n) << (4 * n);
            ph_hi |= (uint32_t)__ldg(p.phase0 + 16 * jj + 8 + n) << (4 * n);
        }
    }
    __syncthreads();
    // phase prefix: data symbol d = s - 2 for s >= 2; consume the rows before the chunk
    if (carrier_thread) {
        const int nd = max(0, grp0 * G - 2);
        const uint8_t *row = bits + 2 * jj;
        for (int d = 0; d < nd; d++, row += K / 4) {
            const unsigned iw = __ldg(reinterpret_cast<const unsigned short *>(row));
            const unsigned qw = __ldg(reinterpret_cast<const unsigned short *>(row + K / 8));
            ph_lo = (ph_lo + 0x11111111u + 2u * sm.spread[(iw ^ qw) & 0xff] + 4u * sm.spread[qw & 0xff]) & 0x77777777u;
            ph_hi = (ph_hi + 0x11111111u + 2u * sm.spread[((iw ^ qw) >> 8) & 0xff] + 4u * sm.spread[qw >> 8]) & 0x77777777u;
        }
    }

    // bit rows one group ahead, digit-reversed carrier positions once per CTA (as in k_symbols)
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
    uint32_t posp[8];
#pragma unroll
    for (int g = 0; g < G; g++) { rows_i[g] = 0; rows_q[g] = 0; }
#pragma unroll
    for (int n = 0; n < 8; n++) posp[n] = 0;
    if (carrier_thread) {
        fetch_rows(grp0, rows_i, rows_q);
#pragma unroll
        for (int n = 0; n < 8; n++) posp[n] = __ldg(reinterpret_cast<const uint32_t *>(p.pos_of_src + 16 * jj) + n);
    }

    for (int grp = grp0; grp < grp1; grp++) {
        const int s0 = grp * G;
        unsigned cur_i[G], cur_q[G];
#pragma unroll
        for (int g = 0; g < G; g++) { cur_i[g] = rows_i[g]; cur_q[g] = rows_q[g]; }
        if (carrier_thread && grp + 1 < grp1) fetch_rows(grp + 1, rows_i, rows_q);
        // ---- 1. carriers into their digit-reversed positions; everything else is zero ----
        for (int i = tid; i < FX_BUF; i += FX_THREADS) sm.buf[i] = 0u;
        __syncthreads();
        if (carrier_thread) {
            uint32_t my_lo = 0, my_hi = 0;
            bool mine = false;
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int s = s0 + g;
                if (s >= 2 && s <= p.L) {
                    const unsigned iw = cur_i[g], qw = cur_q[g];
                    ph_lo = (ph_lo + 0x11111111u + 2u * sm.spread[(iw ^ qw) & 0xff] + 4u * sm.spread[qw & 0xff]) & 0x77777777u;
                    ph_hi = (ph_hi + 0x11111111u + 2u * sm.spread[((iw ^ qw) >> 8) & 0xff] + 4u * sm.spread[qw >> 8]) & 0x77777777u;
                }
                if (g == cg) { my_lo = ph_lo; my_hi = ph_hi; mine = s >= 1 && s <= p.L; }
            }
            if (mine) {
#pragma unroll
                for (int n = 0; n < 16; n++) {
                    const unsigned ph = ((n < 8 ? my_lo >> (4 * n) : my_hi >> (4 * (n - 8)))) & 7u;
                    const short2 c = sm.c8[ph];
                    const int pos = (int)((n & 1) ? posp[n >> 1] >> 16 : posp[n >> 1] & 0xffffu);
                    sm.buf[fxpad(cg * N + pos)] = (uint32_t)(uint16_t)c.x | ((uint32_t)(uint16_t)c.y << 16);
                }
            }
        }
        if (grp == 0 && tii_on && tid < p.tii_count) {
            const short2 c = __ldg(p.tii_val + tid);
            sm.buf[fxpad(__ldg(p.tii_pos + tid))] = (uint32_t)(uint16_t)c.x | ((uint32_t)(uint16_t)c.y << 16);
        }
        __syncthreads();

        // ---- 2. KISS inverse FFT, in place, smallest sub-transforms first ----
        fx_ifft<N>(sm.buf, sm.tw, tid);

        // ---- 3. guard interval (+ window) and store ----
        for (int g = 0; g < G; g++) {
            const int s = s0 + g;
            if (s > p.L) break;
            const int x0 = g * N;
            const int size = fx_size(p, s), pre = size - N, pos = fx_pos(p, s);
            if (W == 0) {
                // plain guard interval: cyclic prefix + body, VEC samples per access (the symbol sizes of the
                // mode are multiples of VEC, so a vector never wraps and stays aligned; fxpad keeps runs of 8)
                constexpr int VEC = (N == 2048 || N == 1024) ? 4 : N == 512 ? 2 : 1;
                for (int o = VEC * tid; o < size; o += VEC * FX_THREADS) {
                    const uint32_t *src = sm.buf + fxpad(x0 + ((o - pre) & (N - 1)));
                    uint32_t *dst = reinterpret_cast<uint32_t *>(out + pos + o);
                    if (VEC == 4) *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src);
                    else if (VEC == 2) *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(src);
                    else *dst = *src;
                }
                continue;
            }
            const bool first = s == 0, last = s == p.L;
            const int lo = (W > 0 && !first) ? -W : 0;
            const int hi = (W > 0 && !last) ? size - W : size;
            for (int o = lo + tid; o < hi; o += FX_THREADS) {
                short2 v = fx_cyclic(sm.buf, x0, N, pre, o);
                if (W > 0 && !first && o < W) {
                    const short wr = sm.win[o + W];
                    v = make_short2(fx_mul(v.x, wr), fx_mul(v.y, wr));
                    short2 f;                                      // falling edge of symbol s - 1 at the same place
                    if (g > 0) {
                        const int psize = fx_size(p, s - 1);
                        const short wf = sm.win[W - 1 - o];       // = w[2W - 1 - (o' - (psize - W))], o' = psize + o
                        const short2 u = fx_cyclic(sm.buf, x0 - N, N, psize - N, psize + o);
                        f = make_short2(fx_mul(u.x, wf), fx_mul(u.y, wf));
                    }
                    else f = sm.tail[o + W];
                    v = make_short2((short)(v.x + f.x), (short)(v.y + f.y));
                }
                out[pos + o] = v;
            }
        }
        if (W > 0) {
            // falling edge of the group's last symbol for the next group
            const int s = min(s0 + G - 1, p.L);
            const int x0 = (s - s0) * N;
            const int size = fx_size(p, s), pre = size - N;
            __syncthreads();                                       // sm.tail was read above
            for (int i = tid; i < 2 * W; i += FX_THREADS) {
                const short2 u = fx_cyclic(sm.buf, x0, N, pre, size - W + i);
                const short wf = sm.win[2 * W - 1 - i];
                sm.tail[i] = make_short2(fx_mul(u.x, wf), fx_mul(u.y, wf));
            }
        }
        __syncthreads();
    }
}

} // namespace dabmod
