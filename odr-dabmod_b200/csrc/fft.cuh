// Register-resident radix-4/8/16 butterflies and the shared-memory Stockham
// pass used by every FFT in this library (OFDM IFFT, CFR FFT, resampler FFTs).
//
// Conventions: unnormalised transforms like FFTW/KISS (reference
// OfdmGenerator.cpp:109-111, Resampler.cpp:94-108).  INV=true evaluates
// sum_k X[k] e^{+j 2 pi k n / N} (FFTW_BACKWARD), INV=false the forward sign.
#pragma once
#include <cuda_runtime.h>

// The pure butterfly arithmetic can also be compiled for the host (tests/test_fft_host.py
// checks it on the CPU); kernels see plain __device__ functions.
#ifndef DABMOD_FN
#define DABMOD_FN __device__ __forceinline__
#endif

namespace dabmod {

// Packed FP32 (Blackwell FADD2 / FMUL2 / FFMA2): one instruction per complex add, same rounding as two scalar
// operations.  The butterflies are bound by instruction issue and instruction supply, not by the FP32 pipe, so the
// instruction count is what counts -- and the packed instructions take their operands through free modifiers
// (cuobjdump -sass of the intrinsics below):
//     R.F32x2.LO_HI      the two halves swapped             FADD2 R8, R2.F32x2.HI_LO, -R4.F32x2.LO_HI.NP
//     -R....NP           ONE half negated                   = a + j b in one instruction
//     R.F32 / immediate  one 32-bit value for both halves   FMUL2 R8, R2.F32x2.HI_LO, R4.F32
// so a multiplication by +-j costs nothing (it folds into the consumer's operand), a +- j b is one FADD2, and a full
// complex product is FMUL2 + FFMA2: two instructions instead of four, no moves.  (Round 1 kept everything that swaps
// re and im scalar, on the assumption that a swapped register pair costs real moves.)
#if defined(__CUDA_ARCH__)
DABMOD_FN float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
DABMOD_FN float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
DABMOD_FN float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
DABMOD_FN float2 cfma(float s, float2 a, float2 c) { return __ffma2_rn(make_float2(s, s), a, c); }
// (a.x b.x - a.y b.y, a.y b.x + a.x b.y) = a * (b.x, b.x) + (-a.y, a.x) * (b.y, b.y)
DABMOD_FN float2 cmul(float2 a, float2 b)
{
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
#else
DABMOD_FN float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DABMOD_FN float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
DABMOD_FN float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
DABMOD_FN float2 cfma(float s, float2 a, float2 c) { return make_float2(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y)); }
DABMOD_FN float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.x, b.y, a.y * b.x));     // the device's rounding
}
#endif
// multiply by +j (INV) or -j (forward): free when the consumer is a packed instruction
template <bool INV>
DABMOD_FN float2 mul_j(float2 a)
{
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// a + (+-j) b and a - (+-j) b: one packed add each
template <bool INV>
DABMOD_FN float2 cadd_j(float2 a, float2 b)
{
    return cadd(a, mul_j<INV>(b));
}
template <bool INV>
DABMOD_FN float2 csub_j(float2 a, float2 b)
{
    return cadd(a, mul_j<!INV>(b));
}
// twiddle table holds e^{+j theta}; the forward transform needs the conjugate
template <bool INV>
DABMOD_FN float2 tw_dir(float2 w)
{
    return INV ? w : make_float2(w.x, -w.y);
}

template <bool INV>
DABMOD_FN void fft2(float2 &a, float2 &b)
{
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <bool INV>
DABMOD_FN void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
    const float2 t2 = cadd(a1, a3), d = csub(a1, a3);
    a0 = cadd(t0, t2);
    a1 = cadd_j<INV>(t1, d);
    a2 = csub(t0, t2);
    a3 = csub_j<INV>(t1, d);
}

// a * e^{+-j pi O / 4}, O odd
template <int O, bool INV>
DABMOD_FN float2 mul_w8(float2 a)
{
    const float h = 0.70710678118654752440f;
    // (1 + j) a = (a.x - a.y, a.x + a.y);  (1 - j) a = (a.x + a.y, a.y - a.x): one packed add each
    const float2 p = cadd(a, make_float2(-a.y, a.x)), m = cadd(a, make_float2(a.y, -a.x));
    if (O == 1) return cscale(INV ? p : m, h);
    if (O == 3) return cscale(INV ? make_float2(-m.x, -m.y) : make_float2(-p.x, -p.y), h);
    if (O == 5) return cscale(INV ? make_float2(-p.x, -p.y) : make_float2(-m.x, -m.y), h);
    return cscale(INV ? m : p, h);
}

// v[0..7] natural order in, natural order out
template <bool INV>
DABMOD_FN void fft8(float2 *v)
{
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    fft4<INV>(e0, e1, e2, e3);
    fft4<INV>(o0, o1, o2, o3);
    // W8^1 = (1 + sj)/sqrt2, W8^2 = sj, W8^3 = (-1 + sj)/sqrt2, s = +1 (INV) / -1
    o1 = mul_w8<1, INV>(o1);
    o3 = mul_w8<3, INV>(o3);
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd_j<INV>(e2, o2); v[6] = csub_j<INV>(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// v[0..15] natural order in, natural order out (4 x 4 decomposition)
template <bool INV>
DABMOD_FN void fft16(float2 *v)
{
    // cos/sin of 2 pi e / 16
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    const float h = 0.70710678118654752440f;
    // column transforms over the stride-4 subsequences
#pragma unroll
    for (int q = 0; q < 4; q++) fft4<INV>(v[q], v[q + 4], v[q + 8], v[q + 12]);
    // now v[q + 4k] = Y_q[k]; twiddle by W16^{qk}
    const float sg = INV ? 1.0f : -1.0f;
    // q=1: k=1 -> e=1, k=2 -> e=2, k=3 -> e=3
    v[1 + 4] = cmul(v[1 + 4], make_float2(c1, sg * s1));
    v[1 + 8] = cmul(v[1 + 8], make_float2(h, sg * h));
    v[1 + 12] = cmul(v[1 + 12], make_float2(s1, sg * c1));
    // q=2: e = 2, 4, 6
    v[2 + 4] = cmul(v[2 + 4], make_float2(h, sg * h));
    v[2 + 8] = mul_j<INV>(v[2 + 8]);
    v[2 + 12] = cmul(v[2 + 12], make_float2(-h, sg * h));
    // q=3: e = 3, 6, 9
    v[3 + 4] = cmul(v[3 + 4], make_float2(s1, sg * c1));
    v[3 + 8] = cmul(v[3 + 8], make_float2(-h, sg * h));
    v[3 + 12] = cmul(v[3 + 12], make_float2(-c1, -sg * s1));
    // row transforms over q for each k: X[k + 4m]
#pragma unroll
    for (int k = 0; k < 4; k++) fft4<INV>(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    // v[4k + m] = X[k + 4m] -> transpose to natural order
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int m = k + 1; m < 4; m++) {
            const float2 t = v[4 * k + m];
            v[4 * k + m] = v[4 * m + k];
            v[4 * m + k] = t;
        }
}

// ---- radix 5 / 10 / 20 (the 4000-point phase transforms of k_resample_q) ----
// 5-point DFT in place: X[k] = sum_n x[n] e^{+-j 2 pi n k / 5}
template <bool INV>
DABMOD_FN void dft5(float2 &x0, float2 &x1, float2 &x2, float2 &x3, float2 &x4)
{
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;   // cos 72, cos 144
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;    // sin 72, sin 144
    const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    const float2 a = cfma(c1, t1, cfma(c2, t2, x0));
    const float2 b = cfma(c2, t1, cfma(c1, t2, x0));
    const float2 p = cfma(s1, t3, cscale(t4, s2));
    const float2 q = cfma(s2, t3, cscale(t4, -s1));
    x0 = cadd(x0, cadd(t1, t2));
    x1 = cadd_j<INV>(a, p);
    x4 = csub_j<INV>(a, p);
    x2 = cadd_j<INV>(b, q);
    x3 = csub_j<INV>(b, q);
}

// 20 = 4 x 5 with coprime factors: the prime-factor index maps n = (5 n1 + 4 n2) mod 20,
// k = (5 k1 + 16 k2) mod 20 need no twiddles between the two stages, and every index below is a
// compile-time register number.  v[0..19] natural order in, natural order out.
template <bool INV>
DABMOD_FN void fft20(float2 *v)
{
#pragma unroll
    for (int n1 = 0; n1 < 4; n1++)
        dft5<INV>(v[(5 * n1) % 20], v[(5 * n1 + 4) % 20], v[(5 * n1 + 8) % 20], v[(5 * n1 + 12) % 20],
                  v[(5 * n1 + 16) % 20]);
#pragma unroll
    for (int k2 = 0; k2 < 5; k2++)
        fft4<INV>(v[(4 * k2) % 20], v[(5 + 4 * k2) % 20], v[(10 + 4 * k2) % 20], v[(15 + 4 * k2) % 20]);
    float2 o[20];
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
        for (int k2 = 0; k2 < 5; k2++) o[(5 * k1 + 16 * k2) % 20] = v[(5 * k1 + 4 * k2) % 20];
#pragma unroll
    for (int i = 0; i < 20; i++) v[i] = o[i];
}

// 10 = 2 x 5 the same way (n = (5 n1 + 2 n2) mod 10, k = (5 k1 + 6 k2) mod 10), but only the
// first five outputs: X[k] = Y0[k] + (-1)^k Y1[k], k < 5.  v[0..9] in, v[0..4] out.
template <bool INV>
DABMOD_FN void fft10_lo(float2 *v)
{
    dft5<INV>(v[0], v[2], v[4], v[6], v[8]);
    dft5<INV>(v[5], v[7], v[9], v[1], v[3]);
    // Y0[k2] sits at v[2 k2], Y1[k2] at v[(5 + 2 k2) % 10]
    const float2 o0 = cadd(v[0], v[5]), o1 = csub(v[2], v[7]), o2 = cadd(v[4], v[9]);
    const float2 o3 = csub(v[6], v[1]), o4 = cadd(v[8], v[3]);
    v[0] = o0; v[1] = o1; v[2] = o2; v[3] = o3; v[4] = o4;
}

template <int R, bool INV>
DABMOD_FN void fft_radix(float2 *v)
{
    static_assert(R == 2 || R == 4 || R == 8 || R == 16, "radix");
    if (R == 2) fft2<INV>(v[0], v[1]);
    if (R == 4) fft4<INV>(v[0], v[1], v[2], v[3]);
    if (R == 8) fft8<INV>(v);
    if (R == 16) fft16<INV>(v);
}

// Shared-memory index padding: one complex slot per 16 keeps the stride-R
// writes of the first Stockham pass and the stride-N/R reads of every pass
// bank-conflict free for 8-byte elements.
__device__ __host__ __forceinline__ constexpr int spad(int i) { return i + (i >> 4); }

// One Stockham radix-R pass over `nfft` independent transforms of size N that
// lie back to back in `buf` (element i of transform g at spad(g*N + i)).
//   Ns   = product of the radices of the passes already done
//   twp  = this pass's twiddles, twp[(r-1)*Ns + k] = e^{+j 2 pi k r / (Ns*R)},
//          r in [1, R), k in [0, Ns): consecutive lanes read consecutive k, so
//          the reads are bank-conflict free (unused when Ns == 1)
// The pass is split in two halves so that the caller can put ONE barrier
// between "everyone has read" and "everyone writes" (in-place operation).
// Each thread owns PER = (nfft*N/R)/NTHREADS butterflies.
template <int R, bool INV, int PER>
struct StockhamPass {
    float2 v[PER][R];

    __device__ __forceinline__ void load(const float2 *buf, int tid, int nthreads, int N, int Ns,
                                         const float2 *twp)
    {
        const int nb = N / R; // butterflies per transform
#pragma unroll
        for (int p = 0; p < PER; p++) {
            const int b = tid + p * nthreads;
            const int g = b / nb, j = b - g * nb;
            const int base = g * N + j;
            const int k = j & (Ns - 1);
#pragma unroll
            for (int r = 0; r < R; r++) v[p][r] = buf[spad(base + r * nb)];
            if (Ns > 1) {
#pragma unroll
                for (int r = 1; r < R; r++) v[p][r] = cmul(v[p][r], tw_dir<INV>(twp[(r - 1) * Ns + k]));
            }
            fft_radix<R, INV>(v[p]);
        }
    }

    __device__ __forceinline__ void store(float2 *buf, int tid, int nthreads, int N, int Ns) const
    {
        const int nb = N / R;
#pragma unroll
        for (int p = 0; p < PER; p++) {
            const int b = tid + p * nthreads;
            const int g = b / nb, j = b - g * nb;
            const int k = j & (Ns - 1);
            const int j0 = (j - k) * R + k;
            const int base = g * N + j0;
#pragma unroll
            for (int r = 0; r < R; r++) buf[spad(base + r * Ns)] = v[p][r];
        }
    }
};

} // namespace dabmod
