// Resampler kernels.  Reference: src/Resampler.cpp:51-112 (geometry, window,
// scale factor), :131-195 (process).
//
// The reference is a Hann-windowed 50 %-overlap FFT resampler: per hop of
// hi = Ni/2 new input samples it transforms the block x_b = [H_{b-1} | H_b] * w
// with a forward FFT of size Ni, re-lays the spectrum out on No bins, scales,
// transforms back with an inverse FFT of size No and emits
//     out_b[n] = y_{b-1}[ho + n] + y_b[n],        n in [0, ho), ho = No/2
// keeping H_b and y_b[ho:] as state for the next hop.
//
// Formulation used here (no sequential state, every hop independent):
// a delay by half a block is a factor (-1)^k on the spectrum, on both the Ni
// and the No grid (Ni, No, hi even), and all steps are linear, so
//     out_b[n] = IFFT_No( relayout( FFT_Ni(c_b) ) * factor )[n],   n in [0, ho)
//     c_b[m]      = (w[m] + w[m + hi]) * H_{b-1}[m]                 m in [0, hi)
//     c_b[hi + m] =  w[hi + m] * H_b[m] + w[m] * H_{b-2}[m]
// H_b = in[b*hi : (b+1)*hi], samples before the stream start are zero, samples
// before the start of this launch come from the `hist` buffer (the last 2*hi
// input samples of the previous launch of the same stream).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fft.cuh"
#include "kernels.cuh"

namespace dabmod {

constexpr int RES_MAX_PASSES = 24;

struct ResParams {
    int ni, no;                    // FFT sizes
    long long total_hops;          // hops in this launch (n_tf * tf_samples / hi)
    float factor;                  // Resampler::myFactor
    const float2 *in;              // total_hops * hi samples
    const float2 *hist;            // 2 * hi samples that precede in[0]
    const float *win;              // ni floats (Resampler::myWindow)
    const float2 *tw_in;           // ni entries e^{+j 2 pi k / ni}
    const float2 *tw_out;          // no entries
    int n_rad_in, n_rad_out;
    unsigned char rad_in[RES_MAX_PASSES], rad_out[RES_MAX_PASSES];
    float2 *scratch;               // gridDim.x * 2 * max(ni, no) entries, or nullptr = shared memory
    void *out;                     // total_hops * ho samples
    PostParams post;
    int dbg;                       // profiling aid ("res_dbg", results are WRONG when non-zero): 1 no input loads,
                                   // 2 no stores, 4 no inverse transforms, 8 no forward transform
};

__device__ __forceinline__ float2 res_load(const ResParams &p, long long s)
{
    return s >= 0 ? __ldg(p.in + s) : __ldg(p.hist + (s + p.ni));
}

// c_b into dst[0, ni)
__device__ __forceinline__ void res_build_block(const ResParams &p, long long hop, float2 *dst, int tid, int nth)
{
    const int hi = p.ni / 2;
    const long long base = hop * hi;
    for (int m = tid; m < hi; m += nth) {
        const float w0 = __ldg(p.win + m), w1 = __ldg(p.win + hi + m);
        const float2 a = res_load(p, base - hi + m);      // H_{b-1}[m]
        const float2 b = res_load(p, base + m);           // H_b[m]
        const float2 c = res_load(p, base - 2 * hi + m);  // H_{b-2}[m]
        dst[m] = make_float2(__fadd_rn(__fmul_rn(w0, a.x), __fmul_rn(w1, a.x)),
                             __fadd_rn(__fmul_rn(w0, a.y), __fmul_rn(w1, a.y)));
        dst[hi + m] = make_float2(__fadd_rn(__fmul_rn(w1, b.x), __fmul_rn(w0, c.x)),
                                  __fadd_rn(__fmul_rn(w1, b.y), __fmul_rn(w0, c.y)));
    }
}

// Spectrum re-layout + scale (Resampler.cpp:153-181): bin k of the No grid from F (Ni grid)
__device__ __forceinline__ float2 res_relayout(const ResParams &p, const float2 *F, int k)
{
    const int ni = p.ni, no = p.no, hi = ni / 2, ho = no / 2;
    float2 v = make_float2(0.f, 0.f);
    if (no > ni) {
        if (k <= hi) v = F[k];                       // k == hi: the input Nyquist bin, copied to both sides
        else if (k >= no - hi) v = F[k - (no - ni)];
    }
    else {
        if (k < ho) v = F[k];
        else v = F[ni - no + k];
        if (k == ho) {                               // average of the two input bins that fold here
            const float2 u = F[ho];
            v = make_float2((v.x + u.x) * 0.5f, (v.y + u.y) * 0.5f);
        }
    }
    return make_float2(v.x * p.factor, v.y * p.factor);
}

// Small DFTs of odd prime size, O(R^2), constants in double precision
template <int R>
__device__ __forceinline__ float2 prime_root(int e)
{
    // e^{+j 2 pi e / R}, e in [0, R)
    if (R == 3) {
        const float c[3] = {1.f, -0.5f, -0.5f};
        const float s[3] = {0.f, 0.86602540378443864676f, -0.86602540378443864676f};
        return make_float2(c[e], s[e]);
    }
    if (R == 5) {
        const float c[5] = {1.f, 0.30901699437494742410f, -0.80901699437494742410f, -0.80901699437494742410f,
                            0.30901699437494742410f};
        const float s[5] = {0.f, 0.95105651629515357212f, 0.58778525229247312917f, -0.58778525229247312917f,
                            -0.95105651629515357212f};
        return make_float2(c[e], s[e]);
    }
    const float c[7] = {1.f, 0.62348980185873353053f, -0.22252093395631440429f, -0.90096886790241912624f,
                        -0.90096886790241912624f, -0.22252093395631440429f, 0.62348980185873353053f};
    const float s[7] = {0.f, 0.78183148246802980871f, 0.97492791218182360702f, 0.43388373911755812048f,
                        -0.43388373911755812048f, -0.97492791218182360702f, -0.78183148246802980871f};
    return make_float2(c[e], s[e]);
}

template <int R, bool INV>
__device__ __forceinline__ void dft_prime(float2 *v)
{
    float2 o[R];
#pragma unroll
    for (int m = 0; m < R; m++) {
        float2 acc = v[0];
#pragma unroll
        for (int r = 1; r < R; r++) {
            const float2 w = tw_dir<INV>(prime_root<R>((r * m) % R));
            acc.x = fmaf(v[r].x, w.x, fmaf(-v[r].y, w.y, acc.x));
            acc.y = fmaf(v[r].x, w.y, fmaf(v[r].y, w.x, acc.y));
        }
        o[m] = acc;
    }
#pragma unroll
    for (int m = 0; m < R; m++) v[m] = o[m];
}

template <int R, bool INV>
__device__ __forceinline__ void butterfly_any(float2 *v)
{
    if (R == 2) fft2<INV>(v[0], v[1]);
    else if (R == 4) fft4<INV>(v[0], v[1], v[2], v[3]);
    else if (R == 8) fft8<INV>(v);
    else if (R == 16) fft16<INV>(v);
    else dft_prime<(R == 3 || R == 5 || R == 7) ? R : 3, INV>(v);
}

// One Stockham pass src -> dst (distinct buffers, any address space), runtime N / Ns.
template <int R, bool INV>
__device__ __forceinline__ void generic_pass(const float2 *src, float2 *dst, int N, int Ns, const float2 *tw,
                                             int tid, int nth)
{
    const int nb = N / R;
    const int step = N / (Ns * R);
    for (int b = tid; b < nb; b += nth) {
        const int k = b % Ns;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = src[b + r * nb];
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; r++) v[r] = cmul(v[r], tw_dir<INV>(__ldg(tw + k * r * step)));
        }
        butterfly_any<R, INV>(v);
        const int j0 = (b - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; r++) dst[j0 + r * Ns] = v[r];
    }
}

// A pass of any prime radix R (11, 13, ...: output rates whose L carries such a factor, e.g. 2.816 Msps = 11/8):
// one output per loop iteration, O(R) terms each, the inter-pass twiddle and the R-th root of unity folded into
// one look-up in the N-entry root table.  Completeness, not speed.
template <bool INV>
__device__ __forceinline__ void generic_pass_prime(const float2 *src, float2 *dst, int N, int Ns, int R, const float2 *tw,
                                                   int tid, int nth)
{
    const int nb = N / R, step = N / (Ns * R), rstep = N / R;
    for (int idx = tid; idx < N; idx += nth) {
        const int m = idx / nb, b = idx - m * nb;
        const int k = b % Ns;
        const long long e1 = (long long)k * step + (long long)m * rstep;   // exponent per unit of r
        float ax = 0.f, ay = 0.f;
        for (int r = 0; r < R; r++) {
            const float2 x = src[b + r * nb];
            const float2 w = tw_dir<INV>(__ldg(tw + (int)((e1 * r) % N)));
            ax = fmaf(x.x, w.x, fmaf(-x.y, w.y, ax));
            ay = fmaf(x.x, w.y, fmaf(x.y, w.x, ay));
        }
        dst[(b - k) * R + k + m * Ns] = make_float2(ax, ay);
    }
}

template <bool INV>
__device__ __forceinline__ void generic_fft(float2 *&src, float2 *&dst, int N, const unsigned char *rad, int n_rad,
                                            const float2 *tw, int tid, int nth)
{
    int Ns = 1;
    for (int i = 0; i < n_rad; i++) {
        const int R = rad[i];
        switch (R) {
            case 16: generic_pass<16, INV>(src, dst, N, Ns, tw, tid, nth); break;
            case 8: generic_pass<8, INV>(src, dst, N, Ns, tw, tid, nth); break;
            case 4: generic_pass<4, INV>(src, dst, N, Ns, tw, tid, nth); break;
            case 2: generic_pass<2, INV>(src, dst, N, Ns, tw, tid, nth); break;
            case 3: generic_pass<3, INV>(src, dst, N, Ns, tw, tid, nth); break;
            case 5: generic_pass<5, INV>(src, dst, N, Ns, tw, tid, nth); break;
            case 7: generic_pass<7, INV>(src, dst, N, Ns, tw, tid, nth); break;
            default: generic_pass_prime<INV>(src, dst, N, Ns, R, tw, tid, nth); break;
        }
        Ns *= R;
        __syncthreads();
        float2 *t = src; src = dst; dst = t;
    }
}

// ---------------------------------------------------------------------------
// k_resample_generic: any Ni / No whose prime factors are at most 251 (2, 3, 5, 7 as register butterflies).
// Persistent CTAs stride over the hops; two ping-pong buffers of max(Ni, No)
// points live in shared memory when they fit, else in a per-CTA slice of
// `scratch` (L2-resident).  This is the completeness path; the TM I hot
// configurations have dedicated kernels below.
// ---------------------------------------------------------------------------
constexpr int RESG_THREADS = 256;

template <bool POST>
__global__ void __launch_bounds__(RESG_THREADS) k_resample_generic(const __grid_constant__ ResParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int nmax = p.ni > p.no ? p.ni : p.no;
    float2 *A = p.scratch ? p.scratch + (size_t)blockIdx.x * 2 * nmax : reinterpret_cast<float2 *>(smem_raw);
    float2 *B = A + nmax;
    const int ho = p.no / 2;
    unsigned clip = 0;
    for (long long hop = blockIdx.x; hop < p.total_hops; hop += gridDim.x) {
        float2 *src = A, *dst = B;
        res_build_block(p, hop, src, tid, RESG_THREADS);
        __syncthreads();
        generic_fft<false>(src, dst, p.ni, p.rad_in, p.n_rad_in, p.tw_in, tid, RESG_THREADS);
        for (int k = tid; k < p.no; k += RESG_THREADS) dst[k] = res_relayout(p, src, k);
        __syncthreads();
        { float2 *t = src; src = dst; dst = t; }
        generic_fft<true>(src, dst, p.no, p.rad_out, p.n_rad_out, p.tw_out, tid, RESG_THREADS);
        const size_t obase = (size_t)hop * ho;
        for (int n = tid; n < ho; n += RESG_THREADS) store_sample<POST>(p.out, obase + n, src[n], p.post, clip);
        __syncthreads();
    }
    if (POST && p.post.format != 0) flush_clip(p.post, clip);
}

} // namespace dabmod
