// Host-side construction of the constant tables the kernels consume.
// Each function cites the reference code whose behaviour it reproduces; the
// implementations are written for the kernels' layouts (per *source* carrier,
// FFT-bin addressed), not translated from the reference.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace dabmod {

// Transmission-mode constants: DabModulator.cpp:84-122 (symbols, carriers,
// spacing, null/symbol sizes), FrequencyInterleaver.cpp:41-66 (beta),
// BlockPartitioner.cpp:44-73 (bytes per TF).
struct ModeInfo {
    int mode;
    int L;          // symbols per TF incl. phase reference, excl. null
    int K;          // carriers
    int N;          // FFT size
    int null_size;
    int sym_size;
    int beta;
    int tf_in_bytes;
    int tf_samples;
};

inline ModeInfo mode_info(int mode)
{
    ModeInfo m{};
    switch (mode) {
        case 0:
        case 1: m = {1, 76, 1536, 2048, 2656, 2552, 511, 0, 0}; break;
        case 2: m = {2, 76, 384, 512, 664, 638, 127, 0, 0}; break;
        case 3: m = {3, 153, 192, 256, 345, 319, 63, 0, 0}; break;
        case 4: m = {4, 76, 768, 1024, 1328, 1276, 255, 0, 0}; break;
        default: throw std::runtime_error("DabModulator::setMode invalid mode size");
    }
    m.tf_in_bytes = (m.L - 1) * m.K / 4;
    m.tf_samples = m.null_size + m.L * m.sym_size;
    return m;
}

// Carrier position (0..K-1 in the reference's carrier buffers) -> FFT bin.
// OfdmGenerator.cpp:77-94: first half of the buffer = positive frequencies
// from bin 1, second half = negative frequencies ending at bin N-1.
inline int bin_of_carrier(const ModeInfo &m, int c)
{
    return c < m.K / 2 ? c + 1 : c + m.N - m.K;
}

// Frequency interleaver (FrequencyInterleaver.cpp:73-92): source symbol j of a
// QPSK block lands on carrier position dest[j].
inline std::vector<int> interleaver_dest(const ModeInfo &m)
{
    std::vector<int> dest;
    dest.reserve(m.K);
    const int N = m.N, K = m.K, lo = (N - K) / 2, hi = N - lo;
    int perm = 0;
    for (int j = 1; j < N; j++) {
        perm = (13 * perm + m.beta) & (N - 1);
        if (perm < lo || perm > hi || perm == N / 2) continue;
        dest.push_back(perm > N / 2 ? perm - (1 + N / 2) : perm + (K - N / 2));
    }
    if ((int)dest.size() != K) throw std::logic_error("interleaver size");
    return dest;
}

// Phase reference symbol (PhaseReference.cpp:35-44,91-124,152-171; EN 300 401
// clause 14.3.2): quarter-turn index per carrier position, value = j^q.
inline std::vector<uint8_t> phase_ref_quarter_turns(const ModeInfo &m)
{
    static const uint8_t h[4][16] = {   // table 43, each row repeats after 16
        {0, 2, 0, 0, 0, 0, 1, 1, 2, 0, 0, 0, 2, 2, 1, 1},
        {0, 3, 2, 3, 0, 1, 3, 0, 2, 1, 2, 3, 2, 3, 3, 0},
        {0, 0, 0, 2, 0, 2, 1, 3, 2, 2, 0, 2, 2, 0, 1, 3},
        {0, 1, 2, 1, 0, 3, 3, 2, 2, 3, 2, 1, 2, 1, 3, 2},
    };
    // {i, n} per block of 32 carriers in buffer order, packed as i*4+n
    static const uint8_t tm1[48] = {
        3, 13, 9, 5, 2, 14, 9, 4, 2, 14, 11, 7, 0, 14, 9, 7, 3, 15, 11, 4, 3, 12, 9, 5,
        1, 6, 8, 13, 3, 6, 10, 15, 2, 5, 10, 15, 1, 6, 11, 15, 2, 6, 10, 13, 1, 7, 9, 14};
    static const uint8_t tm2[12] = {8, 6, 2, 13, 8, 7, 2, 7, 10, 14, 1, 6};
    static const uint8_t tm3[6] = {14, 10, 6, 2, 7, 8};
    static const uint8_t tm4[24] = {0, 13, 8, 6, 0, 13, 10, 6, 2, 13, 11, 4,
                                    0, 5, 9, 14, 2, 6, 8, 15, 3, 5, 11, 14};
    const uint8_t *tab = m.mode == 1 ? tm1 : m.mode == 2 ? tm2 : m.mode == 3 ? tm3 : tm4;
    std::vector<uint8_t> q(m.K);
    for (int blk = 0; blk < m.K / 32; blk++)
        for (int k = 0; k < 32; k++)
            q[blk * 32 + k] = (h[tab[blk] >> 2][k & 15] + (tab[blk] & 3)) & 3;
    return q;
}

// TII carrier pairs (TII.cpp:229-337; EN 300 401 clause 14.8). Returns the
// carrier positions ix where a pair (ix, ix+1) is switched on; empty + false
// when the mode has no TII (TII ctor throws for TM III/IV, DabModulator.cpp:178-190).
inline bool tii_pairs(const ModeInfo &m, int comb, int pattern, std::vector<int> &pairs)
{
    pairs.clear();
    if (m.mode != 1 && m.mode != 2) return false;
    if (pattern < 0 || pattern > 69) throw std::runtime_error("TII::TII pattern not valid!");
    if (comb < 0 || comb > 23) throw std::runtime_error("TII::TII comb not valid!");
    // pattern table = the 70 bytes of weight 4 in ascending order, b0 = MSB
    int word = 0;
    for (int w = 0, n = -1; w < 256; w++)
        if (__builtin_popcount(w) == 4 && ++n == pattern) { word = w; break; }
    auto enable = [&](int k) {
        const int ix = m.K / 2 + k - (k >= 0 ? 1 : 0);
        if (ix < 0 || ix + 1 >= m.K) throw std::runtime_error("TII::enable_carrier invalid k!");
        pairs.push_back(ix);
    };
    for (int b = 0; b < 8; b++) {
        if (!((word >> (7 - b)) & 1)) continue;
        if (m.mode == 1) {
            for (int base : {-768, -384, 1, 385}) enable(base + 2 * comb + 48 * b);
        }
        else {
            enable((b < 4 ? -192 : -191) + 2 * comb + 48 * b);
        }
    }
    return true;
}

// CicEqualizer coefficients per carrier position (CicEqualizer.cpp:29-57),
// float32 arithmetic in the reference's order.
// The reference's constructor takes `size_t spacing` and DabModulator.cpp:172-175 passes a float, so a
// fractional N * rate / 2048000 (e.g. TM III at 2.5 Msps: 312.5) is truncated before the angles are computed.
inline std::vector<float> cic_filter(int K, float spacing_f, int R)
{
    const size_t spacing = (size_t)spacing_f;
    std::vector<float> f(K);
    const float pi = 4.0f * atanf(1.0f);
    for (int i = 0; i < K; i++) {
        const int k = i < (K + 1) / 2 ? i + ((K & 1) ^ 1) : i - K;
        if (k == 0) { f[i] = 1.0f; continue; }
        const float angle = pi * k / spacing;
        float v = sinf(angle / R) / sinf(angle);
        v = fabsf(v) * R;
        f[i] = powf(v, 4);
    }
    return f;
}

// Whether DabModulator inserts a CicEqualizer, and its ratio (DabModulator.cpp:154-176).
inline bool cic_enabled(uint64_t clock_rate, uint64_t output_rate, unsigned &ratio)
{
    ratio = 1;
    if (!clock_rate) return false;
    ratio = (unsigned)(clock_rate / output_rate) / 4;
    if (clock_rate == 400000000) return (ratio & 1) != 0;
    return true;
}

// Raised-cosine rising edge of the OFDM window (GuardIntervalInserter.cpp:96-113)
inline std::vector<float> guard_window(int overlap)
{
    std::vector<float> w(2 * (size_t)overlap);
    for (size_t i = 0; i < w.size(); i++)
        w[i] = (float)(0.5 * (1.0 - cos(M_PI * (double)i / (double)(2 * overlap - 1))));
    return w;
}

// Built-in FIR taps (FIRFilter.cpp:50-71): the 45-tap symmetric low-pass that
// doc/fir-filter/generate-filter.py produces for fs=2.048e6, cutoff 810e3,
// transition 250e3.  Stored as the first half + centre tap.
inline std::vector<float> default_fir_taps()
{
    static const float half[23] = {
        -0.00110450468492f, 0.00120703084394f, -0.000840645749122f, -0.000187368263141f,
        0.00184351124335f, -0.00355578539893f, 0.00419321097434f, -0.00254214904271f,
        -0.00183473504148f, 0.00781436730176f, -0.0125957569107f, 0.0126200336963f,
        -0.00537294941023f, -0.00866683479398f, 0.0249746385962f, -0.0356550291181f,
        0.0319730602205f, -0.00795613788068f, -0.0363943465054f, 0.0938014090061f,
        -0.151176810265f, 0.193567320704f, 0.791776955128f};
    std::vector<float> t(45);
    for (int i = 0; i < 23; i++) t[i] = t[44 - i] = half[i];
    return t;
}

// e^{+j 2 pi k / n} evaluated in double
inline void twiddle_table(int n, std::vector<float> &re_im)
{
    re_im.resize(2 * (size_t)n);
    for (int k = 0; k < n; k++) {
        const double a = 2.0 * M_PI * (double)k / (double)n;
        re_im[2 * k] = (float)cos(a);
        re_im[2 * k + 1] = (float)sin(a);
    }
}

// Twiddles of the OFDM IFFT passes in the layout StockhamPass reads:
// for pass (R, Ns) entry (r-1)*Ns + k = e^{+j 2 pi k r / (Ns R)}; first pass has none.
// Radix plans: N=2048: 16,16,8; 1024: 16,8,8; 512: 8,8,8; 256: 16,16 (kernels.cuh).
inline void symbol_fft_twiddles(int N, std::vector<float> &re_im)
{
    std::vector<int> rad;
    switch (N) {
        case 2048: rad = {16, 16, 8}; break;
        case 1024: rad = {16, 8, 8}; break;
        case 512: rad = {8, 8, 8}; break;
        default: rad = {16, 16}; break;
    }
    re_im.clear();
    int Ns = rad[0];
    for (size_t i = 1; i < rad.size(); i++) {
        const int R = rad[i];
        for (int r = 1; r < R; r++)
            for (int k = 0; k < Ns; k++) {
                const double a = 2.0 * M_PI * (double)k * (double)r / ((double)Ns * (double)R);
                re_im.push_back((float)cos(a));
                re_im.push_back((float)sin(a));
            }
        Ns *= R;
    }
}

// Resampler geometry (Resampler.cpp:51-83): L/M after reduction by the gcd, FFT
// sizes from the `resolution` (= OFDM spacing, DabModulator.cpp:265-268) and the
// float32 scale factor applied between the two transforms.
struct ResamplerPlan {
    uint64_t L, M;
    int ni, no;
    float factor;
};

inline ResamplerPlan resampler_plan(uint64_t in_rate, uint64_t out_rate, int resolution)
{
    uint64_t a = in_rate, b = out_rate;
    while (b) { const uint64_t t = a % b; a = b; b = t; }
    ResamplerPlan p{};
    p.L = out_rate / a;
    p.M = in_rate / a;
    uint64_t factor = (uint64_t)resolution * 2 / p.M;
    if (factor & 1) ++factor;
    p.ni = (int)(factor * p.M);
    p.no = (int)(factor * p.L);
    const int big = p.ni > p.no ? p.ni : p.no;
    p.factor = 1.0f / big * out_rate / in_rate;
    return p;
}

// Hann window of the resampler (Resampler.cpp:85-92); note the (len - 1) divisor.
inline std::vector<float> resampler_window(int ni)
{
    std::vector<float> w(ni);
    for (int i = 0; i < ni; i++) w[i] = (float)(0.5 * (1.0 - cos(2.0 * M_PI * i / (ni - 1))));
    return w;
}

} // namespace dabmod
