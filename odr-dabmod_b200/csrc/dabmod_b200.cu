// C ABI (include/dabmod_b200.h) over the sm_100a kernel family.
// Host orchestration only: tables, buffers, streams, launches, parameters.
#include "../../include/dabmod_b200.h"

#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cerrno>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "internal.h"
#include "kernels.cuh"
#include "resample.cuh"
#include "resample_up.cuh"
#include "resample_q.cuh"
#include "symbols_fixed.cuh"
#include "symbols_warp.cuh"
#include "symbols_warp_g.cuh"
#include "tables.h"

using namespace dabmod;

namespace {

thread_local std::string g_last_error;

#define CUDA_CHECK(expr)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            throw ApiError(DABMOD_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    void alloc(size_t count)
    {
        release();
        if (count == 0) return;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess)
            throw ApiError(DABMOD_B200_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        n = count;
    }
    void upload(const std::vector<T> &v, cudaStream_t s)
    {
        if (v.size() > n) alloc(v.size());
        if (!v.empty())
            CUDA_CHECK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

size_t format_bytes(int fmt)
{
    switch (fmt) {
        case DABMOD_B200_FMT_COMPLEXF: return 8;
        case DABMOD_B200_FMT_S16: return 4;
        case DABMOD_B200_FMT_U8:
        case DABMOD_B200_FMT_S8: return 2;
        default: throw ApiError(DABMOD_B200_EINVAL, "FormatConverter: Invalid format");
    }
}

// PAPRStats.cpp:36-107: peak-to-average power ratio over a sliding window of blocks (symbols)
struct PaprWindow {
    size_t n_blocks = 0;
    std::deque<double> peaks, means;
    void add(double peak, double mean)
    {
        peaks.push_back(peak);
        means.push_back(mean);
        if (means.size() > n_blocks) { means.pop_front(); peaks.pop_front(); }
    }
    double papr() const
    {
        if (means.size() < n_blocks) return 0;
        double peak = 0, rms2 = 0;
        for (size_t i = 0; i < peaks.size(); i++) {
            if (peaks[i] > peak) peak = peaks[i];
            rms2 += means[i];
        }
        rms2 /= (double)peaks.size();
        return 10.0 * std::log10(peak / rms2);
    }
    void clear() { peaks.clear(); means.clear(); }
};

// Host side of the CFR read-outs (OfdmGenerator.cpp:186-306, 419-453): fed with the per-symbol
// records the symbol kernel writes, one transmission frame at a time, in stream order.
struct CfrReadouts {
    static constexpr size_t MAX_CLIP_STATS = 10;      // OfdmGenerator.cpp:38
    PaprWindow before, after;
    std::deque<double> clip_ratios, errclip_ratios, mers;
    size_t mer_index = 0;
    bool clear_request = false;

    void init(int n_sym)
    {
        before = PaprWindow{};
        after = PaprWindow{};
        before.n_blocks = after.n_blocks = (size_t)n_sym * 50;    // OfdmGenerator.cpp:59-61
        clip_ratios.clear(); errclip_ratios.clear(); mers.clear();
        mer_index = 0;
        clear_request = false;
    }
    void add_frame(const CfrSymStat *rec, int n_sym, int N)
    {
        mer_index = (mer_index + 1) % (size_t)n_sym;
        if (clear_request) { before.clear(); after.clear(); clear_request = false; }
        size_t num_clip = 0, num_err = 0;
        for (int i = 0; i < n_sym; i++) {
            const CfrSymStat &r = rec[i];
            before.add(r.peak_before, (double)r.sum_before / N);
            if (i > 0) after.add(r.peak_after, (double)r.sum_after / N);
            if (i > 0 && mer_index == (size_t)i) {
                // ETSI ETR 290 annex C; by Parseval the time-domain sums of the reference
                // (OfdmGenerator.cpp:262-271) are N times these frequency-domain sums
                mers.push_back(r.sum_delta > 0 ? 10.0 * std::log10((double)r.sum_ref / (double)r.sum_delta) : 90);
            }
            num_clip += r.clip;
            num_err += r.errclip;
        }
        const double n_samps = (double)n_sym * N;
        clip_ratios.push_back((double)num_clip / n_samps);
        errclip_ratios.push_back((double)num_err / n_samps);
        while (clip_ratios.size() > MAX_CLIP_STATS) clip_ratios.pop_front();
        while (errclip_ratios.size() > MAX_CLIP_STATS) errclip_ratios.pop_front();
        while (mers.size() > MAX_CLIP_STATS) mers.pop_front();
    }
    static double avg(const std::deque<double> &d)
    {
        double a = 0;
        for (double v : d) a += v;
        return a / (double)d.size();
    }
    std::string clip_stats() const      // OfdmGenerator.cpp:420-441
    {
        std::stringstream ss;
        if (clip_ratios.empty() || errclip_ratios.empty() || mers.empty()) ss << "No stats available";
        else
            ss << "Statistics : " << std::fixed << avg(clip_ratios) * 100 << "% samples clipped, "
               << avg(errclip_ratios) * 100 << "% errors clipped. MER after CFR: " << avg(mers) << " dB";
        return ss.str();
    }
    std::string papr() const            // OfdmGenerator.cpp:443-452
    {
        const double b = before.papr(), a = after.papr();
        std::stringstream ss;
        ss << "PAPR [dB]: " << std::fixed << (b == 0 ? std::string("N/A") : std::to_string(b)) << ", "
           << (a == 0 ? std::string("N/A") : std::to_string(a));
        return ss.str();
    }
};

} // namespace

struct dabmod_b200 {
    dabmod_b200_config cfg{};
    ModeInfo m{};
    std::mutex mtx;                // guards parameters against set_param from RC threads
    int device = 0;
    int sm_count = 148;
    cudaStream_t s_compute = nullptr, s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_done;

    // parameters (host copies)
    std::vector<float> fir_taps;
    int dpd_mode = 0;
    float dpd[33] = {0};
    bool use_cic = false;
    bool tii_supported = false;
    uint64_t tf_counter = 0;       // TFs processed since create/reset (TII parity)
    bool tables_dirty = true;
    int force_chunks = 0;          // test/tuning knob: CTAs per TF in k_symbols (0 = automatic)

    // device tables
    DevBuf<uint16_t> d_bin_of_src;
    DevBuf<uint8_t> d_phase0;
    DevBuf<float> d_cic;
    DevBuf<float> d_twiddle;       // interleaved re/im, per-pass tables (symbol_fft_twiddles)
    DevBuf<float> d_twiddle_w;     // second-pass table of k_symbols_w (TM I only)
    bool use_warp_kernel = true;   // "sym_kernel" knob: 0 = always the CTA-per-symbol-group kernel
    bool use_fir_sym = true;       // "fir_kernel" knob: 0 = always the sample-stream FIR kernel
    int fir_kernel = 3;            //   1 = k_fir_sym, 2 = k_fir_tma where it applies (complexf output),
                                   //   3 = inside the symbol kernel where that applies (k_symbols_w FUSE) and pays
                                   //   (batch size), else like 2; 4 = 3 at any batch size
    int n_twiddle = 0;
    DevBuf<uint16_t> d_tii_bin;
    DevBuf<float> d_tii_val;
    DevBuf<float> d_window;
    DevBuf<float> d_lut;
    DevBuf<float> d_fir_taps;      // taps of a filter longer than MAX_FIR_TAPS (k_fir_long), zero padded
    DevBuf<float2> d_tii_frame;    // one frame through k_symbols: its null symbol is the stream's TII symbol (k_tii_fill)
    DevBuf<float2> d_tii_fir;      // ... and the same through the FIR: the filtered null symbol of a TII frame (fused kernel)
    DevBuf<uint8_t> d_tii_bits;    // that frame's (all-zero) input block
    DevBuf<unsigned long long> d_clipped;
    int tii_count = 0;

    // fixed-point engine (symbols_fixed.cuh)
    DevBuf<uint16_t> d_fx_pos_of_src, d_fx_tii_pos;
    DevBuf<short2> d_fx_tw, d_fx_tii_val;
    DevBuf<short> d_fx_window;
    std::vector<int> fx_p, fx_m;   // kf_factor's (radix, remaining size) list

    // work buffers for max_batch TFs
    DevBuf<uint8_t> d_bits;
    DevBuf<unsigned char> d_out;
    DevBuf<float2> d_tmp;          // symbol-stage output when another kernel follows
    DevBuf<float2> d_tmp2;         // FIR output when the resampler follows

    // resampler (Resampler.cpp:51-112)
    bool has_res = false;
    ResamplerPlan rp{};
    std::vector<unsigned char> rad_in, rad_out;
    DevBuf<float> d_res_win;
    DevBuf<float2> d_tw_in, d_tw_out;
    DevBuf<float2> d_hist;         // last Ni input samples of the stream (zeros at stream start)
    DevBuf<float2> d_scratch;
    int res_grid = 0;
    bool res_up = false;           // k_resample_up applies (Ni = 4096, integer ratio <= 4)
    bool res_q = false;            // k_resample_q applies (Ni = 4096, No = P * 4000, P = 2..5)
    bool allow_res_up = true;      // "res_kernel" knob: 0 = always the generic kernel
    int res_kernel = 1;            //   2 = the older variant of a fast kernel where two exist
    int res_dbg = 0;               // "res_dbg": profiling aid, skips parts of the resampler kernels (wrong results)
    size_t res_smem = 0;

    uint64_t clipped_last = 0;
    bool clipped_pending = false;  // a device-path call left its count in d_clipped (read back on demand)
    uint32_t launches_last = 0;

    // Calls on one handle are stream-ordered by the library: every enqueue records `ev_last` on the stream it
    // used and the next one (whatever its stream) waits for it, because the work buffers, the resampler history
    // and the tables are per handle.  `ev_tables` orders a table rebuild (on s_compute) before a user stream.
    cudaEvent_t ev_last = nullptr, ev_tables = nullptr;
    bool have_last = false;

    // file sink (dabmod_b200_process_batch_to_fd): ring of pinned host buffers, one slice each
    static constexpr int SINK_SLOTS = 3;
    unsigned char *sink_buf[SINK_SLOTS] = {nullptr, nullptr, nullptr};
    size_t sink_cap = 0;
    std::vector<cudaEvent_t> ev_out;

    // CFR read-outs ("clip_stats", "papr"): per-symbol records of the last launch, aggregated on the host
    DevBuf<CfrSymStat> d_cfr_stats;
    CfrReadouts cfr_readouts;
    bool cfr_collect = true;       // false while seek() re-runs a frame that is not part of this handle's range
    size_t cfr_pending = 0;        // TFs whose records wait in d_cfr_stats

    // optional per-kernel timing
    bool profile = false;
    struct Timed { const char *name; cudaEvent_t a, b; };
    std::vector<Timed> timed;
    std::vector<cudaEvent_t> event_pool;
    size_t events_used = 0;
    cudaEvent_t next_event()
    {
        if (events_used == event_pool.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) throw std::runtime_error("cudaEventCreate failed");
            event_pool.push_back(e);
        }
        return event_pool[events_used++];
    }

    size_t out_samples_per_tf() const
    {
        return has_res ? (size_t)m.tf_samples * rp.L / rp.M : (size_t)m.tf_samples;
    }
    bool fixed() const { return cfg.fft_engine == DABMOD_B200_FFT_KISS_FIXED; }
    size_t out_bytes_per_tf() const { return out_samples_per_tf() * (fixed() ? 4 : format_bytes(cfg.format)); }
    bool has_fir() const { return !fir_taps.empty(); }
    bool has_post() const { return dpd_mode != 0 || cfg.format != DABMOD_B200_FMT_COMPLEXF; }
};

namespace {

void build_tables(dabmod_b200 *h)
{
    const ModeInfo &m = h->m;
    const dabmod_b200_config &c = h->cfg;
    cudaStream_t s = h->s_compute;

    const std::vector<int> dest = interleaver_dest(m);
    const std::vector<uint8_t> q = phase_ref_quarter_turns(m);
    std::vector<uint16_t> bin(m.K);
    std::vector<uint8_t> ph0(m.K);
    for (int j = 0; j < m.K; j++) {
        bin[j] = (uint16_t)bin_of_carrier(m, dest[j]);
        ph0[j] = (uint8_t)(2 * q[dest[j]]);
    }
    h->d_bin_of_src.upload(bin, s);
    h->d_phase0.upload(ph0, s);

    unsigned ratio = 1;
    const uint64_t rate = c.output_rate ? c.output_rate : 2048000;
    h->use_cic = cic_enabled(c.clock_rate, rate, ratio);
    std::vector<float> cic_pos;
    if (h->use_cic) {
        cic_pos = cic_filter(m.K, (float)m.N * (float)rate / 2048000.0f, (int)ratio);   // truncates like the reference's size_t parameter
        std::vector<float> cic_src(m.K);
        for (int j = 0; j < m.K; j++) cic_src[j] = cic_pos[dest[j]];
        h->d_cic.upload(cic_src, s);
    }

    // TII symbol (TII.cpp:172-211): carrier pair (ix, ix+1) = phaseref[ix] twice,
    // or phaseref[ix], phaseref[ix+1] for the old variant.
    std::vector<int> pairs;
    h->tii_supported = tii_pairs(m, c.tii_comb, c.tii_pattern, pairs);
    std::vector<uint16_t> tbin;
    std::vector<float> tval;
    if (h->tii_supported && c.tii_enable) {
        static const float qre[4] = {1, 0, -1, 0}, qim[4] = {0, 1, 0, -1};
        for (int ix : pairs) {
            for (int o = 0; o < 2; o++) {
                const int src = (o == 1 && c.tii_old_variant) ? ix + 1 : ix;
                float g = h->use_cic ? cic_pos[ix + o] : 1.0f;
                tbin.push_back((uint16_t)bin_of_carrier(m, ix + o));
                tval.push_back(qre[q[src]] * g);
                tval.push_back(qim[q[src]] * g);
            }
        }
    }
    h->tii_count = (int)tbin.size();
    if (h->tii_count > MAX_TII * 2) throw ApiError(DABMOD_B200_EINVAL, "too many TII carriers");
    h->d_tii_bin.upload(tbin, s);
    h->d_tii_val.upload(tval, s);

    if (c.window_overlap > 0) h->d_window.upload(guard_window(c.window_overlap), s);

    if (h->dpd_mode == DABMOD_B200_DPD_LUT) {
        std::vector<float> lut(h->dpd + 1, h->dpd + 33);
        h->d_lut.upload(lut, s);
    }
    if ((int)h->fir_taps.size() > MAX_FIR_TAPS) {
        std::vector<float> t(h->fir_taps);
        t.resize((t.size() + FIR_CHUNK - 1) / FIR_CHUNK * FIR_CHUNK, 0.0f);
        h->d_fir_taps.upload(t, s);
    }
    if (h->fixed()) {
        // KISS plan of the mode's transform (kiss_fft.c:293-315 kf_factor, :325-355 kiss_fft_alloc)
        const int N = m.N;
        h->fx_p.clear(); h->fx_m.clear();
        {
            int p = 4, rem = N;
            const double floor_sqrt = std::floor(std::sqrt((double)N));
            do {
                while (rem % p) {
                    switch (p) {
                        case 4: p = 2; break;
                        case 2: p = 3; break;
                        default: p += 2; break;
                    }
                    if (p > floor_sqrt) p = rem;
                }
                rem /= p;
                h->fx_p.push_back(p);
                h->fx_m.push_back(rem);
            } while (rem > 1);
        }
        for (int p : h->fx_p)
            if (p != 2 && p != 4) throw ApiError(DABMOD_B200_EUNSUPPORTED, "fixed-point FFT: radix other than 2 and 4");
        if ((int)h->fx_p.size() > FX_MAX_STAGES) throw ApiError(DABMOD_B200_EUNSUPPORTED, "fixed-point FFT: too many stages");
        std::vector<short2> tw(N);
        for (int i = 0; i < N; i++) {
            const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
            const double phase = 2 * pi * i / N;          // inverse transform: -(-2 pi i / N)
            tw[i] = make_short2((short)std::floor(.5 + 32767 * std::cos(phase)),
                                (short)std::floor(.5 + 32767 * std::sin(phase)));
        }
        h->d_fx_tw.upload(tw, s);
        // kf_work reads the input in mixed-radix digit-reversed order: position -> input index
        std::vector<int> idx_of_pos(N), pos_of_idx(N);
        std::function<void(int, int, int, size_t)> fill = [&](int pos0, int idx0, int fstride, size_t level) {
            const int p = h->fx_p[level], mm = h->fx_m[level];
            for (int q = 0; q < p; q++) {
                if (mm == 1) idx_of_pos[pos0 + q] = idx0 + q * fstride;
                else fill(pos0 + q * mm, idx0 + q * fstride, fstride * p, level + 1);
            }
        };
        fill(0, 0, 1, 0);
        for (int i = 0; i < N; i++) pos_of_idx[idx_of_pos[i]] = i;
        std::vector<uint16_t> pos_src(m.K);
        for (int j = 0; j < m.K; j++) pos_src[j] = (uint16_t)pos_of_idx[bin[j]];
        h->d_fx_pos_of_src.upload(pos_src, s);
        std::vector<uint16_t> tpos;
        std::vector<short2> tv;
        for (size_t i = 0; i < tbin.size(); i++) {
            tpos.push_back((uint16_t)pos_of_idx[tbin[i]]);
            // PhaseReference.cpp:139-150: fixed_16{1} = 16384
            tv.push_back(make_short2((short)(tval[2 * i] * 16384.0f), (short)(tval[2 * i + 1] * 16384.0f)));
        }
        h->d_fx_tii_pos.upload(tpos, s);
        h->d_fx_tii_val.upload(tv, s);
        if (c.window_overlap > 0) {
            // GuardIntervalInserter.cpp:103-112: windowFix[i] = fixed((double)(float)value), rounded
            const std::vector<float> wf = guard_window(c.window_overlap);
            std::vector<short> wi(wf.size());
            for (size_t i = 0; i < wf.size(); i++) wi[i] = (short)((double)wf[i] * 16384.0 + 0.5);
            h->d_fx_window.upload(wi, s);
        }
    }
    h->tables_dirty = false;
}

PostParams make_post(dabmod_b200 *h, bool enabled)
{
    PostParams pp{};
    if (!enabled) return pp;
    pp.dpd_mode = h->dpd_mode;
    pp.format = h->cfg.format;
    if (h->dpd_mode == DABMOD_B200_DPD_ODD_POLY) {
        for (int i = 0; i < 5; i++) { pp.am[i] = h->dpd[i]; pp.pm[i] = h->dpd[5 + i]; }
    }
    else if (h->dpd_mode == DABMOD_B200_DPD_LUT) {
        pp.lut_scale = h->dpd[0];
        pp.lut = h->d_lut.p;
    }
    pp.clipped = h->d_clipped.p;
    return pp;
}

template <int N, bool POST, bool OPT>
void launch_symbols_npo(const SymParams &p, int grid, cudaStream_t s)
{
    const size_t smem = sizeof(SymSmem) + (OPT ? sizeof(SymSmemOpt) : 0);
    CUDA_CHECK(cudaFuncSetAttribute(k_symbols<N, POST, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_symbols<N, POST, OPT><<<grid, SYM_THREADS, smem, s>>>(p);
}

template <int N>
void launch_symbols_n(const SymParams &p, bool post, bool opt, int grid, cudaStream_t s)
{
    if (opt) {
        if (post) launch_symbols_npo<N, true, true>(p, grid, s);
        else launch_symbols_npo<N, false, true>(p, grid, s);
    }
    else {
        if (post) launch_symbols_npo<N, true, false>(p, grid, s);
        else launch_symbols_npo<N, false, false>(p, grid, s);
    }
}

template <bool POST>
void launch_fir_p(const FirParams &p, int ntaps, int grid, cudaStream_t s)
{
    if (ntaps <= 16) k_fir<16, POST><<<grid, FIR_THREADS, 0, s>>>(p);
    else if (ntaps <= 32) k_fir<32, POST><<<grid, FIR_THREADS, 0, s>>>(p);
    else if (ntaps <= 45) k_fir<45, POST><<<grid, FIR_THREADS, 0, s>>>(p);   // the reference's default filter
    else if (ntaps <= 64) k_fir<64, POST><<<grid, FIR_THREADS, 0, s>>>(p);
    else k_fir<0, POST><<<grid, FIR_THREADS, 0, s>>>(p);                     // any length, chunked tap loop
}

struct ProfScope {
    dabmod_b200 *h;
    cudaStream_t s;
    size_t slot = 0;
    bool on;
    ProfScope(dabmod_b200 *h_, const char *name, cudaStream_t s_) : h(h_), s(s_), on(h_->profile)
    {
        if (!on) return;
        dabmod_b200::Timed t{name, h->next_event(), h->next_event()};
        cudaEventRecord(t.a, s);
        slot = h->timed.size();
        h->timed.push_back(t);
    }
    void end()
    {
        if (on) cudaEventRecord(h->timed[slot].b, s);
    }
};

// Radix schedule of a Stockham FFT of size n with radices from {16, 8, 4, 2, 5, 3, 7}.
bool fft_radices(int n, std::vector<unsigned char> &rad)
{
    rad.clear();
    if (n < 2) return false;
    while (n % 16 == 0) { rad.push_back(16); n /= 16; }
    while (n % 8 == 0) { rad.push_back(8); n /= 8; }
    while (n % 4 == 0) { rad.push_back(4); n /= 4; }
    while (n % 2 == 0) { rad.push_back(2); n /= 2; }
    for (int f : {5, 3, 7})
        while (n % f == 0) { rad.push_back((unsigned char)f); n /= f; }
    // larger primes: a generic O(R) pass (resample.cuh generic_pass_prime); the radix is stored in a byte
    for (int f = 11; f <= 251 && n > 1; f += 2)
        while (n % f == 0) { rad.push_back((unsigned char)f); n /= f; }
    return n == 1 && rad.size() <= (size_t)RES_MAX_PASSES;
}

// symbols [-> FIR]: d_bits -> dst (float2 stream, or the final format when `last`)
// Fetch the CFR records of the frames processed since the last call and feed the read-outs.
void consume_cfr(dabmod_b200 *h)
{
    if (!h->cfr_pending) return;
    const size_t n_sym = (size_t)h->m.L + 1;
    std::vector<CfrSymStat> rec(h->cfr_pending * n_sym);
    CUDA_CHECK(cudaSetDevice(h->device));
    if (h->have_last) CUDA_CHECK(cudaEventSynchronize(h->ev_last));   // the enqueue that wrote them
    CUDA_CHECK(cudaMemcpy(rec.data(), h->d_cfr_stats.p, rec.size() * sizeof(CfrSymStat), cudaMemcpyDeviceToHost));
    for (size_t tf = 0; tf < h->cfr_pending; tf++)
        h->cfr_readouts.add_frame(rec.data() + tf * n_sym, (int)n_sym, h->m.N);
    h->cfr_pending = 0;
}

void enqueue_front(dabmod_b200 *h, const uint8_t *d_bits, size_t n_tf, void *dst, bool last, size_t tmp_tf0,
                   uint64_t stream_tf, cudaStream_t s, uint32_t &launches)
{
    const ModeInfo &m = h->m;
    const dabmod_b200_config &c = h->cfg;
    const bool fir = h->has_fir();
    const bool post = last && h->has_post();

    SymParams sp{};
    sp.L = m.L; sp.K = m.K; sp.N = m.N;
    sp.null_size = m.null_size; sp.sym_size = m.sym_size;
    sp.tf_in_bytes = m.tf_in_bytes; sp.tf_samples = m.tf_samples;
    sp.G = SYM_POINTS / m.N;
    sp.n_groups = (m.L + 1 + sp.G - 1) / sp.G;
    // chunking: enough CTAs to fill the machine several times over, but chunks
    // long enough to amortise the per-CTA tables and the phase prefix
    {
        const int target_ctas = h->sm_count * 6 * 4;
        int chunks = (int)std::min<size_t>((size_t)sp.n_groups / 2, std::max<size_t>(1, (target_ctas + n_tf - 1) / n_tf));
        chunks = std::max(1, std::min(chunks, 11));
        if (h->force_chunks > 0) chunks = std::max(1, std::min(h->force_chunks, sp.n_groups / 2));
        sp.groups_per_chunk = (sp.n_groups + chunks - 1) / chunks;
        if (sp.groups_per_chunk < 2) sp.groups_per_chunk = 2;
        sp.n_chunks = (sp.n_groups + sp.groups_per_chunk - 1) / sp.groups_per_chunk;
    }
    sp.bin_of_src = h->d_bin_of_src.p;
    sp.phase0 = h->d_phase0.p;
    sp.cic = h->use_cic ? h->d_cic.p : nullptr;
    sp.twiddle = reinterpret_cast<const float2 *>(h->d_twiddle.p);
    sp.n_twiddle = h->n_twiddle;
    sp.tii_count = h->tii_count;
    sp.tii_parity = 0;
    sp.tii_bin = h->d_tii_bin.p;
    sp.tii_val = reinterpret_cast<const float2 *>(h->d_tii_val.p);
    sp.cfr = c.cfr_enable;
    sp.cfr_clip = c.cfr_clip;
    sp.cfr_errclip = c.cfr_errclip;
    sp.cfr_stats = nullptr;
    if (c.cfr_enable && h->cfr_collect) {
        if (!h->d_cfr_stats.p) h->d_cfr_stats.alloc((size_t)c.max_batch * (m.L + 1));
        sp.cfr_stats = h->d_cfr_stats.p + tmp_tf0 * (size_t)(m.L + 1);
        h->cfr_pending = std::max(h->cfr_pending, tmp_tf0 + n_tf);
    }
    sp.gain_mode = c.gain_mode;
    sp.gain_const = c.normalise * c.digital_gain;
    sp.var_factor = c.gain_variance;
    sp.window = c.window_overlap;
    sp.window_tab = h->d_window.p;
    sp.bits = d_bits;
    sp.tf_offset = stream_tf;
    const bool sym_last = !fir;
    sp.out = sym_last ? dst : (void *)(h->d_tmp.p + tmp_tf0 * (size_t)m.tf_samples);
    const bool sym_post = sym_last && post;
    sp.post = make_post(h, sym_post);

    const bool sym_opt = c.cfr_enable != 0 || c.window_overlap > 0;
    // TII frames (every second one) differ from plain frames in the null symbol only, and that symbol is one
    // constant vector per stream (k_tii_fill): with enough frames in the call the warp-per-symbol kernels run and
    // the TII symbols are filled in afterwards; the general kernel handles CicEq and short calls
    const bool tii_fill = h->tii_count > 0 && !h->use_cic && n_tf >= 16;
    const bool warp_ok = h->use_warp_kernel && !sym_opt && !h->use_cic && (h->tii_count == 0 || tii_fill);
    const bool warp_kernel = m.N == SW_N && warp_ok;
    // ... followed by the default-length FIR: the symbol kernel leaves out null symbol and cyclic prefix,
    // k_fir_sym works symbol by symbol on that compact layout (kernels.cuh); not with TII (the null symbol is not zero)
    const bool compact = warp_kernel && fir && h->fir_taps.size() == 45 && h->use_fir_sym && h->tii_count == 0;
    // one frame through the general kernel: its null symbol is the stream's TII symbol (see k_tii_fill)
    auto make_tii_frame = [&]() {
        if (!h->d_tii_frame.p) {
            h->d_tii_frame.alloc((size_t)m.tf_samples);
            h->d_tii_bits.alloc((size_t)m.tf_in_bytes);
            CUDA_CHECK(cudaMemsetAsync(h->d_tii_bits.p, 0, (size_t)m.tf_in_bytes, s));
        }
        SymParams tp = sp;
        tp.bits = h->d_tii_bits.p;
        tp.out = h->d_tii_frame.p;
        tp.tf_offset = 0;                               // frame 0 of a stream carries the TII symbol
        tp.post = PostParams{};
        tp.groups_per_chunk = 2;                        // the null symbol and symbol 1 are all that is needed
        tp.n_chunks = 1;
        ProfScope prof_t(h, "k_symbols", s);
        switch (m.N) {
            case 2048: launch_symbols_n<2048>(tp, false, false, 1, s); break;
            case 1024: launch_symbols_n<1024>(tp, false, false, 1, s); break;
            default: launch_symbols_n<512>(tp, false, false, 1, s); break;
        }
        CUDA_CHECK(cudaGetLastError());
        prof_t.end();
        launches++;
    };
    // TM I without the optional per-carrier features: one warp per symbol (symbols_warp.cuh)
    if (warp_kernel) {
        SymWParams wp{};
        wp.s = sp;
        wp.compact = compact ? 1 : 0;
        wp.twiddle_w = reinterpret_cast<const float2 *>(h->d_twiddle_w.p);
        wp.n_tf = (int)n_tf;
        // persistent: one CTA per SM, every warp takes one contiguous range of the batch's symbols
        const long long n_sym = (long long)n_tf * m.L;
        const int wgrid = (int)std::min<long long>((n_sym + SW_WARPS - 1) / SW_WARPS, h->sm_count);
        // "fir_kernel" = 3: the 45-tap FIR inside the symbol kernel (complexf out of the FIR stage; symbols_warp.cuh FUSE)
        // (from eight symbols per warp on: below that the one symbol more that a warp assembles to finish its range
        // -- see the kernel -- outweighs the saved intermediate; measured 64 / 128 TFs: two kernels 2-4 % faster)
        // ("fir_kernel" = 4: at any batch size -- the tests walk the kernel's edge cases with few frames)
        // With TII the null symbol of every second frame is one constant vector per stream: its filtered samples
        // (all but the last 44, which see symbol 1) are computed once per call -- one frame through k_symbols and
        // k_fir -- and copied into the TII frames; the fused kernel takes the symbol's last 44 samples as the carry
        // into symbol 1 instead of zeros.
        const bool fused = warp_kernel && fir && h->fir_taps.size() == 45 && h->use_fir_sym && !post &&
                           (h->tii_count == 0 || tii_fill) &&
                           (h->fir_kernel >= 4 || (h->fir_kernel == 3 && (long long)n_tf * m.L >= 8LL * h->sm_count * SW_WARPS));
        if (fused && h->tii_count > 0) {
            make_tii_frame();
            if (!h->d_tii_fir.p) h->d_tii_fir.alloc((size_t)m.tf_samples);
            FirParams fp{};
            fp.in = h->d_tii_frame.p;
            fp.out = h->d_tii_fir.p;
            fp.tf_samples = m.tf_samples;
            fp.tiles_per_tf = (m.null_size + FIR_TILE - 1) / FIR_TILE;      // the tiles that cover the null symbol
            fp.ntaps = 45;
            std::memset(fp.taps, 0, sizeof(fp.taps));
            for (size_t j = 0; j < 45; j++) fp.taps[j] = make_float2(h->fir_taps[j], h->fir_taps[j]);
            fp.post = PostParams{};
            ProfScope prof_tf(h, "k_fir", s);
            launch_fir_p<false>(fp, 45, fp.tiles_per_tf, s);
            CUDA_CHECK(cudaGetLastError());
            prof_tf.end();
            launches++;
            wp.tii_tail = h->d_tii_frame.p + (m.null_size - 44);
        }
        ProfScope prof_w(h, fused ? "k_symbols_w_fir" : "k_symbols_w", s);
        if (fused) {
            for (size_t j = 0; j < 45 && j < h->fir_taps.size(); j++) wp.taps[j] = make_float2(h->fir_taps[j], h->fir_taps[j]);
            wp.compact = 0;
            wp.s.out = dst;
            auto kern = k_symbols_w<false, true>;
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SymWSmem)));
            kern<<<wgrid, SW_THREADS, sizeof(SymWSmem), s>>>(wp);
            CUDA_CHECK(cudaGetLastError());
            prof_w.end();
            launches++;
            if (h->tii_count > 0) {
                const int fsz = m.null_size - 44;
                const int fgrid = (int)std::min<size_t>(n_tf * (size_t)((fsz + 255) / 256), (size_t)h->sm_count * 8);
                ProfScope prof_f(h, "k_tii_fill", s);
                k_tii_fill<false><<<fgrid, 256, 0, s>>>(h->d_tii_fir.p, dst, fsz, m.tf_samples, (int)n_tf, stream_tf, PostParams{});
                CUDA_CHECK(cudaGetLastError());
                prof_f.end();
                launches++;
            }
            return;
        }
        if (sym_post) {
            CUDA_CHECK(cudaFuncSetAttribute(k_symbols_w<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(SymWSmem)));
            k_symbols_w<true><<<wgrid, SW_THREADS, sizeof(SymWSmem), s>>>(wp);
        }
        else {
            CUDA_CHECK(cudaFuncSetAttribute(k_symbols_w<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(SymWSmem)));
            k_symbols_w<false><<<wgrid, SW_THREADS, sizeof(SymWSmem), s>>>(wp);
        }
        CUDA_CHECK(cudaGetLastError());
        prof_w.end();
        launches++;
    }
    else if ((m.N == 1024 || m.N == 512) && warp_ok) {
        // TM IV / II without the optional per-carrier features: 64 / R1 symbols per warp (symbols_warp_g.cuh)
        SymWgParams wp{};
        wp.s = sp;
        wp.twiddle = reinterpret_cast<const float2 *>(h->d_twiddle_w.p);
        wp.n_tf = (int)n_tf;
        const int G = 2048 / m.N;
        wp.n_groups = (m.L + G - 1) / G;
        const long long n_grp = (long long)n_tf * wp.n_groups;
        const int wgrid = (int)std::min<long long>((n_grp + SW_WARPS - 1) / SW_WARPS, h->sm_count);
        ProfScope prof_w(h, "k_symbols_wg", s);
        auto go = [&](auto kern, size_t smem) {
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<wgrid, SW_THREADS, smem, s>>>(wp);
        };
        switch (m.N * 2 + (sym_post ? 1 : 0)) {
            case 2048: go(k_symbols_wg<32, false>, sizeof(SymWgSmem<32>)); break;
            case 2049: go(k_symbols_wg<32, true>, sizeof(SymWgSmem<32>)); break;
            case 1024: go(k_symbols_wg<16, false>, sizeof(SymWgSmem<16>)); break;
            default: go(k_symbols_wg<16, true>, sizeof(SymWgSmem<16>)); break;
        }
        CUDA_CHECK(cudaGetLastError());
        prof_w.end();
        launches++;
    }
    else {
    const int grid = (int)(n_tf * sp.n_chunks);
    ProfScope prof_sym(h, "k_symbols", s);
    switch (m.N) {
        case 2048: launch_symbols_n<2048>(sp, sym_post, sym_opt, grid, s); break;
        case 1024: launch_symbols_n<1024>(sp, sym_post, sym_opt, grid, s); break;
        case 512: launch_symbols_n<512>(sp, sym_post, sym_opt, grid, s); break;
        default: launch_symbols_n<256>(sp, sym_post, sym_opt, grid, s); break;
    }
    CUDA_CHECK(cudaGetLastError());
    prof_sym.end();
    launches++;
    }

    if (tii_fill && (warp_kernel || ((m.N == 1024 || m.N == 512) && warp_ok))) {
        // one frame through the general kernel (its null symbol is the TII symbol), then into every TII frame
        make_tii_frame();
        const int fgrid = (int)std::min<size_t>(n_tf * (size_t)((m.null_size + 255) / 256), (size_t)h->sm_count * 8);
        ProfScope prof_f(h, "k_tii_fill", s);
        if (sym_post) k_tii_fill<true><<<fgrid, 256, 0, s>>>(h->d_tii_frame.p, sp.out, m.null_size, m.tf_samples, (int)n_tf, stream_tf, sp.post);
        else k_tii_fill<false><<<fgrid, 256, 0, s>>>(h->d_tii_frame.p, sp.out, m.null_size, m.tf_samples, (int)n_tf, stream_tf, sp.post);
        CUDA_CHECK(cudaGetLastError());
        prof_f.end();
        launches++;
    }

    if (fir && compact) {
        FirSymParams fp{};
        fp.in = reinterpret_cast<const float2 *>(sp.out);
        fp.out = dst;
        fp.L = m.L; fp.null_size = m.null_size; fp.sym_size = m.sym_size; fp.tf_samples = m.tf_samples;
        std::memset(fp.taps, 0, sizeof(fp.taps));
        for (size_t j = 0; j < h->fir_taps.size(); j++) fp.taps[j] = make_float2(h->fir_taps[j], h->fir_taps[j]);
        fp.post = make_post(h, post);
        const dim3 fgrid((unsigned)m.L, (unsigned)n_tf);
        const bool tma = !post && h->fir_kernel >= 2;
        ProfScope prof_fir(h, tma ? "k_fir_tma" : "k_fir_sym", s);
        if (tma) {
            const long long n_items = (long long)n_tf * m.L;
            const int grid = (int)std::min<long long>(n_items, (long long)h->sm_count * FIRT_CTAS_PER_SM);
            CUDA_CHECK(cudaFuncSetAttribute(k_fir_tma<45>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(FirTmaSmem)));
            k_fir_tma<45><<<grid, FIRT_THREADS, sizeof(FirTmaSmem), s>>>(fp, n_items);
        }
        else if (post) k_fir_sym<45, true><<<fgrid, FIRS_THREADS, 0, s>>>(fp);
        else k_fir_sym<45, false><<<fgrid, FIRS_THREADS, 0, s>>>(fp);
        CUDA_CHECK(cudaGetLastError());
        prof_fir.end();
        launches++;
    }
    else if (fir && (int)h->fir_taps.size() > MAX_FIR_TAPS) {
        // any tap count (FIRFilter.cpp:95-141): taps in a device table, window in dynamic shared memory
        FirLongParams fp{};
        fp.in = reinterpret_cast<const float2 *>(sp.out);
        fp.out = dst;
        fp.tf_samples = m.tf_samples;
        fp.tiles_per_tf = (m.tf_samples + FIR_TILE - 1) / FIR_TILE;
        fp.ntaps = (int)h->fir_taps.size();
        fp.taps = h->d_fir_taps.p;
        fp.post = make_post(h, post);
        const size_t smem = fir_long_smem(fp.ntaps);
        const int fgrid = (int)(n_tf * fp.tiles_per_tf);
        ProfScope prof_fir(h, "k_fir_long", s);
        if (post) {
            CUDA_CHECK(cudaFuncSetAttribute(k_fir_long<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fir_long<true><<<fgrid, FIR_THREADS, smem, s>>>(fp);
        }
        else {
            CUDA_CHECK(cudaFuncSetAttribute(k_fir_long<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fir_long<false><<<fgrid, FIR_THREADS, smem, s>>>(fp);
        }
        CUDA_CHECK(cudaGetLastError());
        prof_fir.end();
        launches++;
    }
    else if (fir) {
        FirParams fp{};
        fp.in = reinterpret_cast<const float2 *>(sp.out);
        fp.out = dst;
        fp.tf_samples = m.tf_samples;
        fp.tiles_per_tf = (m.tf_samples + FIR_TILE - 1) / FIR_TILE;
        fp.ntaps = (int)h->fir_taps.size();
        std::memset(fp.taps, 0, sizeof(fp.taps));
        for (size_t j = 0; j < h->fir_taps.size(); j++) fp.taps[j] = make_float2(h->fir_taps[j], h->fir_taps[j]);
        fp.post = make_post(h, post);
        const int fgrid = (int)(n_tf * fp.tiles_per_tf);
        ProfScope prof_fir(h, "k_fir", s);
        if (post) launch_fir_p<true>(fp, (int)h->fir_taps.size(), fgrid, s);
        else launch_fir_p<false>(fp, (int)h->fir_taps.size(), fgrid, s);
        CUDA_CHECK(cudaGetLastError());
        prof_fir.end();
        launches++;
    }
}

// Resampler over n_tf TFs of the stream: in (float2, n_tf*tf_samples) -> d_out
void enqueue_resampler(dabmod_b200 *h, const float2 *in, size_t n_tf, void *d_out, cudaStream_t s,
                       uint32_t &launches)
{
    const ResamplerPlan &rp = h->rp;
    const bool post = h->has_post();
    ResParams p{};
    p.ni = rp.ni; p.no = rp.no;
    p.total_hops = (long long)(n_tf * (size_t)h->m.tf_samples / (size_t)(rp.ni / 2));
    p.factor = rp.factor;
    p.in = in;
    p.hist = h->d_hist.p;
    p.win = h->d_res_win.p;
    p.tw_in = h->d_tw_in.p;
    p.tw_out = h->d_tw_out.p;
    p.n_rad_in = (int)h->rad_in.size();
    p.n_rad_out = (int)h->rad_out.size();
    std::memcpy(p.rad_in, h->rad_in.data(), h->rad_in.size());
    std::memcpy(p.rad_out, h->rad_out.data(), h->rad_out.size());
    p.scratch = h->res_smem ? nullptr : h->d_scratch.p;
    p.out = d_out;
    p.post = make_post(h, post);
    p.dbg = h->res_dbg;
    if (h->res_up && h->allow_res_up) {
        // TM I, integer up-sampling: L transforms of Ni points per hop, all in shared memory
        RuParams pu{};
        pu.r = p;
        pu.L = (int)rp.L;
        if (h->res_kernel != 2) {
            // three teams per SM (resample_up.cuh); "res_kernel" = 2 keeps the two-team kernel with output staging
            const int grid3 = (int)std::min<long long>((p.total_hops + RU3_TEAMS - 1) / RU3_TEAMS, h->sm_count);
            ProfScope prof3(h, "k_resample_up3", s);
            auto go = [&](auto kern) {
                CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Ru3Smem)));
                kern<<<grid3, RU3_THREADS, sizeof(Ru3Smem), s>>>(pu);
            };
            switch (pu.L * 2 + (post ? 1 : 0)) {
                case 4: go(k_resample_up3<false, 2>); break;
                case 5: go(k_resample_up3<true, 2>); break;
                case 6: go(k_resample_up3<false, 3>); break;
                case 7: go(k_resample_up3<true, 3>); break;
                case 8: go(k_resample_up3<false, 4>); break;
                default: go(k_resample_up3<true, 4>); break;
            }
            CUDA_CHECK(cudaGetLastError());
            prof3.end();
            launches++;
        }
        else {
        const int grid = (int)std::min<long long>((p.total_hops + RU_TEAMS - 1) / RU_TEAMS, h->sm_count);
        ProfScope prof(h, "k_resample_up", s);
        if (post) {
            CUDA_CHECK(cudaFuncSetAttribute(k_resample_up<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(RuSmem)));
            k_resample_up<true><<<grid, RU_THREADS, sizeof(RuSmem), s>>>(pu);
        }
        else {
            CUDA_CHECK(cudaFuncSetAttribute(k_resample_up<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(RuSmem)));
            k_resample_up<false><<<grid, RU_THREADS, sizeof(RuSmem), s>>>(pu);
        }
        CUDA_CHECK(cudaGetLastError());
        prof.end();
        launches++;
        }
    }
    else if (h->res_q && h->allow_res_up) {
        // TM I, No = P * 4000 (4 / 6 / 8 / 10 Msps): P phase transforms of 4000 points per hop in shared memory
        RqParams pq{};
        pq.r = p;
        pq.P = rp.no / RQ_Q;
        const int grid = (int)std::min<long long>((p.total_hops + RQ_TEAMS - 1) / RQ_TEAMS, h->sm_count);
        ProfScope prof(h, "k_resample_q", s);
        if (post) {
            CUDA_CHECK(cudaFuncSetAttribute(k_resample_q<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(RqSmem)));
            k_resample_q<true><<<grid, RQ_THREADS, sizeof(RqSmem), s>>>(pq);
        }
        else {
            CUDA_CHECK(cudaFuncSetAttribute(k_resample_q<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(RqSmem)));
            k_resample_q<false><<<grid, RQ_THREADS, sizeof(RqSmem), s>>>(pq);
        }
        CUDA_CHECK(cudaGetLastError());
        prof.end();
        launches++;
    }
    else {
    const int grid = (int)std::min<long long>(p.total_hops, h->res_grid);
    ProfScope prof(h, "k_resample", s);
    if (post) {
        CUDA_CHECK(cudaFuncSetAttribute(k_resample_generic<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)h->res_smem));
        k_resample_generic<true><<<grid, RESG_THREADS, h->res_smem, s>>>(p);
    }
    else {
        CUDA_CHECK(cudaFuncSetAttribute(k_resample_generic<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)h->res_smem));
        k_resample_generic<false><<<grid, RESG_THREADS, h->res_smem, s>>>(p);
    }
    CUDA_CHECK(cudaGetLastError());
    prof.end();
    launches++;
    }
    // stream state for the next launch: the last Ni input samples (Resampler.cpp:143-145,185-191)
    const size_t total = n_tf * (size_t)h->m.tf_samples;
    CUDA_CHECK(cudaMemcpyAsync(h->d_hist.p, in + total - rp.ni, sizeof(float2) * rp.ni, cudaMemcpyDeviceToDevice, s));
}

// Fixed-point engine: one kernel, bits -> int16 I/Q (symbols_fixed.cuh)
void enqueue_fixed(dabmod_b200 *h, const uint8_t *d_bits, size_t n_tf, void *dst, uint64_t stream_tf, cudaStream_t s,
                   uint32_t &launches)
{
    const ModeInfo &m = h->m;
    const dabmod_b200_config &c = h->cfg;
    FixParams fp{};
    fp.L = m.L; fp.K = m.K; fp.N = m.N;
    fp.null_size = m.null_size; fp.sym_size = m.sym_size;
    fp.tf_in_bytes = m.tf_in_bytes; fp.tf_samples = m.tf_samples;
    fp.G = FX_POINTS / m.N;
    fp.n_groups = (m.L + 1 + fp.G - 1) / fp.G;
    {
        // as in k_symbols: enough CTAs to fill the machine a few times, chunks long enough to amortise the
        // tables and the phase prefix; with windowing a TF is one chunk (the falling edges chain through it)
        const int target_ctas = h->sm_count * 8 * 2;
        int chunks = (int)std::min<size_t>((size_t)fp.n_groups / 2, std::max<size_t>(1, (target_ctas + n_tf - 1) / n_tf));
        chunks = std::max(1, std::min(chunks, 11));
        if (h->force_chunks > 0) chunks = std::max(1, std::min(h->force_chunks, fp.n_groups / 2));
        if (c.window_overlap > 0) chunks = 1;
        fp.groups_per_chunk = (fp.n_groups + chunks - 1) / chunks;
        fp.n_chunks = (fp.n_groups + fp.groups_per_chunk - 1) / fp.groups_per_chunk;
    }
    fp.pos_of_src = h->d_fx_pos_of_src.p;
    fp.phase0 = h->d_phase0.p;
    fp.tw = h->d_fx_tw.p;
    fp.n_stages = (int)h->fx_p.size();
    for (int i = 0; i < fp.n_stages; i++) { fp.stage_p[i] = h->fx_p[i]; fp.stage_m[i] = h->fx_m[i]; }
    fp.tii_count = h->tii_count;
    fp.tii_parity = 0;
    fp.tii_pos = h->d_fx_tii_pos.p;
    fp.tii_val = h->d_fx_tii_val.p;
    fp.window = c.window_overlap;
    fp.window_tab = h->d_fx_window.p;
    fp.bits = d_bits;
    fp.out = reinterpret_cast<short2 *>(dst);
    fp.tf_offset = stream_tf;
    ProfScope prof(h, "k_symbols_fix", s);
    const unsigned grid = (unsigned)(n_tf * fp.n_chunks);
    switch (m.N) {
        case 2048: k_symbols_fix<2048><<<grid, FX_THREADS, 0, s>>>(fp); break;
        case 1024: k_symbols_fix<1024><<<grid, FX_THREADS, 0, s>>>(fp); break;
        case 512: k_symbols_fix<512><<<grid, FX_THREADS, 0, s>>>(fp); break;
        default: k_symbols_fix<256><<<grid, FX_THREADS, 0, s>>>(fp); break;
    }
    CUDA_CHECK(cudaGetLastError());
    prof.end();
    launches++;
}

// Enqueue the kernel family for n_tf TFs: d_bits -> d_out.  `tmp_tf0` = index of
// the first TF within the handle's work buffers (for the temp buffer offsets).
void enqueue(dabmod_b200 *h, const uint8_t *d_bits, size_t n_tf, void *d_out, size_t tmp_tf0,
             uint64_t stream_tf, cudaStream_t s, uint32_t &launches)
{
    if (h->fixed()) {
        enqueue_fixed(h, d_bits, n_tf, d_out, stream_tf, s, launches);
        return;
    }
    if (!h->has_res) {
        enqueue_front(h, d_bits, n_tf, d_out, true, tmp_tf0, stream_tf, s, launches);
        return;
    }
    float2 *front = (h->has_fir() ? h->d_tmp2.p : h->d_tmp.p) + tmp_tf0 * (size_t)h->m.tf_samples;
    enqueue_front(h, d_bits, n_tf, front, false, tmp_tf0, stream_tf, s, launches);
    enqueue_resampler(h, front, n_tf, d_out, s, launches);
}

// GuardIntervalInserter.cpp:149-300 needs windowOverlap <= symSize - spacing (it computes
// `remaining_prefix_length` as an unsigned difference); the kernel keeps 2W window entries on chip.
void check_window(int mode, int overlap)
{
    if (overlap < 0) throw ApiError(DABMOD_B200_EINVAL, "windowlen must be >= 0");
    const ModeInfo m = mode_info(mode == 0 ? 1 : mode);
    if (overlap > m.sym_size - m.N)
        throw ApiError(DABMOD_B200_EINVAL, "windowlen " + std::to_string(overlap) +
                                               " exceeds the guard interval of this mode");
    if (2 * overlap > MAX_WINDOW)
        throw ApiError(DABMOD_B200_EUNSUPPORTED, "windowlen above " + std::to_string(MAX_WINDOW / 2));
}

void validate_config(const dabmod_b200_config &c)
{
    if (c.abi_version != DABMOD_B200_ABI_VERSION)
        throw ApiError(DABMOD_B200_EINVAL, "dabmod_b200_config.abi_version mismatch");
    if (c.mode < 0 || c.mode > 4) throw ApiError(DABMOD_B200_EINVAL, "DabModulator::setMode invalid mode size");
    if (c.gain_mode < 0 || c.gain_mode > 2) throw ApiError(DABMOD_B200_EINVAL, "Internal error: invalid gainmode");
    if (c.fir_ntaps < 0 || (c.fir_ntaps > 0 && !c.fir_taps)) throw ApiError(DABMOD_B200_EINVAL, "FIRFilter: taps missing");
    if (c.fir_ntaps > MAX_FIR_TAPS_LONG)
        throw ApiError(DABMOD_B200_EUNSUPPORTED, "FIRFilter: more than " + std::to_string(MAX_FIR_TAPS_LONG) + " taps are not supported");
    if (c.dpd_mode < 0 || c.dpd_mode > 2 || (c.dpd_mode && !c.dpd_coefs))
        throw ApiError(DABMOD_B200_EINVAL, "MemlessPoly: invalid coefficients");
    format_bytes(c.format);
    check_window(c.mode, c.window_overlap);
    if (c.max_batch < 0) throw ApiError(DABMOD_B200_EINVAL, "max_batch < 0");
    if (c.fft_engine != DABMOD_B200_FFT_FLOAT && c.fft_engine != DABMOD_B200_FFT_KISS_FIXED)
        throw ApiError(DABMOD_B200_EINVAL, "unknown fft_engine");
    if (c.fft_engine == DABMOD_B200_FFT_KISS_FIXED) {
        // DabModulator.cpp:249,257,265
        if (c.fir_ntaps > 0) throw ApiError(DABMOD_B200_EINVAL, "fixed point doesn't support fir filter");
        if (c.dpd_mode != 0) throw ApiError(DABMOD_B200_EINVAL, "fixed point doesn't support predistortion");
        if (c.output_rate != 0 && c.output_rate != 2048000)
            throw ApiError(DABMOD_B200_EINVAL, "fixed point doesn't support resampler");
        // OfdmGeneratorFixed has no CFR; CicEqualizer has no fixed-point variant (it would read the
        // complexfix carriers as floats in the reference)
        if (c.cfr_enable) throw ApiError(DABMOD_B200_EINVAL, "fixed point doesn't support crest factor reduction");
        if (c.clock_rate != 0) throw ApiError(DABMOD_B200_EUNSUPPORTED, "fixed point: CicEqualizer is float only");
        if (c.format != DABMOD_B200_FMT_COMPLEXF && c.format != DABMOD_B200_FMT_S16)
            throw ApiError(DABMOD_B200_EINVAL, "fixed point output is s16");
        if (2 * c.window_overlap > FX_MAX_WINDOW)
            throw ApiError(DABMOD_B200_EUNSUPPORTED, "windowlen above " + std::to_string(FX_MAX_WINDOW / 2));
    }
}

int guard(const std::function<void()> &fn)
{
    try {
        fn();
        return DABMOD_B200_OK;
    }
    catch (const ApiError &e) {
        g_last_error = e.what();
        return e.code;
    }
    catch (const std::bad_alloc &) {
        g_last_error = "out of host memory";
        return DABMOD_B200_ENOMEM;
    }
    catch (const std::exception &e) {
        g_last_error = e.what();
        return DABMOD_B200_EINVAL;
    }
}
} // namespace

extern "C" {

const char *dabmod_b200_last_error(void) { return g_last_error.c_str(); }

void dabmod_b200_config_init(dabmod_b200_config *cfg)
{
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->abi_version = DABMOD_B200_ABI_VERSION;
    cfg->mode = 1;
    cfg->gain_mode = DABMOD_B200_GAIN_VAR;
    cfg->output_rate = 2048000;
    cfg->digital_gain = 1.0f;
    cfg->normalise = 1.0f;
    cfg->gain_variance = 4.0f;
    cfg->cfr_clip = 1.0f;
    cfg->cfr_errclip = 1.0f;
    cfg->max_batch = 1;
}

int dabmod_b200_default_fir_taps(float *taps, int cap)
{
    const std::vector<float> t = default_fir_taps();
    if (taps) std::memcpy(taps, t.data(), sizeof(float) * std::min<size_t>(t.size(), cap > 0 ? cap : 0));
    return (int)t.size();
}

int dabmod_b200_create(const dabmod_b200_config *cfg, dabmod_b200 **out)
{
    if (out) *out = nullptr;
    dabmod_b200 *h = nullptr;
    int rc = guard([&] {
        if (!cfg || !out) throw ApiError(DABMOD_B200_EINVAL, "null argument");
        validate_config(*cfg);
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw ApiError(DABMOD_B200_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                                  " (this library has no CPU fallback)");
        if (cfg->device < 0 || cfg->device >= ndev) throw ApiError(DABMOD_B200_EINVAL, "invalid device ordinal");
        CUDA_CHECK(cudaSetDevice(cfg->device));
        cudaDeviceProp prop{};
        CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
        if (prop.major != 10)
            throw ApiError(DABMOD_B200_ECUDA, "device is not compute capability 10.x (built for sm_100a only)");

        h = new dabmod_b200();
        h->cfg = *cfg;
        if (h->cfg.mode == 0) h->cfg.mode = 1;
        if (h->cfg.output_rate == 0) h->cfg.output_rate = 2048000;
        if (h->cfg.max_batch <= 0) h->cfg.max_batch = 1;
        h->device = cfg->device;
        h->sm_count = prop.multiProcessorCount;
        h->m = mode_info(h->cfg.mode);
        if (cfg->fir_ntaps > 0) h->fir_taps.assign(cfg->fir_taps, cfg->fir_taps + cfg->fir_ntaps);
        h->cfg.fir_taps = nullptr;
        h->dpd_mode = cfg->dpd_mode;
        if (cfg->dpd_mode == DABMOD_B200_DPD_ODD_POLY) std::memcpy(h->dpd, cfg->dpd_coefs, 10 * sizeof(float));
        if (cfg->dpd_mode == DABMOD_B200_DPD_LUT) std::memcpy(h->dpd, cfg->dpd_coefs, 33 * sizeof(float));
        h->cfg.dpd_coefs = nullptr;

        CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_tables, cudaEventDisableTiming));

        h->m = mode_info(h->cfg.mode);
        std::vector<float> tw;
        symbol_fft_twiddles(h->m.N, tw);
        h->d_twiddle.upload(tw, h->s_compute);
        h->n_twiddle = (int)(tw.size() / 2);
        if (h->m.N == SW_N) {
            std::vector<float> tww;
            for (int r = 1; r < 32; r++)
                for (int j = 0; j < 64; j++) {
                    const double a = 2.0 * M_PI * (double)j * (double)r / (double)SW_N;
                    tww.push_back((float)cos(a));
                    tww.push_back((float)sin(a));
                }
            h->d_twiddle_w.upload(tww, h->s_compute);
            CUDA_CHECK(cudaStreamSynchronize(h->s_compute));
        }
        else {
            // k_symbols_wg (TM II / III / IV): the N roots of unity, its second-pass table is built from them
            std::vector<float> tww;
            for (int k = 0; k < h->m.N; k++) {
                const double a = 2.0 * M_PI * (double)k / (double)h->m.N;
                tww.push_back((float)cos(a));
                tww.push_back((float)sin(a));
            }
            h->d_twiddle_w.upload(tww, h->s_compute);
            CUDA_CHECK(cudaStreamSynchronize(h->s_compute));
        }
        h->d_clipped.alloc(1);
        CUDA_CHECK(cudaMemsetAsync(h->d_clipped.p, 0, sizeof(unsigned long long), h->s_compute));
        build_tables(h);
        h->cfr_readouts.init(h->m.L + 1);

        const size_t nb = (size_t)h->cfg.max_batch;
        if (h->cfg.output_rate != 2048000) {
            // DabModulator.cpp:260-268: Resampler(2048000, outputRate, spacing)
            const ResamplerPlan rp = resampler_plan(2048000, h->cfg.output_rate, h->m.N);
            if (rp.ni < 4 || rp.no < 2 || h->m.tf_samples % (rp.ni / 2) != 0)
                throw ApiError(DABMOD_B200_EUNSUPPORTED,
                               "Resampler: FFT sizes " + std::to_string(rp.ni) + "/" + std::to_string(rp.no) +
                                   " do not tile the transmission frame (the reference reads out of bounds here)");
            if (!fft_radices(rp.ni, h->rad_in) || !fft_radices(rp.no, h->rad_out))
                throw ApiError(DABMOD_B200_EUNSUPPORTED,
                               "Resampler: FFT size " + std::to_string(rp.no) + " has a prime factor above 251");
            h->rp = rp;
            h->has_res = true;
            h->res_up = rp.ni == RU_NI && rp.M == 1 && rp.L >= 2 && rp.L <= (uint64_t)RU_MAX_L;
            h->res_q = rp.ni == RQ_NI && rp.no > rp.ni && rp.no % RQ_Q == 0 && rp.no / RQ_Q >= RQ_MIN_P &&
                       rp.no / RQ_Q <= RQ_MAX_P;
            h->d_res_win.upload(resampler_window(rp.ni), h->s_compute);
            std::vector<float> t;
            twiddle_table(rp.ni, t);
            h->d_tw_in.alloc(rp.ni);
            CUDA_CHECK(cudaMemcpyAsync(h->d_tw_in.p, t.data(), sizeof(float2) * rp.ni, cudaMemcpyHostToDevice, h->s_compute));
            twiddle_table(rp.no, t);
            h->d_tw_out.alloc(rp.no);
            CUDA_CHECK(cudaMemcpyAsync(h->d_tw_out.p, t.data(), sizeof(float2) * rp.no, cudaMemcpyHostToDevice, h->s_compute));
            CUDA_CHECK(cudaStreamSynchronize(h->s_compute));   // `t` is reused / goes out of scope
            h->d_hist.alloc(rp.ni);
            CUDA_CHECK(cudaMemsetAsync(h->d_hist.p, 0, sizeof(float2) * rp.ni, h->s_compute));
            const size_t nmax = (size_t)std::max(rp.ni, rp.no);
            const size_t need = 2 * nmax * sizeof(float2);
            if (need <= 96 * 1024) {
                h->res_smem = need;
                h->res_grid = h->sm_count * 2;
            }
            else {
                h->res_smem = 0;
                h->res_grid = h->sm_count * 2;
                h->d_scratch.alloc((size_t)h->res_grid * 2 * nmax);
            }
        }
        h->d_bits.alloc(nb * h->m.tf_in_bytes);
        h->d_out.alloc(nb * h->out_bytes_per_tf());
        if (h->has_fir() || h->has_res) h->d_tmp.alloc(nb * (size_t)h->m.tf_samples);
        if (h->has_fir() && h->has_res) h->d_tmp2.alloc(nb * (size_t)h->m.tf_samples);
        CUDA_CHECK(cudaStreamSynchronize(h->s_compute));
        *out = h;
    });
    if (rc != DABMOD_B200_OK && h) {
        dabmod_b200_destroy(h);
    }
    return rc;
}

void dabmod_b200_destroy(dabmod_b200 *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->s_compute) cudaStreamSynchronize(h->s_compute);
    if (h->s_in) cudaStreamSynchronize(h->s_in);
    if (h->s_out) cudaStreamSynchronize(h->s_out);
    if (h->have_last) cudaEventSynchronize(h->ev_last);
    if (h->ev_last) cudaEventDestroy(h->ev_last);
    if (h->ev_tables) cudaEventDestroy(h->ev_tables);
    for (auto e : h->event_pool) cudaEventDestroy(e);
    for (auto e : h->ev_in) cudaEventDestroy(e);
    for (auto e : h->ev_done) cudaEventDestroy(e);
    for (auto e : h->ev_out) cudaEventDestroy(e);
    for (auto &b : h->sink_buf) if (b) cudaFreeHost(b);
    if (h->s_compute) cudaStreamDestroy(h->s_compute);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
}

size_t dabmod_b200_tf_in_bytes(const dabmod_b200 *h) { return h ? (size_t)h->m.tf_in_bytes : 0; }
size_t dabmod_b200_tf_out_bytes(const dabmod_b200 *h) { return h ? h->out_bytes_per_tf() : 0; }
size_t dabmod_b200_tf_out_samples(const dabmod_b200 *h) { return h ? h->out_samples_per_tf() : 0; }

} // extern "C"

namespace {

// Start of every call that enqueues work on stream `s`: order it after the previous call on this handle
// (work buffers, resampler history and tables are per handle) and rebuild the tables if a parameter changed.
void begin_call(dabmod_b200 *h, cudaStream_t s)
{
    if (h->tables_dirty) {
        // kernels of the previous call may still be reading the tables
        if (h->have_last) CUDA_CHECK(cudaStreamWaitEvent(h->s_compute, h->ev_last, 0));
        build_tables(h);
        if (s != h->s_compute) {
            CUDA_CHECK(cudaEventRecord(h->ev_tables, h->s_compute));
            CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_tables, 0));
        }
    }
    if (h->have_last) CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_last, 0));
    h->launches_last = 0;
    h->timed.clear();
    h->events_used = 0;
}

void end_call(dabmod_b200 *h, cudaStream_t s)
{
    CUDA_CHECK(cudaEventRecord(h->ev_last, s));
    h->have_last = true;
}

// FormatConverter::get_num_clipped_samples of a call whose result was left on the device
void refresh_clipped(dabmod_b200 *h)
{
    if (!h->clipped_pending) return;
    CUDA_CHECK(cudaSetDevice(h->device));
    if (h->have_last) CUDA_CHECK(cudaEventSynchronize(h->ev_last));
    unsigned long long v = 0;
    CUDA_CHECK(cudaMemcpy(&v, h->d_clipped.p, sizeof(v), cudaMemcpyDeviceToHost));
    h->clipped_last = v;
    h->clipped_pending = false;
}

// write(2) until everything is out (OutputFile.cpp:56-67 uses fwrite and throws on a short count)
void write_all(int fd, const unsigned char *p, size_t n)
{
    while (n) {
        const ssize_t w = ::write(fd, p, n);
        if (w < 0) {
            if (errno == EINTR) continue;
            throw ApiError(DABMOD_B200_EIO, std::string("OutputFile: write failed: ") + std::strerror(errno));
        }
        p += w;
        n -= (size_t)w;
    }
}

void grow_events(std::vector<cudaEvent_t> &v, size_t n)
{
    while (v.size() < n) {
        cudaEvent_t e;
        CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        v.push_back(e);
    }
}

void seek_locked(dabmod_b200 *h, uint64_t tf_index, const uint8_t *prev_bits, bool on_device)
{
    h->tf_counter = tf_index;
    if (!h->has_res) return;
    CUDA_CHECK(cudaSetDevice(h->device));
    cudaStream_t s = h->s_compute;
    begin_call(h, s);
    if (!prev_bits || tf_index == 0) {
        CUDA_CHECK(cudaMemsetAsync(h->d_hist.p, 0, sizeof(float2) * h->rp.ni, s));
    }
    else {
        // re-run TF tf_index-1 up to the resampler input; keep its last Ni samples
        const uint8_t *d_prev = prev_bits;
        if (!on_device) {
            CUDA_CHECK(cudaMemcpyAsync(h->d_bits.p, prev_bits, h->m.tf_in_bytes, cudaMemcpyHostToDevice, s));
            d_prev = h->d_bits.p;
        }
        float2 *front = h->has_fir() ? h->d_tmp2.p : h->d_tmp.p;
        uint32_t launches = 0;
        consume_cfr(h);
        h->cfr_collect = false;           // the halo frame belongs to the previous shard's read-outs
        try {
            enqueue_front(h, d_prev, 1, front, false, 0, tf_index - 1, s, launches);
        }
        catch (...) {
            h->cfr_collect = true;
            throw;
        }
        h->cfr_collect = true;
        CUDA_CHECK(cudaMemcpyAsync(h->d_hist.p, front + h->m.tf_samples - h->rp.ni, sizeof(float2) * h->rp.ni,
                                   cudaMemcpyDeviceToDevice, s));
    }
    end_call(h, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
}

} // namespace

namespace dabmod {

cudaStream_t compute_stream(dabmod_b200 *h) { return h->s_compute; }
int device_of(const dabmod_b200 *h) { return h->device; }

void seek_device(dabmod_b200 *h, uint64_t tf_index, const uint8_t *d_prev_bits)
{
    std::lock_guard<std::mutex> lock(h->mtx);
    seek_locked(h, tf_index, d_prev_bits, true);
}

// Software pipeline over slices of the batch on three streams, so that PCIe traffic in both directions hides
// behind the kernels:   H2D(i+1) | front + kernels(i) | D2H(i-1)   [| write(i-2) on this thread for a descriptor]
void run_pipeline(dabmod_b200 *h, size_t n_tf, const PipeFront &front, const PipeSink &sink, size_t *out_bytes)
{
    if (out_bytes) *out_bytes = 0;
    if (!h) throw ApiError(DABMOD_B200_EINVAL, "null argument");
    const bool to_fd = sink.to_fd;
    if (to_fd && sink.fd < 0) throw ApiError(DABMOD_B200_EINVAL, "bad file descriptor");
    if (!to_fd && n_tf && !sink.host_out) throw ApiError(DABMOD_B200_EINVAL, "null argument");
    if (n_tf > (size_t)h->cfg.max_batch) throw ApiError(DABMOD_B200_EINVAL, "n_tf exceeds max_batch of the handle");
    const size_t out_tf = h->out_bytes_per_tf();
    if (!to_fd && sink.cap < n_tf * out_tf) throw ApiError(DABMOD_B200_EINVAL, "output buffer too small");
    std::lock_guard<std::mutex> lock(h->mtx);
    CUDA_CHECK(cudaSetDevice(h->device));
    begin_call(h, h->s_compute);
    if (n_tf == 0) return;
    consume_cfr(h);      // the records of the previous launch, before this one overwrites them
    const bool count_clips = h->cfg.format != DABMOD_B200_FMT_COMPLEXF;
    if (count_clips) CUDA_CHECK(cudaMemsetAsync(h->d_clipped.p, 0, sizeof(unsigned long long), h->s_compute));
    h->clipped_pending = false;

    const size_t slice = std::max<size_t>(1, std::min<size_t>(n_tf, ((to_fd ? 32u : 48u) << 20) / out_tf));
    const size_t n_slices = (n_tf + slice - 1) / slice;
    if (to_fd && h->sink_cap < slice * out_tf) {
        for (auto &b : h->sink_buf) {
            if (b) CUDA_CHECK(cudaFreeHost(b));
            b = nullptr;
            CUDA_CHECK(cudaHostAlloc((void **)&b, slice * out_tf, cudaHostAllocDefault));
        }
        h->sink_cap = slice * out_tf;
    }
    grow_events(h->ev_in, n_slices);
    grow_events(h->ev_done, n_slices);
    if (to_fd) grow_events(h->ev_out, n_slices);

    size_t delivered = 0, next_write = 0;
    auto drain = [&](size_t upto) {                      // write the slices [next_write, upto) to the descriptor
        for (; next_write < upto; next_write++) {
            const size_t t0 = next_write * slice, nt = std::min(slice, n_tf - t0);
            CUDA_CHECK(cudaEventSynchronize(h->ev_out[next_write]));
            write_all(sink.fd, h->sink_buf[next_write % dabmod_b200::SINK_SLOTS], nt * out_tf);
            delivered += nt * out_tf;
        }
    };
    try {
        for (size_t i = 0; i < n_slices; i++) {
            const size_t t0 = i * slice, nt = std::min(slice, n_tf - t0);
            if (to_fd && i >= (size_t)dabmod_b200::SINK_SLOTS) drain(i - dabmod_b200::SINK_SLOTS + 1);   // the slot must be free
            front.upload(t0, nt, h->s_in);
            CUDA_CHECK(cudaEventRecord(h->ev_in[i], h->s_in));
            CUDA_CHECK(cudaStreamWaitEvent(h->s_compute, h->ev_in[i], 0));
            const uint8_t *d_blocks = front.encode(t0, nt, h->s_compute);
            enqueue(h, d_blocks, nt, h->d_out.p + t0 * out_tf, t0, h->tf_counter + t0, h->s_compute, h->launches_last);
            CUDA_CHECK(cudaEventRecord(h->ev_done[i], h->s_compute));
            CUDA_CHECK(cudaStreamWaitEvent(h->s_out, h->ev_done[i], 0));
            unsigned char *dst = to_fd ? h->sink_buf[i % dabmod_b200::SINK_SLOTS] : (unsigned char *)sink.host_out + t0 * out_tf;
            CUDA_CHECK(cudaMemcpyAsync(dst, h->d_out.p + t0 * out_tf, nt * out_tf, cudaMemcpyDeviceToHost, h->s_out));
            if (to_fd) {
                CUDA_CHECK(cudaEventRecord(h->ev_out[i], h->s_out));
                if (i >= 1) drain(i);                    // whatever has landed while this slice was enqueued
            }
        }
        end_call(h, h->s_compute);
        if (to_fd) drain(n_slices);
        else {
            CUDA_CHECK(cudaStreamSynchronize(h->s_out));
            delivered = n_tf * out_tf;
        }
    }
    catch (...) {
        cudaStreamSynchronize(h->s_in);
        cudaStreamSynchronize(h->s_compute);
        cudaStreamSynchronize(h->s_out);
        h->tf_counter += n_tf;                           // the stream position moved, like after a throw in the reference
        if (out_bytes) *out_bytes = delivered;
        throw;
    }
    if (count_clips) {
        unsigned long long v = 0;
        CUDA_CHECK(cudaMemcpy(&v, h->d_clipped.p, sizeof(v), cudaMemcpyDeviceToHost));
        h->clipped_last = v;
    }
    consume_cfr(h);
    h->tf_counter += n_tf;
    if (out_bytes) *out_bytes = delivered;
}

} // namespace dabmod

namespace {
// the plain front: the input already is BlockPartitioner blocks in host memory
PipeFront host_blocks_front(dabmod_b200 *h, const uint8_t *bits)
{
    const size_t in_tf = h->m.tf_in_bytes;
    PipeFront f;
    f.upload = [h, bits, in_tf](size_t t0, size_t nt, cudaStream_t s_in) {
        CUDA_CHECK(cudaMemcpyAsync(h->d_bits.p + t0 * in_tf, bits + t0 * in_tf, nt * in_tf, cudaMemcpyHostToDevice, s_in));
    };
    f.encode = [h, in_tf](size_t t0, size_t, cudaStream_t) -> const uint8_t * { return h->d_bits.p + t0 * in_tf; };
    return f;
}
} // namespace

extern "C" {

int dabmod_b200_process_batch_device(dabmod_b200 *h, const uint8_t *d_bits, size_t n_tf, void *d_iq_out,
                                     void *stream)
{
    return guard([&] {
        if (!h || (n_tf && (!d_bits || !d_iq_out))) throw ApiError(DABMOD_B200_EINVAL, "null argument");
        if (n_tf > (size_t)h->cfg.max_batch)
            throw ApiError(DABMOD_B200_EINVAL, "n_tf exceeds max_batch of the handle");
        if ((reinterpret_cast<uintptr_t>(d_bits) & 3) || (reinterpret_cast<uintptr_t>(d_iq_out) & 15))
            throw ApiError(DABMOD_B200_EINVAL, "device buffers must be aligned (bits: 4 bytes, I/Q: 16 bytes)");
        std::lock_guard<std::mutex> lock(h->mtx);
        CUDA_CHECK(cudaSetDevice(h->device));
        cudaStream_t s = stream ? (cudaStream_t)stream : h->s_compute;
        begin_call(h, s);
        if (n_tf == 0) return;
        consume_cfr(h);      // CFR on: waits for the previous call, its records are overwritten by this one
        if (h->cfg.format != DABMOD_B200_FMT_COMPLEXF) {
            CUDA_CHECK(cudaMemsetAsync(h->d_clipped.p, 0, sizeof(unsigned long long), s));
            h->clipped_pending = true;
        }
        enqueue(h, d_bits, n_tf, d_iq_out, 0, h->tf_counter, s, h->launches_last);
        end_call(h, s);
        h->tf_counter += n_tf;
    });
}

int dabmod_b200_process_batch(dabmod_b200 *h, const uint8_t *bits, size_t n_tf, void *iq_out, size_t cap,
                              size_t *out_bytes)
{
    if (out_bytes) *out_bytes = 0;
    return guard([&] {
        if (!h || (n_tf && (!bits || !iq_out))) throw ApiError(DABMOD_B200_EINVAL, "null argument");
        PipeSink sink;
        sink.host_out = iq_out;
        sink.cap = cap;
        run_pipeline(h, n_tf, host_blocks_front(h, bits), sink, out_bytes);
    });
}

int dabmod_b200_process_batch_to_fd(dabmod_b200 *h, const uint8_t *bits, size_t n_tf, int fd, size_t *out_bytes)
{
    if (out_bytes) *out_bytes = 0;
    return guard([&] {
        if (!h || (n_tf && !bits)) throw ApiError(DABMOD_B200_EINVAL, "null argument");
        if (fd < 0) throw ApiError(DABMOD_B200_EINVAL, "bad file descriptor");
        PipeSink sink;
        sink.to_fd = true;
        sink.fd = fd;
        run_pipeline(h, n_tf, host_blocks_front(h, bits), sink, out_bytes);
    });
}

int dabmod_b200_process(dabmod_b200 *h, const uint8_t *bits, size_t nbytes, void *iq_out, size_t cap,
                        size_t *out_bytes)
{
    if (out_bytes) *out_bytes = 0;
    if (h && nbytes != (size_t)h->m.tf_in_bytes) {
        g_last_error = "QpskSymbolMapper::process input size not valid: " + std::to_string(nbytes) +
                       " != " + std::to_string(h->m.tf_in_bytes);
        return DABMOD_B200_EINVAL;
    }
    return dabmod_b200_process_batch(h, bits, 1, iq_out, cap, out_bytes);
}

int dabmod_b200_synchronize(dabmod_b200 *h)
{
    return guard([&] {
        if (!h) throw ApiError(DABMOD_B200_EINVAL, "null handle");
        CUDA_CHECK(cudaSetDevice(h->device));
        CUDA_CHECK(cudaStreamSynchronize(h->s_in));
        CUDA_CHECK(cudaStreamSynchronize(h->s_compute));
        CUDA_CHECK(cudaStreamSynchronize(h->s_out));
        std::lock_guard<std::mutex> lock(h->mtx);
        if (h->have_last) CUDA_CHECK(cudaEventSynchronize(h->ev_last));   // a caller's stream counts too
        refresh_clipped(h);
    });
}

int dabmod_b200_reset(dabmod_b200 *h)
{
    return guard([&] {
        if (!h) throw ApiError(DABMOD_B200_EINVAL, "null handle");
        std::lock_guard<std::mutex> lock(h->mtx);
        h->tf_counter = 0;
        consume_cfr(h);
        h->cfr_readouts.init(h->m.L + 1);
        if (h->has_res) {
            CUDA_CHECK(cudaSetDevice(h->device));
            if (h->have_last) CUDA_CHECK(cudaStreamWaitEvent(h->s_compute, h->ev_last, 0));
            CUDA_CHECK(cudaMemsetAsync(h->d_hist.p, 0, sizeof(float2) * h->rp.ni, h->s_compute));
            end_call(h, h->s_compute);
            CUDA_CHECK(cudaStreamSynchronize(h->s_compute));
        }
    });
}

int dabmod_b200_seek(dabmod_b200 *h, uint64_t tf_index, const uint8_t *prev_bits, size_t nbytes)
{
    return guard([&] {
        if (!h) throw ApiError(DABMOD_B200_EINVAL, "null handle");
        if (prev_bits && tf_index != 0 && h->has_res && nbytes != (size_t)h->m.tf_in_bytes)
            throw ApiError(DABMOD_B200_EINVAL, "seek: prev_bits must be one TF block");
        std::lock_guard<std::mutex> lock(h->mtx);
        seek_locked(h, tf_index, prev_bits, false);
    });
}

void *dabmod_b200_device_out(dabmod_b200 *h) { return h ? (void *)h->d_out.p : nullptr; }

int dabmod_b200_host_register(void *p, size_t bytes)
{
    return guard([&] {
        if (!p || !bytes) throw ApiError(DABMOD_B200_EINVAL, "null argument");
        CUDA_CHECK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    });
}

int dabmod_b200_host_unregister(void *p)
{
    return guard([&] {
        if (!p) throw ApiError(DABMOD_B200_EINVAL, "null argument");
        CUDA_CHECK(cudaHostUnregister(p));
    });
}

uint64_t dabmod_b200_num_clipped_samples(dabmod_b200 *h)
{
    if (!h) return 0;
    std::lock_guard<std::mutex> lock(h->mtx);
    try { refresh_clipped(h); } catch (const std::exception &e) { g_last_error = e.what(); }
    return h->clipped_last;
}
uint32_t dabmod_b200_last_launch_count(const dabmod_b200 *h) { return h ? h->launches_last : 0; }

int dabmod_b200_set_param(dabmod_b200 *h, const char *name, const char *value)
{
    return guard([&] {
        if (!h || !name || !value) throw ApiError(DABMOD_B200_EINVAL, "null argument");
        std::lock_guard<std::mutex> lock(h->mtx);
        const std::string n(name);
        std::stringstream ss(value);
        ss.exceptions(std::stringstream::failbit | std::stringstream::badbit);
        dabmod_b200_config &c = h->cfg;
        if (h->fixed() && (n == "cfr" || n == "clip" || n == "errorclip" || n == "taps" || n == "coefs" ||
                           n == "digital" || n == "mode" || n == "var"))
            throw ApiError(DABMOD_B200_EINVAL, "Parameter '" + n + "' is not exported by the fixed-point chain "
                                               "(no GainControl / CFR / FIRFilter / MemlessPoly there)");
        if (h->fixed() && n == "windowlen") {
            int v = 0;
            std::stringstream(value) >> v;
            if (2 * v > FX_MAX_WINDOW) throw ApiError(DABMOD_B200_EUNSUPPORTED, "windowlen above " + std::to_string(FX_MAX_WINDOW / 2));
        }
        try {
            if (n == "digital") { ss >> c.digital_gain; }
            else if (n == "profile") { int v; ss >> v; h->profile = v != 0; }
            else if (n == "sym_chunks") { int v; ss >> v; h->force_chunks = v < 0 ? 0 : v; }
            else if (n == "sym_kernel") { int v; ss >> v; h->use_warp_kernel = v != 0; }
            else if (n == "fir_kernel") { int v; ss >> v; h->use_fir_sym = v != 0; h->fir_kernel = v; }
            else if (n == "res_kernel") { int v; ss >> v; h->allow_res_up = v != 0; h->res_kernel = v; }
            else if (n == "res_dbg") { int v; ss >> v; h->res_dbg = v; }
            else if (n == "var") { ss >> c.gain_variance; }
            else if (n == "mode") {
                std::string v; ss >> v;
                for (auto &ch : v) ch = (char)tolower(ch);
                if (v == "fix") c.gain_mode = DABMOD_B200_GAIN_FIX;
                else if (v == "max") c.gain_mode = DABMOD_B200_GAIN_MAX;
                else if (v == "var") c.gain_mode = DABMOD_B200_GAIN_VAR;
                else throw ApiError(DABMOD_B200_EINVAL, "Gainmode " + v + " unknown (fix|max|var)");
            }
            else if (n == "windowlen") {
                int v; ss >> v;
                check_window(c.mode, v);
                c.window_overlap = v; h->tables_dirty = true;
            }
            // (each of the three also restarts the PAPR windows: OfdmGenerator.cpp:384-395)
            else if (n == "cfr") { int v; ss >> v; consume_cfr(h); c.cfr_enable = v != 0; h->cfr_readouts.clear_request = true; }
            else if (n == "clip") { consume_cfr(h); ss >> c.cfr_clip; h->cfr_readouts.clear_request = true; }
            else if (n == "errorclip") { consume_cfr(h); ss >> c.cfr_errclip; h->cfr_readouts.clear_request = true; }
            else if (n == "clip_stats" || n == "papr")
                throw ApiError(DABMOD_B200_EINVAL, "Parameter '" + n + "' is read-only");
            else if (n == "tii.enable") { int v; ss >> v; c.tii_enable = v != 0; h->tables_dirty = true; }
            else if (n == "tii.comb") {
                int v; ss >> v;
                if (v < 0 || v > 23) throw ApiError(DABMOD_B200_EINVAL, "TII comb not valid!");
                c.tii_comb = v; h->tables_dirty = true;
            }
            else if (n == "tii.pattern") {
                int v; ss >> v;
                if (v < 0 || v > 69) throw ApiError(DABMOD_B200_EINVAL, "TII pattern not valid!");
                c.tii_pattern = v; h->tables_dirty = true;
            }
            else if (n == "tii.old_variant") { int v; ss >> v; c.tii_old_variant = v != 0; h->tables_dirty = true; }
            else if (n == "taps") {
                int cnt; ss >> cnt;
                if (cnt <= 0) throw ApiError(DABMOD_B200_EINVAL, "FIRFilter: taps file has invalid format.");
                if (cnt > MAX_FIR_TAPS_LONG)
                    throw ApiError(DABMOD_B200_EUNSUPPORTED, "FIRFilter: more than " + std::to_string(MAX_FIR_TAPS_LONG) + " taps");
                if (!h->has_fir()) throw ApiError(DABMOD_B200_ESTATE, "FIRFilter is not part of this chain");
                std::vector<float> t(cnt);
                for (auto &x : t) ss >> x;
                h->fir_taps = t;
                h->tables_dirty = true;        // taps above MAX_FIR_TAPS live in a device table
            }
            else if (n == "coefs") {
                if (h->dpd_mode == 0) throw ApiError(DABMOD_B200_ESTATE, "MemlessPoly is not part of this chain");
                int fmt; ss >> fmt;
                if (fmt == 1) {
                    int nc; ss >> nc;
                    if (nc != 5) throw ApiError(DABMOD_B200_EINVAL, "MemlessPoly: invalid number of coefs");
                    for (int i = 0; i < 10; i++) ss >> h->dpd[i];
                    h->dpd_mode = DABMOD_B200_DPD_ODD_POLY;
                }
                else if (fmt == 2) {
                    for (int i = 0; i < 33; i++) ss >> h->dpd[i];
                    h->dpd_mode = DABMOD_B200_DPD_LUT;
                    h->tables_dirty = true;
                }
                else throw ApiError(DABMOD_B200_EINVAL, "MemlessPoly: coef file has unknown format");
            }
            else {
                throw ApiError(DABMOD_B200_EINVAL, "Parameter '" + n + "' is not exported by controllable dabmod_b200");
            }
        }
        catch (const std::ios_base::failure &) {
            throw ApiError(DABMOD_B200_EINVAL, "could not parse value for parameter '" + n + "'");
        }
    });
}

int dabmod_b200_get_param(dabmod_b200 *h, const char *name, char *buf, size_t cap)
{
    return guard([&] {
        if (!h || !name || !buf || cap == 0) throw ApiError(DABMOD_B200_EINVAL, "null argument");
        std::lock_guard<std::mutex> lock(h->mtx);
        const std::string n(name);
        const dabmod_b200_config &c = h->cfg;
        std::stringstream ss;
        if (n == "digital") ss << c.digital_gain;
        else if (n == "var") ss << c.gain_variance;
        else if (n == "mode") ss << (c.gain_mode == 0 ? "fix" : c.gain_mode == 1 ? "max" : "var");
        else if (n == "windowlen") ss << c.window_overlap;
        else if (n == "cfr") ss << c.cfr_enable;
        else if (n == "clip") ss << std::fixed << c.cfr_clip;
        else if (n == "errorclip") ss << std::fixed << c.cfr_errclip;
        else if (n == "clip_stats") { consume_cfr(h); ss << h->cfr_readouts.clip_stats(); }
        else if (n == "papr") { consume_cfr(h); ss << h->cfr_readouts.papr(); }
        else if (n == "tii.enable") ss << (c.tii_enable ? 1 : 0);
        else if (n == "tii.comb") ss << c.tii_comb;
        else if (n == "tii.pattern") ss << c.tii_pattern;
        else if (n == "tii.old_variant") ss << (c.tii_old_variant ? 1 : 0);
        else if (n == "ntaps") ss << h->fir_taps.size();
        else if (n == "rate") ss << c.output_rate;
        else if (n == "num_clipped_samples") { refresh_clipped(h); ss << h->clipped_last; }
        else throw ApiError(DABMOD_B200_EINVAL, "Parameter '" + n + "' is not exported by controllable dabmod_b200");
        const std::string s = ss.str();
        if (s.size() + 1 > cap) throw ApiError(DABMOD_B200_EINVAL, "buffer too small");
        std::memcpy(buf, s.c_str(), s.size() + 1);
    });
}

int dabmod_b200_table_interleaver(int mode, int32_t *idx, int cap)
{
    return guard([&] {
        const ModeInfo m = mode_info(mode);
        if (!idx || cap < m.K) throw ApiError(DABMOD_B200_EINVAL, "buffer too small");
        const std::vector<int> d = interleaver_dest(m);
        for (int j = 0; j < m.K; j++) idx[j] = d[j];
    });
}

int dabmod_b200_table_phase_ref(int mode, uint8_t *q, int cap)
{
    return guard([&] {
        const ModeInfo m = mode_info(mode);
        if (!q || cap < m.K) throw ApiError(DABMOD_B200_EINVAL, "buffer too small");
        const std::vector<uint8_t> v = phase_ref_quarter_turns(m);
        std::memcpy(q, v.data(), m.K);
    });
}

int dabmod_b200_table_tii(int mode, int comb, int pattern, uint8_t *acp, int cap)
{
    return guard([&] {
        const ModeInfo m = mode_info(mode);
        if (!acp || cap < m.K) throw ApiError(DABMOD_B200_EINVAL, "buffer too small");
        std::vector<int> pairs;
        if (!tii_pairs(m, comb, pattern, pairs))
            throw ApiError(DABMOD_B200_EUNSUPPORTED, "TII::TII DAB mode " + std::to_string(mode) + " not valid!");
        std::memset(acp, 0, m.K);
        for (int ix : pairs) acp[ix] = 1;
    });
}

int dabmod_b200_table_cic(int n_carriers, float spacing, int ratio, float *filter)
{
    return guard([&] {
        if (!filter || n_carriers <= 0) throw ApiError(DABMOD_B200_EINVAL, "bad argument");
        const std::vector<float> f = cic_filter(n_carriers, spacing, ratio);
        std::memcpy(filter, f.data(), sizeof(float) * n_carriers);
    });
}

int dabmod_b200_resampler_sizes(uint64_t in_rate, uint64_t out_rate, int resolution, int *fft_in, int *fft_out)
{
    return guard([&] {
        if (!in_rate || !out_rate || resolution <= 0 || !fft_in || !fft_out)
            throw ApiError(DABMOD_B200_EINVAL, "bad argument");
        ResamplerPlan rp = resampler_plan(in_rate, out_rate, resolution);
        *fft_in = rp.ni;
        *fft_out = rp.no;
    });
}

int dabmod_b200_kernel_time(dabmod_b200 *h, int idx, char *name, size_t cap, float *ms)
{
    return guard([&] {
        if (!h || !ms) throw ApiError(DABMOD_B200_EINVAL, "null argument");
        std::lock_guard<std::mutex> lock(h->mtx);
        if (idx < 0 || (size_t)idx >= h->timed.size())
            throw ApiError(DABMOD_B200_ESTATE, "no such timed launch (set_param(\"profile\", \"1\") first)");
        const dabmod_b200::Timed &t = h->timed[idx];
        CUDA_CHECK(cudaEventSynchronize(t.b));
        CUDA_CHECK(cudaEventElapsedTime(ms, t.a, t.b));
        if (name && cap) {
            std::strncpy(name, t.name, cap - 1);
            name[cap - 1] = 0;
        }
    });
}

} // extern "C"
