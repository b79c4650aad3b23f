/* TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load the library this
 * file is built into (oracle/_ref/libdabmod_ref.so).
 *
 * Drives the UNMODIFIED reference blocks (compiled from /root/reference by
 * oracle/Makefile) through the reference's own Flowgraph, wired exactly as
 * DabModulator::process does for the hot path (DabModulator.cpp:386-417):
 *
 *   [bits] -> QpskSymbolMapper -> FrequencyInterleaver --+
 *                                 PhaseReference --------+-> DifferentialModulator --+
 *   NullSymbol ----------------------------------------------------------------------+-> SignalMultiplexer
 *   PhaseReference -> TII (TM I/II only) ---------------------------------------------+
 *   -> [CicEqualizer] -> OfdmGeneratorCF32 -> GainControl -> GuardIntervalInserter
 *   -> [FIRFilter] -> [Resampler] -> [MemlessPoly] -> [FormatConverter] -> OutputMemory
 *
 * The only node that is not reference code is BitsInput below, which stands in
 * for BlockPartitioner (it hands the caller's per-TF byte block to the graph).
 *
 * PipelinedModCodec stages (GainControl, FIRFilter, MemlessPoly) each delay the
 * stream by one call (ModPlugin.cpp:90-115); ref_process() therefore returns 0
 * bytes for the first `ref_latency()` calls and afterwards the result for the
 * TF fed `ref_latency()` calls earlier.
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <stdexcept>

#include "Buffer.h"
#include "ModPlugin.h"
#include "Flowgraph.h"
#include "ConfigParser.h"
#include "QpskSymbolMapper.h"
#include "FrequencyInterleaver.h"
#include "PhaseReference.h"
#include "DifferentialModulator.h"
#include "NullSymbol.h"
#include "TII.h"
#include "SignalMultiplexer.h"
#include "CicEqualizer.h"
#include "OfdmGenerator.h"
#include "GainControl.h"
#include "GuardIntervalInserter.h"
#include "FIRFilter.h"
#include "Resampler.h"
#include "MemlessPoly.h"
#include "FormatConverter.h"
#include "OutputMemory.h"

extern "C" {

/* Mirrors the fields of mod_settings_t (ConfigParser.h:45-96) that
 * parametrise the hot path. */
struct ref_cfg {
    int32_t  mode;            /* 1..4 */
    int32_t  gain_mode;       /* 0 fix, 1 max, 2 var (GainControl.h:45) */
    uint64_t output_rate;     /* 2048000 = no resampler */
    uint64_t clock_rate;      /* 0 = no CicEqualizer */
    float    digital_gain;
    float    normalise;
    float    gain_variance;
    int32_t  window_overlap;
    int32_t  cfr_enable;
    float    cfr_clip;
    float    cfr_errclip;
    int32_t  tii_enable;
    int32_t  tii_comb;
    int32_t  tii_pattern;
    int32_t  tii_old_variant;
    int32_t  poly_threads;
    const char *fir_taps_file;   /* NULL/"" = none, "default" = built-in taps */
    const char *poly_coef_file;  /* NULL/"" = none */
    const char *format;          /* NULL/"" = complexf, else s16|u8|s8 */
    /* stop_after: build the chain only up to and including this stage, so
     * intermediate buffers can be compared: qpsk, freq, diff, mux, ciceq,
     * ofdm, gain, guard, fir, resampler, poly, format, NULL/"" = all. */
    const char *stop_after;
    /* 1 = FFTEngine::KISS: the fixed-point chain of DabModulator.cpp:144-224 (complexfix carriers,
     * OfdmGeneratorFixed, no GainControl; FIR / resampler / predistortion are rejected there) */
    int32_t  fixed_point;
};

} // extern "C"

namespace {

class BitsInput : public ModInput {
public:
    const uint8_t *data = nullptr;
    size_t len = 0;
    int process(Buffer *dataOut) override
    {
        dataOut->setData(data, len);
        return (int)len;
    }
    const char *name() override { return "BitsInput"; }
};

struct Mode { unsigned L, K, N, nullSize, symSize; size_t tf_bytes; };

Mode mode_params(int mode)
{
    /* DabModulator.cpp:84-122, BlockPartitioner.cpp:44-73 */
    switch (mode) {
        case 1: return {76, 1536, 2048, 2656, 2552, 75u * 384u};
        case 2: return {76, 384, 512, 664, 638, 75u * 96u};
        case 3: return {153, 192, 256, 345, 319, 152u * 48u};
        case 4: return {76, 768, 1024, 1328, 1276, 75u * 192u};
        default: throw std::runtime_error("invalid mode");
    }
}

struct Harness {
    mod_settings_t s;
    Mode m;
    std::string fmt;
    int latency = 0;
    Buffer out;
    std::shared_ptr<BitsInput> input;
    std::unique_ptr<Flowgraph> fg;
    std::shared_ptr<OfdmGeneratorCF32> ofdm;
    std::string err;
};

thread_local std::string g_err;

} // namespace

extern "C" {

const char *ref_last_error(void) { return g_err.c_str(); }

void *ref_create(const ref_cfg *c)
{
    using namespace std;
    try {
        auto h = make_unique<Harness>();
        mod_settings_t &s = h->s;
        s.dabMode = c->mode;
        s.gainMode = (GainMode)c->gain_mode;
        s.outputRate = c->output_rate ? c->output_rate : 2048000;
        s.clockRate = c->clock_rate;
        s.digitalgain = c->digital_gain;
        s.normalise = c->normalise;
        s.gainmodeVariance = c->gain_variance;
        s.ofdmWindowOverlap = c->window_overlap;
        s.enableCfr = c->cfr_enable != 0;
        s.cfrClip = c->cfr_clip;
        s.cfrErrorClip = c->cfr_errclip;
        s.tiiConfig.enable = c->tii_enable != 0;
        s.tiiConfig.comb = c->tii_comb;
        s.tiiConfig.pattern = c->tii_pattern;
        s.tiiConfig.old_variant = c->tii_old_variant != 0;
        s.filterTapsFilename = c->fir_taps_file ? c->fir_taps_file : "";
        s.polyCoefFilename = c->poly_coef_file ? c->poly_coef_file : "";
        s.polyNumThreads = c->poly_threads;
        s.showProcessTime = false;
        h->fmt = c->format ? c->format : "";
        const string stop = c->stop_after ? c->stop_after : "";

        h->m = mode_params(c->mode);
        const Mode &m = h->m;
        const unsigned mode = c->mode;
        const bool fixedPoint = c->fixed_point != 0;
        if (fixedPoint) {
            /* DabModulator.cpp:249,257,265: "fixed point doesn't support ..." */
            if (!s.filterTapsFilename.empty()) throw std::runtime_error("fixed point doesn't support fir filter");
            if (!s.polyCoefFilename.empty()) throw std::runtime_error("fixed point doesn't support predistortion");
            if (s.outputRate != 2048000) throw std::runtime_error("fixed point doesn't support resampler");
            s.fftEngine = FFTEngine::KISS;
        }

        h->fg = make_unique<Flowgraph>(false);
        Flowgraph &fg = *h->fg;
        h->input = make_shared<BitsInput>();
        auto output = make_shared<OutputMemory>(&h->out);

        auto cifMap = make_shared<QpskSymbolMapper>(m.K, fixedPoint);
        if (stop == "qpsk") {
            fg.connect(h->input, cifMap);
            fg.connect(cifMap, output);
            return h.release();
        }
        auto cifFreq = make_shared<FrequencyInterleaver>(mode, fixedPoint);
        if (stop == "freq") {
            fg.connect(h->input, cifMap);
            fg.connect(cifMap, cifFreq);
            fg.connect(cifFreq, output);
            return h.release();
        }
        auto cifRef = make_shared<PhaseReference>(mode, fixedPoint);
        auto cifDiff = make_shared<DifferentialModulator>(m.K, fixedPoint);
        if (stop == "diff") {
            fg.connect(h->input, cifMap);
            fg.connect(cifMap, cifFreq);
            fg.connect(cifRef, cifDiff);
            fg.connect(cifFreq, cifDiff);
            fg.connect(cifDiff, output);
            return h.release();
        }
        auto cifNull = make_shared<NullSymbol>(m.K, fixedPoint ? sizeof(complexfix) : sizeof(complexf));
        auto cifSig = make_shared<SignalMultiplexer>();

        /* DabModulator.cpp:154-176 */
        bool useCicEq = false;
        unsigned cic_ratio = 1;
        if (s.clockRate) {
            cic_ratio = s.clockRate / s.outputRate;
            cic_ratio /= 4;
            if (s.clockRate == 400000000) {
                if (cic_ratio & 1) useCicEq = true;
            }
            else {
                useCicEq = true;
            }
        }
        shared_ptr<CicEqualizer> cifCicEq;
        if (useCicEq) {
            cifCicEq = make_shared<CicEqualizer>(m.K,
                    (float)m.N * (float)s.outputRate / 2048000.0f, cic_ratio);
        }

        /* DabModulator.cpp:178-190: TII is instantiated whenever the mode
         * supports it, enabled or not. */
        shared_ptr<TII> tii;
        shared_ptr<PhaseReference> tiiRef;
        try {
            tii = make_shared<TII>(s.dabMode, s.tiiConfig, fixedPoint);
            tiiRef = make_shared<PhaseReference>(mode, fixedPoint);
        }
        catch (const TIIError &) {
            tii.reset();
        }

        shared_ptr<ModPlugin> cifOfdm;
        shared_ptr<GainControl> cifGain;
        if (fixedPoint) {
            cifOfdm = make_shared<OfdmGeneratorFixed>(1 + m.L, m.K, m.N);     /* DabModulator.cpp:208-213 */
        }
        else {
            auto ofdm = make_shared<OfdmGeneratorCF32>(1 + m.L, m.K, m.N,
                    s.enableCfr, s.cfrClip, s.cfrErrorClip);
            h->ofdm = ofdm;
            cifOfdm = ofdm;
            cifGain = make_shared<GainControl>(m.N, s.gainMode, s.digitalgain,
                    s.normalise, s.gainmodeVariance);
        }
        auto cifGuard = make_shared<GuardIntervalInserter>(m.L, m.N,
                m.nullSize, m.symSize, s.ofdmWindowOverlap, s.fftEngine);

        shared_ptr<FIRFilter> cifFilter;
        if (!s.filterTapsFilename.empty()) {
            cifFilter = make_shared<FIRFilter>(s.filterTapsFilename);
        }
        shared_ptr<MemlessPoly> cifPoly;
        if (!s.polyCoefFilename.empty()) {
            cifPoly = make_shared<MemlessPoly>(s.polyCoefFilename, s.polyNumThreads);
        }
        shared_ptr<Resampler> cifRes;
        if (s.outputRate != 2048000) {
            cifRes = make_shared<Resampler>(2048000, s.outputRate, m.N);
        }
        shared_ptr<FormatConverter> fmtConv;
        if (!h->fmt.empty() && h->fmt != "complexf") {
            fmtConv = make_shared<FormatConverter>(false, h->fmt);
        }

        /* DabModulator.cpp:386-395, cifPart replaced by BitsInput */
        fg.connect(h->input, cifMap);
        fg.connect(cifMap, cifFreq);
        fg.connect(cifRef, cifDiff);
        fg.connect(cifFreq, cifDiff);
        fg.connect(cifNull, cifSig);
        fg.connect(cifDiff, cifSig);
        if (tii) {
            fg.connect(tiiRef, tii);
            fg.connect(tii, cifSig);
        }

        /* DabModulator.cpp:397-417 */
        struct Stage { const char *name; shared_ptr<ModPlugin> p; int lat; };
        const Stage stages[] = {
            {"ciceq", cifCicEq, 0}, {"ofdm", cifOfdm, 0}, {"gain", cifGain, 1},
            {"guard", cifGuard, 0}, {"fir", cifFilter, 1}, {"resampler", cifRes, 0},
            {"poly", cifPoly, 1}, {"format", fmtConv, 0},
        };
        shared_ptr<ModPlugin> prev = cifSig;
        if (stop != "mux") {
            for (const auto &st : stages) {
                if (st.p) {
                    fg.connect(prev, st.p);
                    prev = st.p;
                    h->latency += st.lat;
                }
                if (stop == st.name) break;
            }
        }
        fg.connect(prev, output);
        return h.release();
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

/* OfdmGeneratorCF32::get_parameter ("clip_stats", "papr", "cfr", ...) of the chain's generator */
int ref_ofdm_get_parameter(void *hp, const char *name, char *buf, size_t cap)
{
    auto h = static_cast<Harness*>(hp);
    try {
        if (!h->ofdm) throw std::runtime_error("chain stops before the OfdmGenerator");
        const std::string v = h->ofdm->get_parameter(name);
        if (v.size() + 1 > cap) throw std::runtime_error("buffer too small");
        memcpy(buf, v.c_str(), v.size() + 1);
        return 0;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

int ref_latency(void *hp) { return static_cast<Harness*>(hp)->latency; }

size_t ref_tf_bytes(void *hp) { return static_cast<Harness*>(hp)->m.tf_bytes; }

/* Feed one TF (BlockPartitioner output bytes); copies the bytes that fall out
 * of OutputMemory into `out`. Returns the byte count (0 while the pipelined
 * stages are priming), or -1 on error. */
long ref_process(void *hp, const uint8_t *bits, size_t nbytes, void *out, size_t cap)
{
    auto h = static_cast<Harness*>(hp);
    try {
        h->input->data = bits;
        h->input->len = nbytes;
        h->out.setLength(0);
        const bool ran = h->fg->run();
        if (!ran) return 0;
        const size_t n = h->out.getLength();
        if (n > cap) {
            g_err = "output buffer too small";
            return -1;
        }
        if (out && n) memcpy(out, h->out.getData(), n);
        return (long)n;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

void ref_destroy(void *hp) { delete static_cast<Harness*>(hp); }

} // extern "C"
