/* TEST INFRASTRUCTURE -- not product code.
 *
 * Proves the drop-in claim of odr-dabmod_b200/adapter/B200OfdmChain: the
 * adapter is compiled against the UNMODIFIED reference headers and run inside
 * the reference's own Flowgraph (src/Flowgraph.cpp), between a source that
 * stands in for BlockPartitioner and the reference's OutputMemory:
 *
 *     BitsInput -> B200OfdmChain -> OutputMemory
 *
 * tests/test_adapter.py compares its output with the all-reference graph of
 * ref_harness.cpp.  Built into oracle/_ref/libdabmod_adapter.so together with
 * the reference's operator runtime objects; links against libdabmod_b200.so.
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include "Buffer.h"
#include "ConfigParser.h"
#include "Flowgraph.h"
#include "ModPlugin.h"
#include "OutputMemory.h"

#include "B200OfdmChain.h"

extern "C" {
struct ref_cfg {          /* same layout as oracle/ref_harness.cpp */
    int32_t  mode;
    int32_t  gain_mode;
    uint64_t output_rate;
    uint64_t clock_rate;
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
    const char *fir_taps_file;
    const char *poly_coef_file;
    const char *format;
    const char *stop_after;   /* unused here */
    int32_t  fixed_point;     /* 1 = fftEngine KISS: the adapter selects its fixed-point engine */
};
}

namespace {

class BitsInput : public ModInput, public ModMetadata {
public:
    const uint8_t *data = nullptr;
    size_t len = 0;
    int32_t calls = 0;
    /* one timestamp per call, numbered by the call: what EtiReader's metadata looks like to the chain */
    meta_vec_t process_metadata(const meta_vec_t &) override
    {
        flowgraph_metadata md;
        md.ts.fct = calls++;
        md.ts.fp = 0;
        md.ts.timestamp_sec = 0;
        md.ts.timestamp_pps = 0;
        return {md};
    }
    int process(Buffer *dataOut) override
    {
        dataOut->setData(data, len);
        return (int)len;
    }
    const char *name() override { return "BitsInput"; }
};

struct Harness {
    mod_settings_t s;
    Buffer out;
    std::shared_ptr<BitsInput> input;
    std::shared_ptr<B200OfdmChain> chain;
    std::shared_ptr<OutputMemory> output;
    std::unique_ptr<Flowgraph> fg;
};

thread_local std::string g_err;

} // namespace

extern "C" {

const char *adp_last_error(void) { return g_err.c_str(); }

void *adp_create2(const ref_cfg *c, int device, int depth);
void *adp_create(const ref_cfg *c, int device) { return adp_create2(c, device, 0); }

/* depth > 0: the adapter's N-TF pipeline (call i returns TF i - depth, metadata delayed alike) */
void *adp_create2(const ref_cfg *c, int device, int depth)
{
    try {
        auto h = std::make_unique<Harness>();
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
        s.showProcessTime = false;
        if (c->fixed_point) s.fftEngine = FFTEngine::KISS;

        h->fg = std::make_unique<Flowgraph>(false);
        h->input = std::make_shared<BitsInput>();
        h->chain = std::make_shared<B200OfdmChain>(s, c->format ? c->format : "", device, false, depth);
        h->output = std::make_shared<OutputMemory>(&h->out);
        h->fg->connect(h->input, h->chain);
        h->fg->connect(h->chain, h->output);
        return h.release();
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

long adp_process(void *hp, const uint8_t *bits, size_t nbytes, void *out, size_t cap)
{
    auto h = static_cast<Harness *>(hp);
    try {
        h->input->data = bits;
        h->input->len = nbytes;
        h->out.setLength(0);
        if (!h->fg->run()) return 0;
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

/* RemoteControllable::set_parameter / get_parameter through the adapter */
int adp_set_parameter(void *hp, const char *name, const char *value)
{
    try {
        static_cast<Harness *>(hp)->chain->set_parameter(name, value);
        return 0;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

int adp_get_parameter(void *hp, const char *name, char *buf, size_t cap)
{
    try {
        const std::string v = static_cast<Harness *>(hp)->chain->get_parameter(name);
        if (v.size() + 1 > cap) return -1;
        memcpy(buf, v.c_str(), v.size() + 1);
        return 0;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

/* frame count of the timestamp that reached OutputMemory with the last run (-1: none) */
int adp_last_metadata_fct(void *hp)
{
    const auto md = static_cast<Harness *>(hp)->output->get_latest_metadata();
    return md.empty() ? -1 : md.front().ts.fct;
}

/* the reference's separate "tii" controllable (src/TII.cpp:106-127) as the adapter exposes it */
int adp_tii_set(void *hp, const char *name, const char *value)
{
    try {
        auto rc = static_cast<Harness *>(hp)->chain->tii_control();
        if (rc->get_rc_name() != "tii") throw std::runtime_error("controllable is not named tii");
        rc->set_parameter(name, value);
        return 0;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

/* a field of the mod_settings_t the chain was built from: remote-control changes must land there */
double adp_setting(void *hp, const char *name)
{
    const mod_settings_t &s = static_cast<Harness *>(hp)->s;
    const std::string n(name);
    if (n == "digital") return s.digitalgain;
    if (n == "var") return s.gainmodeVariance;
    if (n == "mode") return (double)(int)s.gainMode;
    if (n == "windowlen") return (double)s.ofdmWindowOverlap;
    if (n == "cfr") return s.enableCfr;
    if (n == "clip") return s.cfrClip;
    if (n == "errorclip") return s.cfrErrorClip;
    if (n == "tii.enable") return s.tiiConfig.enable;
    if (n == "tii.comb") return s.tiiConfig.comb;
    if (n == "tii.pattern") return s.tiiConfig.pattern;
    return -1e9;
}

void adp_destroy(void *hp) { delete static_cast<Harness *>(hp); }

} // extern "C"
