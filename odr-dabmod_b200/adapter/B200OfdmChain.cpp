/* See B200OfdmChain.h.  Host glue only: every sample is computed by the CUDA
 * kernels behind dabmod_b200_process(). */
#include "B200OfdmChain.h"

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "dabmod_b200.h"

namespace {

[[noreturn]] void fail(const char* what)
{
    throw std::runtime_error(std::string("B200OfdmChain: ") + what + ": " + dabmod_b200_last_error());
}

/* FIRFilter::load_filter_taps (src/FIRFilter.cpp:95-141) */
std::vector<float> load_taps(const std::string& file)
{
    std::vector<float> taps;
    if (file == "default") {
        taps.resize(dabmod_b200_default_fir_taps(nullptr, 0));
        dabmod_b200_default_fir_taps(taps.data(), (int)taps.size());
        return taps;
    }
    std::ifstream in(file);
    if (!in) throw std::runtime_error("FIRFilter: Could not open file with taps! " + file);
    int n = 0;
    in >> n;
    if (n <= 0) throw std::runtime_error("FIRFilter: taps file has invalid format.");
    taps.resize(n);
    for (int i = 0; i < n; i++) {
        in >> taps[i];
        if (in.fail()) throw std::runtime_error("FIRFilter: file " + file + " should contain more taps");
    }
    return taps;
}

/* MemlessPoly::load_coefficients (src/MemlessPoly.cpp:145-235): returns the
 * dpd_mode and the coefficient block in dabmod_b200_config layout */
int load_coefs(const std::string& file, std::vector<float>& coefs)
{
    std::ifstream in(file);
    if (!in) throw std::runtime_error("MemlessPoly: Could not open file with coefs!");
    int fmt = 0;
    in >> fmt;
    if (fmt == 1) {
        int n = 0;
        in >> n;
        if (n != 5) throw std::runtime_error("MemlessPoly: invalid number of coefs: " + std::to_string(n));
        coefs.resize(10);
        for (auto& c : coefs) in >> c;
        if (in.fail()) throw std::runtime_error("MemlessPoly: coefs file invalid !");
        return DABMOD_B200_DPD_ODD_POLY;
    }
    if (fmt == 2) {
        coefs.resize(33);
        for (auto& c : coefs) in >> c;
        if (in.fail()) throw std::runtime_error("MemlessPoly: coefs file invalid !");
        return DABMOD_B200_DPD_LUT;
    }
    throw std::runtime_error("MemlessPoly: coef file has unknown format " + std::to_string(fmt));
}

} // namespace

B200OfdmChain::B200OfdmChain(const mod_settings_t& s, const std::string& format, int device, bool fixedPoint) :
    ModCodec(),
    RemoteControllable("b200chain")
{
    dabmod_b200_config c;
    dabmod_b200_config_init(&c);
    c.device = device;
    c.mode = (int32_t)s.dabMode;
    c.gain_mode = (int32_t)s.gainMode;
    c.output_rate = s.outputRate;
    c.clock_rate = s.clockRate;
    c.digital_gain = s.digitalgain;
    c.normalise = s.normalise;
    c.gain_variance = s.gainmodeVariance;
    c.window_overlap = (int32_t)s.ofdmWindowOverlap;
    c.cfr_enable = s.enableCfr;
    c.cfr_clip = s.cfrClip;
    c.cfr_errclip = s.cfrErrorClip;
    c.tii_enable = s.tiiConfig.enable;
    c.tii_comb = s.tiiConfig.comb;
    c.tii_pattern = s.tiiConfig.pattern;
    c.tii_old_variant = s.tiiConfig.old_variant;
    /* the fixed-point chain of DabModulator.cpp:144,194-224 (fftEngine = KISS there) */
    if (fixedPoint || s.fftEngine == FFTEngine::KISS) c.fft_engine = DABMOD_B200_FFT_KISS_FIXED;

    std::vector<float> taps, coefs;
    if (!s.filterTapsFilename.empty()) {
        taps = load_taps(s.filterTapsFilename);
        c.fir_ntaps = (int32_t)taps.size();
        c.fir_taps = taps.data();
    }
    if (!s.polyCoefFilename.empty()) {
        c.dpd_mode = load_coefs(s.polyCoefFilename, coefs);
        c.dpd_coefs = coefs.data();
    }
    if (format.empty() || format == "complexf") c.format = DABMOD_B200_FMT_COMPLEXF;
    else if (format == "s16") c.format = DABMOD_B200_FMT_S16;
    else if (format == "u8") c.format = DABMOD_B200_FMT_U8;
    else if (format == "s8") c.format = DABMOD_B200_FMT_S8;
    else throw std::runtime_error("FormatConverter: Invalid format " + format);
    c.max_batch = 1;

    if (dabmod_b200_create(&c, &m_handle) != DABMOD_B200_OK) fail("create");

    /* names of the replaced blocks' parameters (GainControl.cpp:520-603, TII.cpp:339-376,
     * OfdmGenerator.cpp:63-67, GuardIntervalInserter.cpp:100-103) */
    const char* params[][2] = {
        {"digital", "Digital Gain"},
        {"mode", "Gainmode (fix|max|var)"},
        {"var", "Variance setting for gainmode var (default: 4)"},
        {"tii.enable", "enable TII [0-1]"},
        {"tii.comb", "TII comb number [0-23]"},
        {"tii.pattern", "TII pattern number [0-69]"},
        {"tii.old_variant", "select old TII variant for old (buggy) receivers"},
        {"windowlen", "Window length for OFDM windowng [0 to disable]"},
        {"cfr", "Enable crest factor reduction"},
        {"clip", "CFR: Clip to amplitude"},
        {"errorclip", "CFR: Limit error"},
        {"clip_stats", "CFR: statistics (clip ratio, errorclip ratio)"},
        {"papr", "PAPR measurements (before CFR, after CFR)"},
        {"taps", "FIR taps: count followed by the taps"},
        {"coefs", "Predistorter coefficient file content"},
    };
    for (const auto& p : params) m_parameters.push_back({p[0], p[1]});
}

B200OfdmChain::~B200OfdmChain()
{
    dabmod_b200_destroy(m_handle);
}

int B200OfdmChain::process(Buffer* const dataIn, Buffer* dataOut)
{
    /* same check and message as QpskSymbolMapper::process (src/QpskSymbolMapper.cpp) */
    if (dataIn->getLength() != dabmod_b200_tf_in_bytes(m_handle)) {
        throw std::runtime_error("B200OfdmChain::process input size not valid: " +
                                 std::to_string(dataIn->getLength()));
    }
    dataOut->setLength(dabmod_b200_tf_out_bytes(m_handle));
    size_t n = 0;
    if (dabmod_b200_process(m_handle, reinterpret_cast<const uint8_t*>(dataIn->getData()), dataIn->getLength(),
                            dataOut->getData(), dataOut->getLength(), &n) != DABMOD_B200_OK) {
        fail("process");
    }
    return (int)n;
}

size_t B200OfdmChain::get_num_clipped_samples() const
{
    return dabmod_b200_num_clipped_samples(m_handle);
}

void B200OfdmChain::set_parameter(const std::string& parameter, const std::string& value)
{
    if (dabmod_b200_set_param(m_handle, parameter.c_str(), value.c_str()) != DABMOD_B200_OK) {
        throw ParameterError(dabmod_b200_last_error());
    }
}

const std::string B200OfdmChain::get_parameter(const std::string& parameter) const
{
    char buf[256];
    if (dabmod_b200_get_param(m_handle, parameter.c_str(), buf, sizeof(buf)) != DABMOD_B200_OK) {
        throw ParameterError(dabmod_b200_last_error());
    }
    return buf;
}

const json::map_t B200OfdmChain::get_all_values() const
{
    json::map_t map;
    for (const char* p : {"digital", "mode", "var", "tii.enable", "tii.comb", "tii.pattern", "tii.old_variant"}) {
        map[p].v = get_parameter(p);
    }
    return map;
}
