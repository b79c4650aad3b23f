/* See B200OfdmChain.h.  Host glue only: every sample is computed by the CUDA
 * kernels behind dabmod_b200_process(). */
#include "B200OfdmChain.h"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "B200Files.h"
#include "dabmod_b200.h"

namespace {

[[noreturn]] void fail(const char* what)
{
    throw std::runtime_error(std::string("B200OfdmChain: ") + what + ": " + dabmod_b200_last_error());
}

} // namespace

B200OfdmChain::B200OfdmChain(mod_settings_t& s, const std::string& format, int device, bool fixedPoint,
                             int pipelineDepth, int maxBatch) :
    ModCodec(),
    RemoteControllable("b200chain"),
    m_settings(s),
    m_tii(*this),
    m_depth(pipelineDepth > 0 ? (size_t)pipelineDepth : 0)
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
    if (fixedPoint || s.fftEngine == FFTEngine::KISS || static_cast<int>(s.fftEngine) == 4 /* b200_fixed */ ||
        static_cast<int>(s.fftEngine) == 6 /* b200_eti_fixed */) c.fft_engine = DABMOD_B200_FFT_KISS_FIXED;

    std::vector<float> taps, coefs;
    if (!s.filterTapsFilename.empty()) {
        taps = b200files::load_taps(s.filterTapsFilename);
        c.fir_ntaps = (int32_t)taps.size();
        c.fir_taps = taps.data();
    }
    if (!s.polyCoefFilename.empty()) {
        c.dpd_mode = b200files::load_coefs(s.polyCoefFilename, coefs);
        c.dpd_coefs = coefs.data();
    }
    c.format = b200files::format_code(format);
    c.max_batch = std::max<int32_t>(m_depth > 0 ? (int32_t)m_depth : 1, maxBatch);

    if (dabmod_b200_create(&c, &m_handle) != DABMOD_B200_OK) fail("create");
    if (m_depth > 0) {
        m_in.resize(m_depth * dabmod_b200_tf_in_bytes(m_handle));
        m_ready.resize(m_depth * dabmod_b200_tf_out_bytes(m_handle));
    }

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
    const size_t in_tf = dabmod_b200_tf_in_bytes(m_handle), out_tf = dabmod_b200_tf_out_bytes(m_handle);
    dataOut->setLength(out_tf);
    if (m_depth > 0) {
        /* call i: TF i in, TF i - D out; every D-th call turns the D collected TFs into one batch */
        const size_t slot = m_calls % m_depth;
        const bool primed = m_calls >= m_depth;
        if (primed) memcpy(dataOut->getData(), m_ready.data() + slot * out_tf, out_tf);
        /* before the first result: an empty buffer and 0, which ends this flowgraph iteration
         * (src/Flowgraph.cpp:331-336) -- what a PipelinedModCodec's first call amounts to (its input was
         * swapped away before the output is sized from it, src/ModPlugin.cpp:96-109) */
        else dataOut->setLength(0);
        memcpy(m_in.data() + slot * in_tf, dataIn->getData(), in_tf);
        m_calls++;
        if (slot == m_depth - 1) {
            size_t nb = 0;
            if (dabmod_b200_process_batch(m_handle, m_in.data(), m_depth, m_ready.data(), m_ready.size(), &nb) !=
                DABMOD_B200_OK) {
                fail("process_batch");
            }
        }
        return primed ? (int)out_tf : 0;
    }
    size_t n = 0;
    if (dabmod_b200_process(m_handle, reinterpret_cast<const uint8_t*>(dataIn->getData()), dataIn->getLength(),
                            dataOut->getData(), dataOut->getLength(), &n) != DABMOD_B200_OK) {
        fail("process");
    }
    return (int)n;
}

meta_vec_t B200OfdmChain::process_metadata(const meta_vec_t& metadataIn)
{
    if (m_depth == 0) return metadataIn;
    /* PipelinedModCodec::process_metadata (src/ModPlugin.cpp:117-128) with a FIFO of D + 1 instead of 2 */
    m_metadata_fifo.push_back(metadataIn);
    if (m_metadata_fifo.size() == m_depth + 1) {
        auto r = std::move(m_metadata_fifo.front());
        m_metadata_fifo.pop_front();
        return r;
    }
    return {};
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
    write_back(parameter);
}

/* the accepted value goes into DabModulator's settings, where a rebuilt chain will find it */
void B200OfdmChain::write_back(const std::string& parameter)
{
    const std::string v = get_parameter(parameter);
    std::stringstream ss(v);
    if (parameter == "digital") ss >> m_settings.digitalgain;
    else if (parameter == "var") ss >> m_settings.gainmodeVariance;
    else if (parameter == "mode")
        m_settings.gainMode = v == "fix" ? GainMode::GAIN_FIX : v == "max" ? GainMode::GAIN_MAX : GainMode::GAIN_VAR;
    else if (parameter == "windowlen") ss >> m_settings.ofdmWindowOverlap;
    else if (parameter == "cfr") { int b = 0; ss >> b; m_settings.enableCfr = b != 0; }
    else if (parameter == "clip") ss >> m_settings.cfrClip;
    else if (parameter == "errorclip") ss >> m_settings.cfrErrorClip;
    else if (parameter == "tii.enable") { int b = 0; ss >> b; m_settings.tiiConfig.enable = b != 0; }
    else if (parameter == "tii.comb") ss >> m_settings.tiiConfig.comb;
    else if (parameter == "tii.pattern") ss >> m_settings.tiiConfig.pattern;
    else if (parameter == "tii.old_variant") { int b = 0; ss >> b; m_settings.tiiConfig.old_variant = b != 0; }
}

B200TiiControl::B200TiiControl(B200OfdmChain& chain) : RemoteControllable("tii"), m_chain(chain)
{
    /* src/TII.cpp:120-127 */
    m_parameters.push_back({"enable", "enable TII [0-1]"});
    m_parameters.push_back({"comb", "TII comb number [0-23]"});
    m_parameters.push_back({"pattern", "TII pattern number [0-69]"});
    m_parameters.push_back({"old_variant", "select old TII variant for old (buggy) receivers: 0 = off, 1 = on"});
}

void B200TiiControl::set_parameter(const std::string& parameter, const std::string& value)
{
    m_chain.set_parameter("tii." + parameter, value);
}

const std::string B200TiiControl::get_parameter(const std::string& parameter) const
{
    return m_chain.get_parameter("tii." + parameter);
}

const json::map_t B200TiiControl::get_all_values() const
{
    json::map_t map;
    for (const char* p : {"enable", "comb", "pattern", "old_variant"}) map[p].v = get_parameter(p);
    return map;
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
    for (const char* p : {"digital", "mode", "var", "tii.enable", "tii.comb", "tii.pattern", "tii.old_variant",
                          "windowlen", "cfr", "clip", "errorclip", "clip_stats", "papr"}) {
        map[p].v = get_parameter(p);
    }
    return map;
}
