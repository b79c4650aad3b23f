/* B200Blocks -- boundary shape "B1" of SURVEY.md section 8(b): translation-unit substitution.
 *
 * This one file REPLACES the reference's translation units
 *   QpskSymbolMapper.cpp FrequencyInterleaver.cpp PhaseReference.cpp DifferentialModulator.cpp NullSymbol.cpp
 *   TII.cpp SignalMultiplexer.cpp CicEqualizer.cpp OfdmGenerator.cpp GainControl.cpp GuardIntervalInserter.cpp
 *   FIRFilter.cpp Resampler.cpp MemlessPoly.cpp FormatConverter.cpp OutputMemory.cpp
 * at link time.  It defines the SAME classes against the UNMODIFIED reference headers, so DabModulator.cpp,
 * ConfigParser.cpp, DabMod.cpp, Flowgraph.cpp, ModPlugin.cpp and every header stay byte-identical: DabModulator
 * constructs and wires its blocks exactly as it always does (src/DabModulator.cpp:131-417), selects them with the
 * engines it always had (fft_engine = fftw | kiss), and the remote control sees the same controllables with the
 * same parameters ("gain", "tii", "ofdm", "guardinterval", "firfilter", "memlesspoly").
 *
 * None of the blocks computes.  Each process() sets the output length the reference block would produce (downstream
 * blocks and the flowgraph only look at lengths) and forwards a 24-byte TOKEN in the first bytes of the buffer --
 * {magic, chain id, sequence number} -- so that the token experiences exactly the swaps and one-call delays of the
 * PipelinedModCodecs (GainControl, FIRFilter, MemlessPoly: src/ModPlugin.cpp:90-115, unmodified).  The first block,
 * QpskSymbolMapper, keeps the transmission frame's bytes (what BlockPartitioner produced) under the sequence number;
 * the last one, OutputMemory, finds them again through the token and runs the whole chain on the GPU with ONE
 * dabmod_b200_process() call (include/dabmod_b200.h), straight into DabModulator's output buffer.  The output file is
 * therefore aligned like the reference's, including the transmission frames its pipelined blocks never flush.
 * (A frame runs when its token arrives at OutputMemory: the CFR / PAPR read-outs of "ofdm" cover the frames that have
 * left the chain, i.e. they trail the reference's by the 1-3 calls of pipeline delay.)
 *
 * The graph describes itself through the same token: QpskSymbolMapper opens a Chain record on its first frame, and
 * every block the token passes through deposits what its constructor was given (the pieces of mod_settings_t
 * DabModulator hands to it, the references included; TII, which is not on the token's way, leaves a note in its
 * output that SignalMultiplexer picks up).  When the first token reaches OutputMemory the record is complete and the
 * handle is created from it.  Nothing depends on the order in which the blocks were constructed or wired.  Remote
 * control writes go to the referenced settings (so they survive a modulator restart like the reference's) and to
 * the handle.
 *
 * Host glue only; every sample comes from the CUDA kernels behind the C ABI.
 */
#include <algorithm>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>

#include "B200Files.h"
#include "CicEqualizer.h"
#include "DifferentialModulator.h"
#include "FIRFilter.h"
#include "FormatConverter.h"
#include "FrequencyInterleaver.h"
#include "GainControl.h"
#include "GuardIntervalInserter.h"
#include "Log.h"
#include "MemlessPoly.h"
#include "NullSymbol.h"
#include "OfdmGenerator.h"
#include "OutputMemory.h"
#include "PhaseReference.h"
#include "QpskSymbolMapper.h"
#include "Resampler.h"
#include "SignalMultiplexer.h"
#include "TII.h"
#include "dabmod_b200.h"

namespace {

constexpr uint64_t TOKEN_MAGIC = 0xb200c0fdb10c0001ull;

struct Token {
    uint64_t magic, chain, seq;
};

/* what one DabModulator graph's constructors were given */
struct Chain {
    uint64_t id = 0;
    unsigned mode = 0;
    bool fixed = false;
    bool cic = false;
    size_t cic_ratio = 1;
    tii_config_t* tii = nullptr;
    bool* cfr = nullptr;
    float *cfr_clip = nullptr, *cfr_errclip = nullptr;
    GainMode* gain_mode = nullptr;
    float *digital = nullptr, *variance = nullptr;
    float normalise = 1.0f;
    size_t* window = nullptr;
    std::vector<float> taps;
    size_t out_rate = 2048000;
    int dpd_mode = DABMOD_B200_DPD_NONE;
    std::vector<float> coefs;
    std::string format;

    std::mutex mtx;
    dabmod_b200* handle = nullptr;
    uint64_t next_seq = 0;
    std::map<uint64_t, std::vector<uint8_t>> frames;   /* sequence number -> BlockPartitioner bytes */

    ~Chain() { dabmod_b200_destroy(handle); }
};

std::mutex g_mtx;
std::map<uint64_t, std::shared_ptr<Chain>> g_chains;     /* live graphs by id */
std::map<const void*, std::shared_ptr<Chain>> g_of_block;   /* the graph a block was last seen in */
uint64_t g_next_id = 1;

/* the first block of the path opens the record of its graph (once per QpskSymbolMapper instance) */
std::shared_ptr<Chain> chain_of_source(const void* block)
{
    std::lock_guard<std::mutex> lock(g_mtx);
    auto it = g_of_block.find(block);
    if (it != g_of_block.end() && g_chains.count(it->second->id)) return it->second;
    auto c = std::make_shared<Chain>();
    c->id = g_next_id++;
    g_chains[c->id] = c;
    g_of_block[block] = c;
    return c;
}

/* the graph a block is part of, or nullptr before its first frame */
std::shared_ptr<Chain> chain_of(const void* block)
{
    std::lock_guard<std::mutex> lock(g_mtx);
    auto it = g_of_block.find(block);
    return it == g_of_block.end() || !g_chains.count(it->second->id) ? nullptr : it->second;
}

std::shared_ptr<Chain> chain_by_id(uint64_t id)
{
    std::lock_guard<std::mutex> lock(g_mtx);
    auto it = g_chains.find(id);
    return it == g_chains.end() ? nullptr : it->second;
}

void forget(const void* block)
{
    std::lock_guard<std::mutex> lock(g_mtx);
    g_of_block.erase(block);
}

/* sizes: every buffer on the path is at least a token long */
void put_token(Buffer* out, size_t len, const Token& t)
{
    out->setLength(std::max(len, sizeof(Token)));
    memcpy(out->getData(), &t, sizeof(t));
    out->setLength(len);                                 /* the length the reference block reports */
}

Token get_token(const Buffer* in, const char* who)
{
    Token t{};
    /* Buffer keeps its storage when it shrinks (src/Buffer.cpp:128-147), the token is there whatever the length */
    if (in->getData() == nullptr) throw std::runtime_error(std::string(who) + ": empty input buffer");
    memcpy(&t, in->getData(), sizeof(t));
    if (t.magic != TOKEN_MAGIC) throw std::runtime_error(std::string(who) + ": input does not come from the B200 chain");
    return t;
}

/* Forwards the token and hands `deposit` the graph's record while its handle does not exist yet (the first frames),
 * so that the block can leave what its constructor was given. */
template <typename F> void pass(const void* block, const Buffer* in, Buffer* out, size_t len, const char* who, F deposit)
{
    const Token t = get_token(in, who);
    put_token(out, len, t);
    auto c = chain_by_id(t.chain);
    if (!c) return;
    {
        std::lock_guard<std::mutex> lock(c->mtx);
        if (c->handle) return;
        deposit(*c);
    }
    std::lock_guard<std::mutex> lock(g_mtx);
    g_of_block[block] = c;
}

void pass(const void* block, const Buffer* in, Buffer* out, size_t len, const char* who)
{
    pass(block, in, out, len, who, [](Chain&) {});
}

/* what TII leaves in its output for SignalMultiplexer */
constexpr uint64_t TII_MAGIC = 0xb200c0fdb10c0002ull;
struct TiiNote {
    uint64_t magic;
    tii_config_t* conf;
    const void* block;
};

[[noreturn]] void not_exported(const std::string& parameter, const std::string& rc_name)
{
    throw ParameterError("Parameter '" + parameter + "' is not exported by controllable " + rc_name);
}

/* remote control: the handle (when it exists) validates and applies, the caller then stores into the settings */
void handle_set(const std::shared_ptr<Chain>& c, const char* name, const std::string& value)
{
    if (!c) return;                                      /* no frame yet: the settings are all there is */
    std::lock_guard<std::mutex> lock(c->mtx);
    if (c->handle && dabmod_b200_set_param(c->handle, name, value.c_str()) != DABMOD_B200_OK) {
        throw ParameterError(dabmod_b200_last_error());
    }
}

bool handle_get(const std::shared_ptr<Chain>& c, const char* name, std::string& value)
{
    if (!c) return false;
    std::lock_guard<std::mutex> lock(c->mtx);
    char buf[512];
    if (!c->handle || dabmod_b200_get_param(c->handle, name, buf, sizeof(buf)) != DABMOD_B200_OK) return false;
    value = buf;
    return true;
}

template <typename T> T parse(const std::string& value)
{
    std::stringstream ss(value);
    T v{};
    ss >> v;
    if (ss.fail()) throw ParameterError("cannot parse '" + value + "'");
    return v;
}

/* the handle, from what the constructors were given; caller holds c.mtx */
void create_handle(Chain& c)
{
    dabmod_b200_config cfg;
    dabmod_b200_config_init(&cfg);
    cfg.mode = (int32_t)c.mode;
    cfg.max_batch = 1;
    cfg.output_rate = c.out_rate;
    if (c.fixed) cfg.fft_engine = DABMOD_B200_FFT_KISS_FIXED;
    /* DabModulator derived the CIC ratio from clockRate / outputRate / 4 (src/DabModulator.cpp:154-168); a clock
     * that reproduces it (400 MHz is the one value the library treats specially, and the reference only builds the
     * equaliser there when the ratio is odd) */
    if (c.cic) cfg.clock_rate = (uint64_t)c.cic_ratio * 4u * c.out_rate;
    if (c.gain_mode) cfg.gain_mode = (int32_t)*c.gain_mode;
    if (c.digital) cfg.digital_gain = *c.digital;
    if (c.variance) cfg.gain_variance = *c.variance;
    cfg.normalise = c.normalise;
    if (c.window) cfg.window_overlap = (int32_t)*c.window;
    if (c.cfr) {
        cfg.cfr_enable = *c.cfr;
        cfg.cfr_clip = *c.cfr_clip;
        cfg.cfr_errclip = *c.cfr_errclip;
    }
    if (c.tii) {
        cfg.tii_enable = c.tii->enable;
        cfg.tii_comb = c.tii->comb;
        cfg.tii_pattern = c.tii->pattern;
        cfg.tii_old_variant = c.tii->old_variant;
    }
    if (!c.taps.empty()) {
        cfg.fir_ntaps = (int32_t)c.taps.size();
        cfg.fir_taps = c.taps.data();
    }
    if (c.dpd_mode != DABMOD_B200_DPD_NONE) {
        cfg.dpd_mode = c.dpd_mode;
        cfg.dpd_coefs = c.coefs.data();
    }
    cfg.format = b200files::format_code(c.format);
    if (dabmod_b200_create(&cfg, &c.handle) != DABMOD_B200_OK) {
        throw std::runtime_error(std::string("B200Blocks: ") + dabmod_b200_last_error());
    }
}

size_t carriers_of(unsigned mode)
{
    switch (mode) {
        case 1: return 1536;
        case 2: return 384;
        case 3: return 192;
        case 4: return 768;
        default: throw std::runtime_error("invalid DAB mode");
    }
}

} // namespace

/* ------------------------------------------------------------------ QpskSymbolMapper (src/QpskSymbolMapper.cpp) */
QpskSymbolMapper::QpskSymbolMapper(size_t carriers, bool fixedPoint) : ModCodec(), m_fixedPoint(fixedPoint), m_carriers(carriers)
{
}

int QpskSymbolMapper::process(Buffer* const dataIn, Buffer* dataOut)
{
    auto c = chain_of_source(this);
    /* the reference's size rule: whole symbols of carriers * 2 bits */
    if (dataIn->getLength() == 0 || dataIn->getLength() % (m_carriers / 4) != 0) {
        throw std::runtime_error("QpskSymbolMapper::process input size not valid: " + std::to_string(dataIn->getLength()) +
                                 "(input size) % (" + std::to_string(m_carriers) + " (carriers) / 4) != 0");
    }
    Token t{TOKEN_MAGIC, c->id, 0};
    {
        std::lock_guard<std::mutex> lock(c->mtx);
        if (!c->handle) {
            c->fixed = m_fixedPoint;
            c->mode = m_carriers == 1536 ? 1 : m_carriers == 384 ? 2 : m_carriers == 192 ? 3 : m_carriers == 768 ? 4 : 0;
        }
        t.seq = c->next_seq++;
        const uint8_t* p = reinterpret_cast<const uint8_t*>(dataIn->getData());
        c->frames[t.seq].assign(p, p + dataIn->getLength());
        /* the pipelined blocks hold at most one frame each; anything older was dropped by a flowgraph stop */
        while (c->frames.size() > 8) c->frames.erase(c->frames.begin());
    }
    /* four carriers per input byte */
    put_token(dataOut, dataIn->getLength() * 4 * (m_fixedPoint ? sizeof(complexfix) : sizeof(complexf)), t);
    return 1;
}

/* ---------------------------------------------------------- FrequencyInterleaver (src/FrequencyInterleaver.cpp) */
FrequencyInterleaver::FrequencyInterleaver(size_t mode, bool fixedPoint) :
    ModCodec(), m_fixedPoint(fixedPoint), m_carriers(carriers_of((unsigned)mode)), m_indices(nullptr)
{
}

FrequencyInterleaver::~FrequencyInterleaver() { forget(this); }

int FrequencyInterleaver::process(Buffer* const dataIn, Buffer* dataOut)
{
    pass(this, dataIn, dataOut, dataIn->getLength(), "FrequencyInterleaver::process");
    return 1;
}

/* ------------------------------------------------------------------------ PhaseReference (src/PhaseReference.cpp) */
PhaseReference::PhaseReference(unsigned int dabmode, bool fixedPoint) :
    ModInput(), d_dabmode(dabmode), d_fixedPoint(fixedPoint), d_carriers(carriers_of(dabmode))
{
}

int PhaseReference::process(Buffer* dataOut)
{
    if (dataOut == nullptr) throw std::runtime_error("PhaseReference::process received a NULL buffer");
    dataOut->setLength(d_carriers * (d_fixedPoint ? sizeof(complexfix) : sizeof(complexf)));
    return 1;
}

/* --------------------------------------------------------- DifferentialModulator (src/DifferentialModulator.cpp) */
DifferentialModulator::DifferentialModulator(size_t carriers, bool fixedPoint) : ModMux(), m_carriers(carriers), m_fixedPoint(fixedPoint)
{
}

DifferentialModulator::~DifferentialModulator() { forget(this); }

/* dataIn[0] = phase reference symbol, dataIn[1] = the interleaved data symbols (src/DabModulator.cpp:388-389) */
int DifferentialModulator::process(std::vector<Buffer*> dataIn, Buffer* dataOut)
{
    if (dataIn.size() != 2) throw std::runtime_error("DifferentialModulator::process nb of input streams not 2!");
    pass(this, dataIn[1], dataOut, dataIn[0]->getLength() + dataIn[1]->getLength(), "DifferentialModulator::process");
    return 1;
}

/* ---------------------------------------------------------------------------------- NullSymbol (src/NullSymbol.cpp) */
NullSymbol::NullSymbol(size_t numCarriers, size_t typeSize) : ModInput(), m_numCarriers(numCarriers), m_typeSize(typeSize)
{
}

NullSymbol::~NullSymbol() { forget(this); }

int NullSymbol::process(Buffer* dataOut)
{
    dataOut->setLength(m_numCarriers * m_typeSize);
    return 1;
}

/* ------------------------------------------------------------------------------------------------ TII (src/TII.cpp) */
TII::TII(unsigned int dabmode, tii_config_t& tii_config, bool fixedPoint) :
    ModCodec(), RemoteControllable("tii"), m_dabmode(dabmode), m_conf(tii_config), m_fixedPoint(fixedPoint)
{
    RC_ADD_PARAMETER(enable, "enable TII [0-1]");
    RC_ADD_PARAMETER(comb, "TII comb number [0-23]");
    RC_ADD_PARAMETER(pattern, "TII pattern number [0-69]");
    RC_ADD_PARAMETER(old_variant, "select old TII variant for old (buggy) receivers [0-1]");
    /* the reference's constructor rejects what EN 300 401 clause 14.8 has no table for; DabModulator catches
     * TIIError and runs without TII (src/DabModulator.cpp:180-192) */
    if (dabmode != 1 && dabmode != 2) {
        throw TIIError("TII::TII invalid DAB mode " + std::to_string(dabmode));
    }
    if (m_conf.pattern < 0 || m_conf.pattern > 69) throw TIIError("TII::TII pattern not valid!");
    if (m_conf.comb < 0 || m_conf.comb > 23) throw TIIError("TII::TII comb not valid!");
    m_carriers = carriers_of(dabmode);
}

const char* TII::name()
{
    std::stringstream ss;
    ss << "TII(c:" << m_conf.comb << " p:" << m_conf.pattern << " vrnt:" << (m_conf.old_variant ? "old" : "new") << ")";
    m_name = ss.str();
    return m_name.c_str();
}

int TII::process(Buffer* dataIn, Buffer* dataOut)
{
    if (dataIn == nullptr || dataOut == nullptr) throw TIIError("TII::process received a NULL buffer");
    dataOut->setLength(m_carriers * (m_fixedPoint ? sizeof(complexfix) : sizeof(complexf)));
    /* TII is not on the token's way (its input is a phase reference): SignalMultiplexer picks this note up */
    const TiiNote note{TII_MAGIC, &m_conf, this};
    memcpy(dataOut->getData(), &note, sizeof(note));
    return 1;
}

void TII::enable_carrier(int) {}
void TII::prepare_pattern() {}

void TII::set_parameter(const std::string& parameter, const std::string& value)
{
    auto c = chain_of(this);
    if (parameter == "enable") {
        const int v = parse<int>(value);
        handle_set(c, "tii.enable", value);
        m_conf.enable = v != 0;
    }
    else if (parameter == "pattern") {
        const int v = parse<int>(value);
        if (v < 0 || v > 69) throw ParameterError("TII pattern not valid!");
        handle_set(c, "tii.pattern", value);
        m_conf.pattern = v;
    }
    else if (parameter == "comb") {
        const int v = parse<int>(value);
        if (v < 0 || v > 23) throw ParameterError("TII comb not valid!");
        handle_set(c, "tii.comb", value);
        m_conf.comb = v;
    }
    else if (parameter == "old_variant") {
        const int v = parse<int>(value);
        handle_set(c, "tii.old_variant", value);
        m_conf.old_variant = v != 0;
    }
    else not_exported(parameter, get_rc_name());
}

const std::string TII::get_parameter(const std::string& parameter) const
{
    std::stringstream ss;
    if (parameter == "enable") ss << (m_conf.enable ? 1 : 0);
    else if (parameter == "pattern") ss << m_conf.pattern;
    else if (parameter == "comb") ss << m_conf.comb;
    else if (parameter == "old_variant") ss << (m_conf.old_variant ? 1 : 0);
    else not_exported(parameter, get_rc_name());
    return ss.str();
}

const json::map_t TII::get_all_values() const
{
    json::map_t map;
    for (const char* p : {"enable", "pattern", "comb", "old_variant"}) map[p].v = get_parameter(p);
    return map;
}

/* ------------------------------------------------------------------- SignalMultiplexer (src/SignalMultiplexer.cpp) */
SignalMultiplexer::SignalMultiplexer() : ModMux() {}
SignalMultiplexer::~SignalMultiplexer() {}

/* dataIn[0] = null symbol, dataIn[1] = the data symbols, dataIn[2] = (optional) the TII symbol */
int SignalMultiplexer::process(std::vector<Buffer*> dataIn, Buffer* dataOut)
{
    if (dataIn.size() != 2 && dataIn.size() != 3) throw std::runtime_error("SignalMultiplexer::process needs 2 or 3 inputs");
    const size_t null_len = dataIn[dataIn.size() == 3 ? 2 : 0]->getLength();
    const Buffer* tii = dataIn.size() == 3 ? dataIn[2] : nullptr;
    const void* tii_block = nullptr;
    pass(this, dataIn[1], dataOut, null_len + dataIn[1]->getLength(), "SignalMultiplexer::process", [&](Chain& c) {
        TiiNote note{};
        if (tii && tii->getData() && tii->getLength() >= sizeof(note)) memcpy(&note, tii->getData(), sizeof(note));
        if (note.magic == TII_MAGIC) {
            c.tii = note.conf;
            tii_block = note.block;
        }
    });
    if (tii_block) {
        auto c = chain_of(this);
        std::lock_guard<std::mutex> lock(g_mtx);
        if (c) g_of_block[tii_block] = c;
    }
    return 1;
}

/* ---------------------------------------------------------------------------- CicEqualizer (src/CicEqualizer.cpp) */
CicEqualizer::CicEqualizer(size_t nbCarriers, size_t spacing, int R) :
    ModCodec(), myNbCarriers(nbCarriers), mySpacing(spacing), myFilter(1, (float)std::max(R, 1))   /* the ratio, kept */
{
}

CicEqualizer::~CicEqualizer() { forget(this); }

int CicEqualizer::process(Buffer* const dataIn, Buffer* dataOut)
{
    pass(this, dataIn, dataOut, dataIn->getLength(), "CicEqualizer::process", [&](Chain& c) {
        c.cic = true;
        c.cic_ratio = (size_t)myFilter[0];
    });
    return 1;
}

/* -------------------------------------------------------------------------- OfdmGenerator (src/OfdmGenerator.cpp) */
OfdmGeneratorCF32::OfdmGeneratorCF32(size_t nbSymbols, size_t nbCarriers, size_t spacing, bool& enableCfr, float& cfrClip,
                                     float& cfrErrorClip, bool inverse) :
    ModCodec(), RemoteControllable("ofdm"),
    myFftPlan(nullptr), myFftIn(nullptr), myFftOut(nullptr),
    myNbSymbols(nbSymbols), myNbCarriers(nbCarriers), mySpacing(spacing),
    myCfr(enableCfr), myCfrClip(cfrClip), myCfrErrorClip(cfrErrorClip),
    myCfrFft(nullptr), myCfrPostClip(nullptr), myCfrPostFft(nullptr),
    myPaprBeforeCFR(1), myPaprAfterCFR(1), myPaprClearRequest(false)
{
    if (!inverse) throw std::runtime_error("OfdmGenerator: only the inverse transform of the modulator is accelerated");
    if (nbCarriers > spacing) throw std::runtime_error("OfdmGenerator nbCarriers > spacing!");
    RC_ADD_PARAMETER(cfr, "Enable crest factor reduction");
    RC_ADD_PARAMETER(clip, "CFR: Clip to amplitude");
    RC_ADD_PARAMETER(errorclip, "CFR: Limit error");
    RC_ADD_PARAMETER(clip_stats, "CFR: statistics (clip ratio, errorclip ratio)");
    RC_ADD_PARAMETER(papr, "PAPR measurements (before CFR, after CFR)");
}

OfdmGeneratorCF32::~OfdmGeneratorCF32() { forget(this); }

int OfdmGeneratorCF32::process(Buffer* const dataIn, Buffer* dataOut)
{
    if (dataIn->getLength() != myNbSymbols * myNbCarriers * sizeof(complexf)) {
        throw std::runtime_error("OfdmGenerator::process input size not valid! IN " + std::to_string(dataIn->getLength()) +
                                 " != " + std::to_string(myNbSymbols * myNbCarriers * sizeof(complexf)));
    }
    pass(this, dataIn, dataOut, myNbSymbols * mySpacing * sizeof(complexf), "OfdmGenerator::process", [&](Chain& c) {
        c.cfr = &myCfr;
        c.cfr_clip = &myCfrClip;
        c.cfr_errclip = &myCfrErrorClip;
    });
    return 1;
}

OfdmGeneratorCF32::cfr_iter_stat_t OfdmGeneratorCF32::cfr_one_iteration(complexf*, const complexf*) { return {}; }

void OfdmGeneratorCF32::set_parameter(const std::string& parameter, const std::string& value)
{
    auto c = chain_of(this);
    if (parameter == "cfr") {
        const int v = parse<int>(value);
        handle_set(c, "cfr", value);
        std::lock_guard<std::mutex> lock(myCfrRcMutex);
        myCfr = v != 0;
    }
    else if (parameter == "clip") {
        const float v = parse<float>(value);
        handle_set(c, "clip", value);
        std::lock_guard<std::mutex> lock(myCfrRcMutex);
        myCfrClip = v;
    }
    else if (parameter == "errorclip") {
        const float v = parse<float>(value);
        handle_set(c, "errorclip", value);
        std::lock_guard<std::mutex> lock(myCfrRcMutex);
        myCfrErrorClip = v;
    }
    else if (parameter == "clip_stats" || parameter == "papr") {
        throw ParameterError("Parameter '" + parameter + "' is read-only");
    }
    else not_exported(parameter, get_rc_name());
}

const std::string OfdmGeneratorCF32::get_parameter(const std::string& parameter) const
{
    std::stringstream ss;
    std::string v;
    if (parameter == "cfr") ss << myCfr;
    else if (parameter == "clip") ss << std::fixed << myCfrClip;
    else if (parameter == "errorclip") ss << std::fixed << myCfrErrorClip;
    else if (parameter == "clip_stats" || parameter == "papr") {
        /* the device-side per-symbol records, aggregated into the reference's strings by the library */
        auto c = chain_of(this);
        if (handle_get(c, parameter.c_str(), v)) ss << v;
        else ss << (parameter == "papr" ? "0 0" : "No stats available");
    }
    else not_exported(parameter, get_rc_name());
    return ss.str();
}

const json::map_t OfdmGeneratorCF32::get_all_values() const
{
    json::map_t map;
    for (const char* p : {"cfr", "clip", "errorclip", "clip_stats", "papr"}) map[p].v = get_parameter(p);
    return map;
}

OfdmGeneratorFixed::OfdmGeneratorFixed(size_t nbSymbols, size_t nbCarriers, size_t spacing, bool inverse) :
    ModCodec(), myFftIn(nullptr), myFftOut(nullptr), myNbSymbols(nbSymbols), myNbCarriers(nbCarriers), mySpacing(spacing)
{
    if (!inverse) throw std::runtime_error("OfdmGenerator: only the inverse transform of the modulator is accelerated");
    if (nbCarriers > spacing) throw std::runtime_error("OfdmGenerator nbCarriers > spacing!");
}

OfdmGeneratorFixed::~OfdmGeneratorFixed() { forget(this); }

int OfdmGeneratorFixed::process(Buffer* const dataIn, Buffer* dataOut)
{
    if (dataIn->getLength() != myNbSymbols * myNbCarriers * sizeof(complexfix)) {
        throw std::runtime_error("OfdmGenerator::process input size not valid!");
    }
    pass(this, dataIn, dataOut, myNbSymbols * mySpacing * sizeof(complexfix), "OfdmGenerator::process", [](Chain& c) { c.fixed = true; });
    return 1;
}

/* ------------------------------------------------------------------------------ GainControl (src/GainControl.cpp) */
GainControl::GainControl(size_t framesize, GainMode& gainMode, float& digGain, float normalise, float& varVariance) :
    PipelinedModCodec(), RemoteControllable("gain"),
    m_frameSize(framesize), m_digGain(digGain), m_normalise(normalise), m_var_variance_rc(varVariance), m_gainmode(gainMode)
{
    RC_ADD_PARAMETER(digital, "Digital Gain");
    RC_ADD_PARAMETER(mode, "Gainmode (fix|max|var)");
    RC_ADD_PARAMETER(var, "Variance setting for gainmode var (default: 4)");
    start_pipeline_thread();
}

GainControl::~GainControl()
{
    stop_pipeline_thread();
    forget(this);
}

int GainControl::internal_process(Buffer* const dataIn, Buffer* dataOut)
{
    pass(this, dataIn, dataOut, dataIn->getLength(), "GainControl::internal_process", [&](Chain& c) {
        c.gain_mode = &m_gainmode;
        c.digital = &m_digGain;
        c.variance = &m_var_variance_rc;
        c.normalise = m_normalise;
    });
    return 1;
}

void GainControl::set_parameter(const std::string& parameter, const std::string& value)
{
    auto c = chain_of(this);
    if (parameter == "digital") {
        const float v = parse<float>(value);
        handle_set(c, "digital", value);
        m_digGain = v;
    }
    else if (parameter == "mode") {
        std::string m = parse<std::string>(value);
        std::transform(m.begin(), m.end(), m.begin(), [](char ch) { return (char)std::tolower(ch); });
        if (m != "fix" && m != "max" && m != "var") throw ParameterError("Gainmode " + m + " unknown");
        handle_set(c, "mode", m);
        std::lock_guard<std::mutex> lock(m_mutex);
        m_gainmode = m == "fix" ? GainMode::GAIN_FIX : m == "max" ? GainMode::GAIN_MAX : GainMode::GAIN_VAR;
    }
    else if (parameter == "var") {
        const float v = parse<float>(value);
        handle_set(c, "var", value);
        std::lock_guard<std::mutex> lock(m_mutex);
        m_var_variance_rc = v;
    }
    else not_exported(parameter, get_rc_name());
}

const std::string GainControl::get_parameter(const std::string& parameter) const
{
    std::stringstream ss;
    if (parameter == "digital") ss << std::fixed << m_digGain;
    else if (parameter == "mode") ss << (m_gainmode == GainMode::GAIN_FIX ? "fix" : m_gainmode == GainMode::GAIN_MAX ? "max" : "var");
    else if (parameter == "var") ss << std::fixed << m_var_variance_rc;
    else not_exported(parameter, get_rc_name());
    return ss.str();
}

const json::map_t GainControl::get_all_values() const
{
    json::map_t map;
    for (const char* p : {"digital", "mode", "var"}) map[p].v = get_parameter(p);
    return map;
}

/* -------------------------------------------------------- GuardIntervalInserter (src/GuardIntervalInserter.cpp) */
GuardIntervalInserter::Params::Params(size_t nbSymbols, size_t spacing, size_t nullSize, size_t symSize, size_t& windowOverlap) :
    nbSymbols(nbSymbols), spacing(spacing), nullSize(nullSize), symSize(symSize), windowOverlap(windowOverlap)
{
}

GuardIntervalInserter::GuardIntervalInserter(size_t nbSymbols, size_t spacing, size_t nullSize, size_t symSize,
                                             size_t& windowOverlap, FFTEngine fftEngine) :
    ModCodec(), RemoteControllable("guardinterval"), m_fftEngine(fftEngine),
    m_params(nbSymbols, spacing, nullSize, symSize, windowOverlap)
{
    if (nullSize == 0) throw std::logic_error("NULL symbol must be present");
    RC_ADD_PARAMETER(windowlen, "Window length for OFDM windowng [0 to disable]");
    /* the reference's rule (update_window): the window may not exceed the guard interval */
    update_window(windowOverlap);
}

void GuardIntervalInserter::update_window(size_t new_window_overlap)
{
    if (new_window_overlap > m_params.symSize - m_params.spacing) {
        throw std::out_of_range("Window overlap too large: " + std::to_string(new_window_overlap));
    }
    std::lock_guard<std::mutex> lock(m_params.windowMutex);
    m_params.windowOverlap = new_window_overlap;
}

int GuardIntervalInserter::process(Buffer* const dataIn, Buffer* dataOut)
{
    const size_t sample = m_fftEngine == FFTEngine::FFTW ? sizeof(complexf) : sizeof(complexfix);
    if (dataIn->getLength() != (m_params.nbSymbols + 1) * m_params.spacing * sample) {
        throw std::runtime_error("GuardIntervalInserter::process error on input size");
    }
    pass(this, dataIn, dataOut, (m_params.nullSize + m_params.nbSymbols * m_params.symSize) * sample,
         "GuardIntervalInserter::process", [&](Chain& c) { c.window = &m_params.windowOverlap; });
    return 1;
}

void GuardIntervalInserter::set_parameter(const std::string& parameter, const std::string& value)
{
    if (parameter == "windowlen") {
        const size_t v = parse<size_t>(value);
        const size_t old = m_params.windowOverlap;
        try { update_window(v); }
        catch (const std::out_of_range& e) { throw ParameterError(e.what()); }
        try { handle_set(chain_of(this), "windowlen", value); }
        catch (...) { update_window(old); throw; }
    }
    else not_exported(parameter, get_rc_name());
}

const std::string GuardIntervalInserter::get_parameter(const std::string& parameter) const
{
    std::stringstream ss;
    if (parameter == "windowlen") ss << m_params.windowOverlap;
    else not_exported(parameter, get_rc_name());
    return ss.str();
}

const json::map_t GuardIntervalInserter::get_all_values() const
{
    json::map_t map;
    map["windowlen"].v = get_parameter("windowlen");
    return map;
}

/* ------------------------------------------------------------------------------------ FIRFilter (src/FIRFilter.cpp) */
FIRFilter::FIRFilter(std::string& taps_file) : PipelinedModCodec(), RemoteControllable("firfilter"), m_taps_file(taps_file)
{
    RC_ADD_PARAMETER(ntaps, "(Read-only) number of filter taps.");
    RC_ADD_PARAMETER(tapsfile, "Filename containing filter taps. When written to, the new file gets automatically loaded.");
    load_filter_taps(m_taps_file);
    start_pipeline_thread();
}

FIRFilter::~FIRFilter()
{
    stop_pipeline_thread();
    forget(this);
}

void FIRFilter::load_filter_taps(const std::string& tapsFile)
{
    std::vector<float> taps = b200files::load_taps(tapsFile);
    if (auto c = chain_of(this)) {
        std::lock_guard<std::mutex> lock(c->mtx);
        if (c->handle) {
            std::stringstream ss;
            ss.precision(9);
            ss << taps.size();
            for (float t : taps) ss << " " << t;
            if (dabmod_b200_set_param(c->handle, "taps", ss.str().c_str()) != DABMOD_B200_OK) {
                throw std::runtime_error(dabmod_b200_last_error());
            }
        }
        c->taps = taps;
    }
    std::lock_guard<std::mutex> lock(m_taps_mutex);
    m_taps = std::move(taps);
}

int FIRFilter::internal_process(Buffer* const dataIn, Buffer* dataOut)
{
    pass(this, dataIn, dataOut, dataIn->getLength(), "FIRFilter::internal_process", [&](Chain& c) {
        std::lock_guard<std::mutex> lock(m_taps_mutex);
        c.taps = m_taps;
    });
    return 1;
}

void FIRFilter::set_parameter(const std::string& parameter, const std::string& value)
{
    if (parameter == "ntaps") throw ParameterError("Parameter 'ntaps' is read-only");
    else if (parameter == "tapsfile") {
        try {
            load_filter_taps(value);
            m_taps_file = value;
        }
        catch (const std::runtime_error& e) { throw ParameterError(e.what()); }
    }
    else not_exported(parameter, get_rc_name());
}

const std::string FIRFilter::get_parameter(const std::string& parameter) const
{
    std::stringstream ss;
    if (parameter == "ntaps") {
        std::lock_guard<std::mutex> lock(m_taps_mutex);
        ss << m_taps.size();
    }
    else if (parameter == "tapsfile") ss << m_taps_file;
    else not_exported(parameter, get_rc_name());
    return ss.str();
}

const json::map_t FIRFilter::get_all_values() const
{
    json::map_t map;
    for (const char* p : {"ntaps", "tapsfile"}) map[p].v = get_parameter(p);
    return map;
}

/* ------------------------------------------------------------------------------------ Resampler (src/Resampler.cpp) */
Resampler::Resampler(size_t inputRate, size_t outputRate, size_t resolution) :
    ModCodec(), myFftPlan1(nullptr), myFftPlan2(nullptr), myFftIn(nullptr), myFftOut(nullptr), myBufferIn(nullptr),
    myBufferOut(nullptr), myFront(nullptr), myBack(nullptr), myWindow(nullptr), myFactor(1.0f)
{
    /* the ratio and the FFT sizes of src/Resampler.cpp:51-76, from the library (it computes with the same sizes) */
    int fft_in = 0, fft_out = 0;
    if (dabmod_b200_resampler_sizes(inputRate, outputRate, (int)resolution, &fft_in, &fft_out) != DABMOD_B200_OK) {
        throw std::runtime_error(std::string("Resampler: ") + dabmod_b200_last_error());
    }
    myFftSizeIn = (size_t)fft_in;
    myFftSizeOut = (size_t)fft_out;
    size_t a = inputRate, b = outputRate;
    while (b) { const size_t t = a % b; a = b; b = t; }
    L = outputRate / a;
    M = inputRate / a;
    if (inputRate != 2048000) throw std::runtime_error("Resampler: the modulator resamples from 2 048 000 samples/s");
}

Resampler::~Resampler() { forget(this); }

int Resampler::process(Buffer* const dataIn, Buffer* dataOut)
{
    pass(this, dataIn, dataOut, dataIn->getLength() * L / M, "Resampler::process", [&](Chain& c) { c.out_rate = 2048000 / M * L; });
    return 1;
}

/* -------------------------------------------------------------------------------- MemlessPoly (src/MemlessPoly.cpp) */
MemlessPoly::MemlessPoly(std::string& coefs_file, unsigned int) :
    PipelinedModCodec(), RemoteControllable("memlesspoly"), m_coefs_am(), m_coefs_pm(), m_coefs_file(coefs_file)
{
    RC_ADD_PARAMETER(ncoefs, "(Read-only) number of coefficients.");
    RC_ADD_PARAMETER(coefs, "Predistortion coefficients, same format as file.");
    RC_ADD_PARAMETER(coeffile, "Filename containing coefficients. When set, the file gets loaded.");
    std::ifstream in(coefs_file);
    if (!in) throw std::runtime_error("MemlessPoly: Could not open file with coefs!");
    load_coefficients(in);
    start_pipeline_thread();
}

MemlessPoly::~MemlessPoly()
{
    stop_pipeline_thread();
    forget(this);
}

void MemlessPoly::worker_thread(worker_t*) {}

void MemlessPoly::load_coefficients(std::istream& coefData)
{
    std::stringstream text;
    text << coefData.rdbuf();
    std::vector<float> coefs;
    std::istringstream parse_in(text.str());
    const int mode = b200files::load_coefs(parse_in, coefs);
    if (auto c = chain_of(this)) {
        std::lock_guard<std::mutex> lock(c->mtx);
        if (c->handle && dabmod_b200_set_param(c->handle, "coefs", text.str().c_str()) != DABMOD_B200_OK) {
            throw std::runtime_error(dabmod_b200_last_error());
        }
        c->dpd_mode = mode;
        c->coefs = coefs;
    }
    std::lock_guard<std::mutex> lock(m_coefs_mutex);
    m_dpd_type = mode == DABMOD_B200_DPD_LUT ? dpd_type_t::lookup_table : dpd_type_t::odd_only_poly;
    if (mode == DABMOD_B200_DPD_LUT) {
        m_lut_scalefactor = coefs[0];
        for (size_t i = 0; i < lut_entries; i++) m_lut[i] = coefs[1 + i];    /* real correction factors */
    }
    else {
        m_coefs_am.assign(coefs.begin(), coefs.begin() + 5);
        m_coefs_pm.assign(coefs.begin() + 5, coefs.end());
    }
    m_dpd_settings_valid = true;
}

std::string MemlessPoly::serialise_coefficients() const
{
    /* the file format (src/MemlessPoly.cpp:145-235): 1, count, AM/AM then AM/PM -- or 2, scale factor, 32 factors */
    std::lock_guard<std::mutex> lock(m_coefs_mutex);
    std::stringstream ss;
    ss.precision(9);
    if (m_dpd_type == dpd_type_t::lookup_table) {
        ss << "2\n" << m_lut_scalefactor;
        for (size_t i = 0; i < lut_entries; i++) ss << "\n" << m_lut[i].real();
    }
    else {
        ss << "1\n" << m_coefs_am.size();
        for (float v : m_coefs_am) ss << "\n" << v;
        for (float v : m_coefs_pm) ss << "\n" << v;
    }
    return ss.str();
}

int MemlessPoly::internal_process(Buffer* const dataIn, Buffer* dataOut)
{
    pass(this, dataIn, dataOut, dataIn->getLength(), "MemlessPoly::internal_process", [&](Chain& c) {
        std::lock_guard<std::mutex> lock(m_coefs_mutex);
        c.dpd_mode = m_dpd_type == dpd_type_t::lookup_table ? DABMOD_B200_DPD_LUT : DABMOD_B200_DPD_ODD_POLY;
        c.coefs.clear();
        if (m_dpd_type == dpd_type_t::lookup_table) {
            c.coefs.push_back(m_lut_scalefactor);
            for (const auto& l : m_lut) c.coefs.push_back(l.real());
        }
        else {
            c.coefs.insert(c.coefs.end(), m_coefs_am.begin(), m_coefs_am.end());
            c.coefs.insert(c.coefs.end(), m_coefs_pm.begin(), m_coefs_pm.end());
        }
    });
    return 1;
}

void MemlessPoly::set_parameter(const std::string& parameter, const std::string& value)
{
    if (parameter == "ncoefs") throw ParameterError("Parameter 'ncoefs' is read-only");
    else if (parameter == "coeffile") {
        try {
            std::ifstream in(value);
            if (!in) throw std::runtime_error("MemlessPoly: Could not open file with coefs!");
            load_coefficients(in);
            m_coefs_file = value;
        }
        catch (const std::runtime_error& e) { throw ParameterError(e.what()); }
    }
    else if (parameter == "coefs") {
        try {
            std::istringstream in(value);
            load_coefficients(in);
            m_coefs_file = "<set via RC>";
        }
        catch (const std::runtime_error& e) { throw ParameterError(e.what()); }
    }
    else not_exported(parameter, get_rc_name());
}

const std::string MemlessPoly::get_parameter(const std::string& parameter) const
{
    std::stringstream ss;
    if (parameter == "ncoefs") {
        std::lock_guard<std::mutex> lock(m_coefs_mutex);
        ss << (m_dpd_type == dpd_type_t::lookup_table ? lut_entries : m_coefs_am.size());
    }
    else if (parameter == "coefs") ss << serialise_coefficients();
    else if (parameter == "coeffile") ss << m_coefs_file;
    else not_exported(parameter, get_rc_name());
    return ss.str();
}

const json::map_t MemlessPoly::get_all_values() const
{
    json::map_t map;
    for (const char* p : {"ncoefs", "coefs", "coeffile"}) map[p].v = get_parameter(p);
    return map;
}

/* ------------------------------------------------------------------------ FormatConverter (src/FormatConverter.cpp) */
FormatConverter::FormatConverter(bool input_is_complexfix_wide, const std::string& format_out) :
    ModCodec(), m_input_complexfix_wide(input_is_complexfix_wide), m_format_out(format_out)
{
    if (input_is_complexfix_wide) throw std::runtime_error("FormatConverter: the DEXTER sample format is not accelerated");
    get_format_size(format_out);                          /* throws on an unknown format */
}

FormatConverter::~FormatConverter()
{
    etiLog.level(debug) << "FormatConverter: " << m_num_clipped_samples.load() << " clipped";
    forget(this);
}

int FormatConverter::process(Buffer* const dataIn, Buffer* dataOut)
{
    /* float components in, one converted component each out */
    const size_t components = dataIn->getLength() / sizeof(float);
    pass(this, dataIn, dataOut, components * (get_format_size(m_format_out) / 2), "FormatConverter::process",
         [&](Chain& c) { c.format = m_format_out; });
    /* the count of the frame the GPU converted last (the reference reports the last frame it converted) */
    if (auto c = chain_of(this)) {
        std::lock_guard<std::mutex> lock(c->mtx);
        if (c->handle) m_num_clipped_samples.store(dabmod_b200_num_clipped_samples(c->handle));
    }
    return 1;
}

const char* FormatConverter::name() { return "FormatConverter"; }

size_t FormatConverter::get_num_clipped_samples() const { return m_num_clipped_samples.load(); }

size_t FormatConverter::get_format_size(const std::string& format)
{
    if (format == "s16") return 4;
    if (format == "u8" || format == "s8") return 2;
    throw std::runtime_error("FormatConverter: Invalid format " + format);
}

/* ------------------------------------------------------------------------------ OutputMemory (src/OutputMemory.cpp) */
OutputMemory::OutputMemory(Buffer* dataOut) : ModOutput(), m_dataOut(dataOut)
{
}

OutputMemory::~OutputMemory()
{
    /* the graph ends with its sink: the record, the handle and every block's entry go with it */
    std::shared_ptr<Chain> c;
    {
        std::lock_guard<std::mutex> lock(g_mtx);
        auto it = g_of_block.find(this);
        if (it != g_of_block.end()) {
            c = it->second;
            g_chains.erase(c->id);
            for (auto b = g_of_block.begin(); b != g_of_block.end();) {
                if (b->second == c) b = g_of_block.erase(b);
                else ++b;
            }
        }
    }
}

int OutputMemory::process(Buffer* dataIn)
{
    const Token t = get_token(dataIn, "OutputMemory::process");
    auto c = chain_by_id(t.chain);
    if (!c) throw std::runtime_error("OutputMemory::process: frame of a graph that no longer exists");
    {
        std::lock_guard<std::mutex> lock(g_mtx);
        g_of_block[this] = c;
    }
    std::lock_guard<std::mutex> lock(c->mtx);
    auto it = c->frames.find(t.seq);
    if (it == c->frames.end()) throw std::runtime_error("OutputMemory::process: transmission frame lost on the way");
    if (!c->handle) create_handle(*c);
    const size_t out_bytes = dabmod_b200_tf_out_bytes(c->handle);
    if (out_bytes != dataIn->getLength()) {
        throw std::runtime_error("OutputMemory::process: the chain announced " + std::to_string(dataIn->getLength()) +
                                 " bytes, the GPU produces " + std::to_string(out_bytes));
    }
    m_dataOut->setLength(out_bytes);
    size_t n = 0;
    if (dabmod_b200_process(c->handle, it->second.data(), it->second.size(), m_dataOut->getData(), out_bytes, &n) !=
        DABMOD_B200_OK) {
        throw std::runtime_error(std::string("B200Blocks: ") + dabmod_b200_last_error());
    }
    c->frames.erase(c->frames.begin(), std::next(it));
    m_dataOut->setLength(n);
    return (int)m_dataOut->getLength();
}

meta_vec_t OutputMemory::process_metadata(const meta_vec_t& metadataIn)
{
    m_metadata = metadataIn;
    return {};
}

meta_vec_t OutputMemory::get_latest_metadata() { return m_metadata; }
