/* See B200EtiChain.h.  Host glue only: every coded bit and every sample is computed by the CUDA kernels behind
 * dabmod_b200_process_eti_batch(). */
#include "B200EtiChain.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "FicSource.h"
#include "FrameMultiplexer.h"
#include "SubchannelSource.h"
#include "dabmod_b200.h"

namespace {

constexpr size_t ETI_FRAME = 6144;
B200EtiChain* g_active = nullptr;

[[noreturn]] void fail(const char* what, const char* detail)
{
    throw std::runtime_error(std::string("B200EtiChain: ") + what + ": " + detail);
}

dabmod_b200_stream describe(size_t framesize, size_t out_bytes, size_t start_cu, const std::vector<PuncturingRule>& rules)
{
    dabmod_b200_stream s{};
    s.framesize = (uint32_t)framesize;
    s.out_bytes = (uint32_t)out_bytes;
    s.start_cu = (uint32_t)start_cu;
    if (rules.size() > sizeof(s.rules) / sizeof(s.rules[0])) fail("stream", "more than 8 puncturing rules");
    s.n_rules = (uint32_t)rules.size();
    for (size_t i = 0; i < rules.size(); i++) s.rules[i] = {(uint32_t)rules[i].length(), rules[i].pattern()};
    return s;
}

} // namespace

int B200SwapOutput::process(Buffer* dataIn)
{
    m_dataOut->swap(*dataIn);
    return (int)m_dataOut->getLength();
}

B200EtiChain::B200EtiChain(EtiSource& etiSource, mod_settings_t& settings, const std::string& format, int device,
                           bool fixedPoint, int batchTfs) :
    ModInput(),
    m_eti(etiSource),
    m_device(device),
    m_batch(batchTfs > 0 ? (size_t)batchTfs : 1),
    m_mode(settings.dabMode)
{
    /* DabModulator::setMode (src/DabModulator.cpp:84-126) */
    if (m_mode < 1 or m_mode > 4) throw std::runtime_error("DabModulator::setMode invalid mode size");
    m_trace = getenv("ODR_DABMOD_B200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    m_chain.reset(new B200OfdmChain(settings, format, device, fixedPoint, 0, (int)m_batch));
    /* process() returns the byte count as an int */
    const size_t out_tf = dabmod_b200_tf_out_bytes(m_chain->handle());
    if (m_batch * out_tf > 0x7fffffffu) {
        m_batch = 0x7fffffffu / out_tf;
        m_chain.reset(new B200OfdmChain(settings, format, device, fixedPoint, 0, (int)m_batch));
    }
    g_active = this;
    if (m_trace) {
        fprintf(stderr, "B200EtiChain: modulator handle for batches of %zu TFs created in %.3f s\n", m_batch,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
}

B200EtiChain::~B200EtiChain()
{
    if (g_active == this) g_active = nullptr;
    if (m_trace) {
        fprintf(stderr, "B200EtiChain: %zu frames in %.3f s since the first frame\n", m_n_frames,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - m_t_first).count());
        fprintf(stderr, "B200EtiChain: %zu frames collected in %.3f s, %zu batches in %.3f s on the GPU path "
                        "(%zu host ranges page-locked in %.3f s)\n",
                m_n_frames, m_t_collect, m_n_batches, m_t_gpu, m_pinned.size(), m_t_pin);
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (auto& p : m_pinned) dabmod_b200_host_unregister(p.first);
    dabmod_b200_coder_destroy(m_coder);
    m_chain.reset();
    if (m_trace) {
        fprintf(stderr, "B200EtiChain: released in %.3f s\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
}

bool B200EtiChain::flush()
{
    if (m_cif == 0 or m_collected < m_cif) return false;
    m_flush = true;
    return true;
}

bool B200EtiChain::flush_active()
{
    return g_active != nullptr and g_active->flush();
}

/* the construction of the per-stream blocks, FrameMultiplexer and BlockPartitioner of src/DabModulator.cpp:140-142,
 * 300-383, from the sources the EtiSource has built for this multiplex */
void B200EtiChain::build_coder()
{
    const unsigned mode = m_mode;       /* the configured mode, like DabModulator::process (src/DabModulator.cpp:133) */
    std::vector<dabmod_b200_stream> streams;
    auto fic = m_eti.getFic();
    streams.push_back(describe(fic->getFramesize(), mode == 3 ? 384 : 288, 0, fic->get_rules()));
    m_subs.clear();
    for (const auto& sub : m_eti.getSubchannels()) {
        streams.push_back(describe(sub->framesize(), sub->framesizeCu() * 8, sub->startAddress(), sub->get_rules()));
        m_subs.push_back({sub->framesize(), sub->startAddress(), sub->protection()});
    }
    m_cif = mode == 1 ? 4 : mode == 4 ? 2 : 1;
    if (dabmod_b200_coder_create(m_device, (int)mode, streams.data(), (int)streams.size(), (int)(m_batch * m_cif),
                                 &m_coder) != DABMOD_B200_OK) {
        fail("coder_create", dabmod_b200_coder_last_error());
    }
    m_offsets.resize(streams.size());
    for (size_t i = 0; i < streams.size(); i++) m_offsets[i] = dabmod_b200_coder_stream_offset(m_coder, (int)i);
    m_frames.assign(m_batch * m_cif * ETI_FRAME, 0);
}

bool B200EtiChain::same_multiplex() const
{
    const auto subs = m_eti.getSubchannels();
    if (subs.size() != m_subs.size()) return false;
    for (size_t i = 0; i < subs.size(); i++) {
        if (subs[i]->framesize() != m_subs[i].framesize or subs[i]->startAddress() != m_subs[i].start or
            subs[i]->protection() != m_subs[i].protection) {
            return false;
        }
    }
    return true;
}

/* Page-locks the storage the batch is about to land in.  B200SwapOutput makes the graph's two Buffers alternate,
 * each keeps its storage once it has the size of a batch (Buffer::setLength only reallocates to grow,
 * src/Buffer.cpp:128-147), so at most two ranges are registered, each once. */
void B200EtiChain::pin(Buffer* dataOut, size_t need)
{
    auto it = std::find_if(m_pinned.begin(), m_pinned.end(),
                           [&](const std::pair<void*, size_t>& p) { return p.first == dataOut->getData(); });
    if (it != m_pinned.end() and it->second >= need) {
        dataOut->setLength(need);
        return;
    }
    if (it != m_pinned.end()) {             /* registered, too small: release it before setLength frees it */
        dabmod_b200_host_unregister(it->first);
        m_pinned.erase(it);
    }
    dataOut->setLength(0);                  /* nothing to carry over into a new allocation */
    dataOut->setLength(need);
    /* a full batch only: the short batch of a flush does not justify a registration */
    if (need == m_batch * dabmod_b200_tf_out_bytes(m_chain->handle())) {
        if (dabmod_b200_host_register(dataOut->getData(), need) == DABMOD_B200_OK) {
            m_pinned.push_back({dataOut->getData(), need});
        }
        else if (m_trace) fprintf(stderr, "B200EtiChain: host_register: %s\n", dabmod_b200_last_error());
    }
}

int B200EtiChain::process(Buffer* dataOut)
{
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    if (m_n_frames == 0) m_t_first = t0;
    auto since = [](clk::time_point t) { return std::chrono::duration<double>(clk::now() - t).count(); };
    if (not m_flush) {
        if (not m_coder) build_coder();
        else if (not same_multiplex()) {
            /* what FrameMultiplexer::process reports (src/FrameMultiplexer.cpp:68-83) */
            throw FrameMultiplexerError("FrameMultiplexer detected subchannel size change from " +
                                        std::to_string(m_subs.size()) + " to " +
                                        std::to_string(m_eti.getSubchannels().size()));
        }
        /* the frame's payload at the coder's offsets (the coder reads nothing else of a frame) */
        uint8_t* frame = m_frames.data() + m_collected * ETI_FRAME;
        auto fic = m_eti.getFic();
        fic->process(&m_tmp);
        memcpy(frame + m_offsets[0], m_tmp.getData(), m_tmp.getLength());
        for (const auto& md : fic->process_metadata({})) m_meta.push_back(md);
        size_t i = 1;
        for (const auto& sub : m_eti.getSubchannels()) {
            sub->process(&m_tmp);
            memcpy(frame + m_offsets[i++], m_tmp.getData(), m_tmp.getLength());
        }
        m_collected++;
        m_n_frames++;
        m_t_collect += since(t0);
    }

    const bool full = m_collected == m_batch * m_cif;
    if (not full and not m_flush) {
        dataOut->setLength(0);
        return 0;
    }
    m_flush = false;
    const size_t n_tf = m_collected / m_cif, n_frames = n_tf * m_cif;
    if (n_tf == 0) {
        dataOut->setLength(0);
        return 0;
    }
    const size_t need = n_tf * dabmod_b200_tf_out_bytes(m_chain->handle());
    const auto t1 = clk::now();
    pin(dataOut, need);
    m_t_pin += since(t1);
    const auto t2 = clk::now();
    size_t nb = 0;
    if (dabmod_b200_process_eti_batch(m_chain->handle(), m_coder, m_frames.data(), n_frames, dataOut->getData(),
                                      dataOut->getLength(), &nb) != DABMOD_B200_OK) {
        fail("process_eti_batch", dabmod_b200_coder_last_error());
    }
    m_t_gpu += since(t2);
    m_n_batches++;
    /* frames of an incomplete TF (only after a flush) stay for the next batch */
    const size_t rest = m_collected - n_frames;
    if (rest) memmove(m_frames.data(), m_frames.data() + n_frames * ETI_FRAME, rest * ETI_FRAME);
    m_collected = rest;
    m_meta_out = std::move(m_meta);
    m_meta.clear();
    dataOut->setLength(nb);
    return (int)nb;
}

meta_vec_t B200EtiChain::process_metadata(const meta_vec_t&)
{
    meta_vec_t r = std::move(m_meta_out);
    m_meta_out.clear();
    return r;
}
