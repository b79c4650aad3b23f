/* See B200EtiChain.h.  Host glue only: every coded bit and every sample is computed by the CUDA kernels behind
 * dabmod_b200_process_eti_batch(). */
#include "B200EtiChain.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
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
    m_async = getenv("ODR_DABMOD_B200_SYNC") == nullptr;
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
    if (m_job.valid()) {
        try { m_job.get(); } catch (...) {}
    }
    if (m_trace) {
        fprintf(stderr, "B200EtiChain: %zu frames in %.3f s since the first frame\n", m_n_frames,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - m_t_first).count());
        fprintf(stderr, "B200EtiChain: %zu frames collected in %.3f s, %zu batches in %.3f s on the GPU path, of "
                        "which the caller waited %.3f s (%zu host ranges page-locked in %.3f s)\n",
                m_n_frames, m_t_collect, m_n_batches, m_t_gpu, m_t_wait, m_pinned.size(), m_t_pin);
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
    if (not m_job.valid() and (m_cif == 0 or m_collected < m_cif)) return false;
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
    m_collecting.assign(m_batch * m_cif * ETI_FRAME, 0);
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

/* Sizes and page-locks the Buffer a batch is about to land in.  The storages circulate: a finished batch is swapped
 * into the graph's edge Buffer, B200SwapOutput swaps that with DabModulator's output Buffer, and what comes back is
 * the storage of an older batch -- each keeps its size (Buffer::setLength only reallocates to grow,
 * src/Buffer.cpp:128-147), so at most four ranges are registered, each once. */
void B200EtiChain::pin(Buffer& buf, size_t need)
{
    auto it = std::find_if(m_pinned.begin(), m_pinned.end(),
                           [&](const std::pair<void*, size_t>& p) { return p.first == buf.getData(); });
    if (it != m_pinned.end() and it->second >= need) {
        buf.setLength(need);
        return;
    }
    if (it != m_pinned.end()) {             /* registered, too small: release it before setLength frees it */
        dabmod_b200_host_unregister(it->first);
        m_pinned.erase(it);
    }
    buf.setLength(0);                       /* nothing to carry over into a new allocation */
    buf.setLength(need);
    /* a full batch only: the short batch of a flush does not justify a registration */
    if (need == m_batch * dabmod_b200_tf_out_bytes(m_chain->handle())) {
        if (dabmod_b200_host_register(buf.getData(), need) == DABMOD_B200_OK) {
            m_pinned.push_back({buf.getData(), need});
        }
        else if (m_trace) fprintf(stderr, "B200EtiChain: host_register: %s\n", dabmod_b200_last_error());
    }
}

/* Hands the collected whole transmission frames to the GPU: the call runs on a worker thread while the caller
 * reads and parses the frames of the next batch (the N-TF form of PipelinedModCodec, src/ModPlugin.cpp:90-115). */
void B200EtiChain::launch()
{
    using clk = std::chrono::steady_clock;
    const size_t n_tf = m_collected / m_cif, n_frames = n_tf * m_cif;
    if (n_tf == 0) return;
    const int slot = m_next_slot;
    m_next_slot ^= 1;
    const auto t1 = clk::now();
    pin(m_out[slot], n_tf * dabmod_b200_tf_out_bytes(m_chain->handle()));
    m_t_pin += std::chrono::duration<double>(clk::now() - t1).count();
    m_frames[slot].swap(m_collecting);                  /* the batch's frames; m_collecting takes the free array */
    if (m_collecting.size() != m_frames[slot].size()) m_collecting.assign(m_frames[slot].size(), 0);
    /* frames of an incomplete TF (only after a flush) open the next batch */
    const size_t rest = m_collected - n_frames;
    if (rest) memcpy(m_collecting.data(), m_frames[slot].data() + n_frames * ETI_FRAME, rest * ETI_FRAME);
    m_collected = rest;
    m_job_meta = std::move(m_meta);
    m_meta.clear();
    m_job_slot = slot;
    const uint8_t* frames = m_frames[slot].data();
    Buffer* out = &m_out[slot];
    m_job = std::async(std::launch::async, [this, frames, n_frames, out]() -> double {
        const auto t2 = clk::now();
        size_t nb = 0;
        if (dabmod_b200_process_eti_batch(m_chain->handle(), m_coder, frames, n_frames, out->getData(),
                                          out->getLength(), &nb) != DABMOD_B200_OK) {
            fail("process_eti_batch", dabmod_b200_coder_last_error());
        }
        out->setLength(nb);
        return std::chrono::duration<double>(clk::now() - t2).count();
    });
}

/* the finished batch (waits for it), swapped into dataOut; 0 when none is in flight */
int B200EtiChain::collect(Buffer* dataOut)
{
    if (not m_job.valid()) {
        dataOut->setLength(0);
        return 0;
    }
    const auto t0 = std::chrono::steady_clock::now();
    m_t_gpu += m_job.get();                             /* rethrows what the worker threw */
    m_t_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    m_n_batches++;
    dataOut->swap(m_out[m_job_slot]);
    m_meta_out = std::move(m_job_meta);
    m_job_meta.clear();
    return (int)dataOut->getLength();
}

int B200EtiChain::process(Buffer* dataOut)
{
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    if (m_n_frames == 0) m_t_first = t0;
    if (m_flush) {
        /* end of the stream: first the batch in flight, then (next call) whatever whole TFs are left */
        m_flush = false;
        if (m_job.valid()) return collect(dataOut);
        launch();
        return collect(dataOut);
    }
    if (not m_coder) build_coder();
    else if (m_reconfigure or not same_multiplex()) {
        /* A changed multiplex ends this graph: FrameMultiplexerError, with what FrameMultiplexer::process reports
         * (src/FrameMultiplexer.cpp:68-83), makes run_modulator build a new modulator (src/DabMod.cpp:744-749).
         * The frames collected so far are still good: they leave first (one or two calls, whose own frames are
         * dropped -- the new modulator waits for the next frame with FP 0 anyway, src/DabMod.cpp:691-700). */
        if (not m_reconfigure) {
            m_reconfigure = true;
            m_reconfigure_what = "FrameMultiplexer detected subchannel size change from " + std::to_string(m_subs.size()) +
                                 " to " + std::to_string(m_eti.getSubchannels().size());
        }
        if (m_job.valid()) return collect(dataOut);
        if (m_collected >= m_cif) {
            launch();
            return collect(dataOut);
        }
        throw FrameMultiplexerError(m_reconfigure_what);
    }
    /* the frame's payload at the coder's offsets (the coder reads nothing else of a frame) */
    uint8_t* frame = m_collecting.data() + m_collected * ETI_FRAME;
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
    m_t_collect += std::chrono::duration<double>(clk::now() - t0).count();

    if (m_collected < m_batch * m_cif) {
        dataOut->setLength(0);
        return 0;
    }
    /* a batch is complete: take the previous one back (it has had a whole batch of host time), start this one */
    const int n = collect(dataOut);
    launch();
    if (not m_async) return collect(dataOut);           /* ODR_DABMOD_B200_SYNC: no batch in flight between calls */
    return n;
}

meta_vec_t B200EtiChain::process_metadata(const meta_vec_t&)
{
    meta_vec_t r = std::move(m_meta_out);
    m_meta_out.clear();
    return r;
}
