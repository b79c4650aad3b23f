/* B200EtiChain -- the whole of DabModulator's graph on the GPU, inside the reference program.
 *
 * A ModInput (reference src/ModPlugin.h:70-77) that stands where DabModulator::process builds and runs
 *   FicSource / SubchannelSource -> PrbsGenerator -> ConvEncoder -> PuncturingEncoder [-> TimeInterleaver]
 *   -> FrameMultiplexer -> BlockPartitioner -> (the OFDM chain of B200OfdmChain.h) -> OutputMemory
 * (reference src/DabModulator.cpp:131-417).  It reads the frame the EtiSource has just parsed
 * (EtiReader::loadEtiData or the EDI reader; src/EtiReader.cpp:93-284) from the sources themselves -- the FIC and
 * subchannel bytes -- takes the puncturing rules from them (get_rules(), every protection profile including UEP),
 * collects `batchTfs` transmission frames, and runs them through dabmod_b200_process_eti_batch as one batch: coder
 * and modulator chained on the device, I/Q straight into the output Buffer.
 *
 * One call per ETI frame, like DabModulator::process.  While a batch fills, process() returns 0 with an empty buffer
 * (the flowgraph iteration ends there, src/Flowgraph.cpp:331-336, as it does on the three of four TM I frames where
 * BlockPartitioner has no block yet).  The call that completes batch k starts it on a worker thread and returns
 * batch k - 1 (batchTfs transmission frames), so the host reads and parses the next frames while the GPU works --
 * PipelinedModCodec's one-call delay (src/ModPlugin.cpp:90-115) with a batch as the unit; ODR_DABMOD_B200_SYNC in
 * the environment returns batch k itself.
 * The metadata of every frame of the batch (FicSource::process_metadata, collected like
 * BlockPartitioner::process_metadata does for one TF, src/BlockPartitioner.cpp:126-140) leaves with it.
 * flush() makes the next call emit what is pending (end of a file): the batch in flight, then, on a second
 * flush + call, the whole transmission frames collected so far.
 *
 * A changed multiplex (number, size, position or protection of the subchannels) first drains the transmission frames
 * collected so far, then throws FrameMultiplexerError with the reference's message (src/FrameMultiplexer.cpp:68-83):
 * run_modulator restarts the modulator on it (src/DabMod.cpp:744-749).
 *
 * B200SwapOutput is the OutputMemory of this graph: same role and metadata handling
 * (src/OutputMemory.cpp:62-97), but it exchanges the two buffers' storage instead of copying it -- a batch is
 * hundreds of megabytes.
 */
#pragma once

#include <chrono>
#include <future>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "B200OfdmChain.h"
#include "Buffer.h"
#include "EtiReader.h"
#include "ModPlugin.h"
#include "OutputMemory.h"

struct dabmod_b200_coder;

/* "modulator.fft_engine = b200_eti | b200_eti_fixed": channel coding + OFDM chain on the GPU */
constexpr FFTEngine FFTENGINE_B200_ETI = static_cast<FFTEngine>(5);
constexpr FFTEngine FFTENGINE_B200_ETI_FIXED = static_cast<FFTEngine>(6);

inline bool b200_engine_is_eti(FFTEngine e) { return e == FFTENGINE_B200_ETI or e == FFTENGINE_B200_ETI_FIXED; }
inline bool b200_engine_is_fixed(FFTEngine e) { return e == FFTENGINE_B200_FIXED or e == FFTENGINE_B200_ETI_FIXED; }
inline bool b200_engine(FFTEngine e) { return static_cast<int>(e) >= 3 and static_cast<int>(e) <= 6; }

class B200SwapOutput : public OutputMemory
{
public:
    explicit B200SwapOutput(Buffer* dataOut) : OutputMemory(dataOut) {}
    int process(Buffer* dataIn) override;
    const char* name() override { return "B200SwapOutput"; }
};

class B200EtiChain : public ModInput, public ModMetadata
{
public:
    B200EtiChain(EtiSource& etiSource, mod_settings_t& settings, const std::string& format, int device = 0,
                 bool fixedPoint = false, int batchTfs = 64);
    virtual ~B200EtiChain();
    B200EtiChain(const B200EtiChain&) = delete;
    B200EtiChain& operator=(const B200EtiChain&) = delete;

    int process(Buffer* dataOut) override;
    const char* name() override { return "B200EtiChain"; }
    meta_vec_t process_metadata(const meta_vec_t& metadataIn) override;

    /* the OFDM chain's controllables ("b200chain" and "tii"): enrol them like B200OfdmChain's */
    B200OfdmChain& chain() { return *m_chain; }

    /* The next process() call takes no new frame and emits the batch in flight or, when there is none, the whole
     * TFs collected so far.  Returns false when there is nothing left to emit. */
    bool flush();
    /* flush() on the live instance, for the end-of-file branch of run_modulator (src/DabMod.cpp:612-617) */
    static bool flush_active();

private:
    void build_coder();
    bool same_multiplex() const;
    void pin(Buffer& buf, size_t need);
    void launch();
    int collect(Buffer* dataOut);

    EtiSource& m_eti;
    std::unique_ptr<B200OfdmChain> m_chain;
    dabmod_b200_coder* m_coder = nullptr;
    int m_device;
    size_t m_batch;                 /* TFs per batch */
    unsigned m_mode;
    /* ODR_DABMOD_B200_TRACE: where the time went, printed at destruction */
    bool m_trace = false;
    double m_t_collect = 0, m_t_gpu = 0, m_t_pin = 0;
    size_t m_n_frames = 0, m_n_batches = 0;
    std::chrono::steady_clock::time_point m_t_first;
    size_t m_cif = 0;               /* ETI frames per TF */
    size_t m_collected = 0;         /* frames in m_frames */
    bool m_flush = false;
    bool m_reconfigure = false;     /* the multiplex has changed: drain, then throw */
    std::string m_reconfigure_what;
    /* m_batch * m_cif frames of 6144 bytes, payload at the coder's offsets: the one being filled, and the two
     * that batches in flight read from */
    std::vector<uint8_t> m_collecting, m_frames[2];
    Buffer m_out[2];                /* where the batches in flight land */
    int m_next_slot = 0, m_job_slot = 0;
    std::future<double> m_job;      /* the batch on the GPU (seconds it took) */
    meta_vec_t m_job_meta;
    bool m_async = true;
    double m_t_wait = 0;
    std::vector<int> m_offsets;
    /* what the coder was built for: framesize / startAddress / protection per subchannel */
    struct Sub { size_t framesize, start, protection; };
    std::vector<Sub> m_subs;
    meta_vec_t m_meta, m_meta_out;
    std::vector<std::pair<void*, size_t>> m_pinned;
    Buffer m_tmp;
};
