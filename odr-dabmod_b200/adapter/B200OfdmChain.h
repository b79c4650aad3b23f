/* B200OfdmChain -- the reference-facing side of libdabmod_b200.
 *
 * A ModCodec (reference src/ModPlugin.h:80-87) that takes the buffer
 * BlockPartitioner produces and emits the buffer OutputMemory receives, i.e.
 * it stands where DabModulator wires
 *   cifMap -> cifFreq -> cifDiff(+cifRef) -> cifSig(+cifNull/tii) -> [cifCicEq]
 *   -> cifOfdm -> cifGain -> cifGuard -> [cifFilter] -> [cifRes] -> [cifPoly]
 *   -> [m_formatConverter]
 * (reference src/DabModulator.cpp:386-417), and runs that chain on a B200
 * through the C ABI of include/dabmod_b200.h.
 *
 * Compiled against the UNMODIFIED reference headers (Buffer.h, ModPlugin.h,
 * ConfigParser.h, RemoteControl.h); see INTEGRATION.md for the DabModulator
 * hunk that selects it.  Same conventions as the blocks it replaces:
 *   - process() returns the output byte count (0 only while a pipeline primes, see below), size mismatches
 *     throw std::runtime_error;
 *   - parameters are remote-controllable under the names of the replaced
 *     blocks ("digital", "mode", "var", "tii.enable", ...);
 *   - pipelineDepth = 0: the TF comes back in the same call and metadata passes through unchanged.
 *   - pipelineDepth = D > 0: the N-TF generalisation of PipelinedModCodec (src/ModPlugin.cpp:90-128, which has
 *     D = 1): call i hands in TF i and receives TF i - D; every D-th call runs the D collected TFs through
 *     dabmod_b200_process_batch as ONE batch, so the GPU sees D frames at a time instead of one.  The first D
 *     calls return 0 with an empty buffer (the flowgraph iteration ends there, src/Flowgraph.cpp:331-336), which
 *     is what a PipelinedModCodec's first call amounts to, and the metadata (frame timestamps,
 *     src/ModPlugin.h:50-55) is delayed by the same D calls.
 */
#pragma once

#include <deque>
#include <string>
#include <vector>

#include "Buffer.h"
#include "ConfigParser.h"
#include "ModPlugin.h"
#include "RemoteControl.h"

struct dabmod_b200;
class B200OfdmChain;

/* The engines this adapter adds to `modulator.fft_engine` (src/ConfigParser.cpp:66-85): "b200" = the float chain,
 * "b200_fixed" = the KISS fixed-point chain, both on the GPU.  They are further values of the reference's
 * `enum class FFTEngine` (src/ConfigParser.h:39-43: FFTW, KISS, DEXTER = 0, 1, 2); named here so that the
 * reference header stays untouched. */
constexpr FFTEngine FFTENGINE_B200 = static_cast<FFTEngine>(3);
constexpr FFTEngine FFTENGINE_B200_FIXED = static_cast<FFTEngine>(4);

/* The reference enrols TII as its own controllable "tii" with the parameters enable / comb / pattern /
 * old_variant (src/TII.cpp:106-127, 339-376): same name and parameters here, forwarded to the chain's handle. */
class B200TiiControl : public RemoteControllable
{
public:
    explicit B200TiiControl(B200OfdmChain& chain);
    void set_parameter(const std::string& parameter, const std::string& value) override;
    const std::string get_parameter(const std::string& parameter) const override;
    const json::map_t get_all_values() const override;

private:
    B200OfdmChain& m_chain;
};

class B200OfdmChain : public ModCodec, public ModMetadata, public RemoteControllable
{
public:
    /* `format`: "" / "complexf" / "s16" / "u8" / "s8" (what DabModulator hands to
     * FormatConverter); device: CUDA ordinal. */
    /* fixedPoint (or settings.fftEngine == KISS) selects the fixed-point engine: the chain DabModulator builds
     * for FFTEngine::KISS (DabModulator.cpp:144-224), int16 I/Q out, bit-exact with it */
    /* `settings` is DabModulator's own mod_settings_t (src/DabModulator.h:70 holds it by reference too): remote
     * control changes are written back into it, so they survive a modulator restart like those of the blocks
     * that hold references into it (src/GainControl.h:72-77, src/OfdmGenerator.h:50-56, src/TII.h:82,
     * src/GuardIntervalInserter.h:48-54). */
    /* maxBatch: the largest batch the handle must take (dabmod_b200_config.max_batch), for callers that drive the
     * handle's batch entry points themselves (B200EtiChain); process() needs max(1, pipelineDepth). */
    B200OfdmChain(mod_settings_t& settings, const std::string& format, int device = 0,
                  bool fixedPoint = false, int pipelineDepth = 0, int maxBatch = 0);
    virtual ~B200OfdmChain();
    B200OfdmChain(const B200OfdmChain&) = delete;
    B200OfdmChain& operator=(const B200OfdmChain&) = delete;

    int process(Buffer* const dataIn, Buffer* dataOut) override;
    const char* name() override { return "B200OfdmChain"; }
    meta_vec_t process_metadata(const meta_vec_t& metadataIn) override;

    /* FormatConverter::get_num_clipped_samples (src/FormatConverter.cpp:186-189) */
    size_t get_num_clipped_samples() const;

    /* RemoteControllable */
    void set_parameter(const std::string& parameter, const std::string& value) override;
    const std::string get_parameter(const std::string& parameter) const override;
    const json::map_t get_all_values() const override;

    /* the "tii" controllable (enrol it next to the chain: rcs.enrol(chain->tii_control())) */
    RemoteControllable* tii_control() { return &m_tii; }

    /* the C-ABI handle behind the chain (include/dabmod_b200.h) */
    dabmod_b200* handle() const { return m_handle; }

private:
    dabmod_b200* m_handle = nullptr;
    mod_settings_t& m_settings;
    B200TiiControl m_tii;
    void write_back(const std::string& parameter);
    size_t m_depth = 0;             /* TFs per batch, 0 = no pipeline */
    size_t m_calls = 0;
    std::vector<uint8_t> m_in;      /* the TFs collected for the next batch */
    std::vector<uint8_t> m_ready;   /* the outputs of the last batch */
    std::deque<meta_vec_t> m_metadata_fifo;
};
