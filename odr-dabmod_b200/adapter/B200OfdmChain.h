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
 *   - process() returns the output byte count, 0 never (there is no priming
 *     latency: unlike the PipelinedModCodec stages the TF comes back in the
 *     same call), size mismatches throw std::runtime_error;
 *   - parameters are remote-controllable under the names of the replaced
 *     blocks ("digital", "mode", "var", "tii.enable", ...);
 *   - metadata (frame timestamps) passes through unchanged, so no ModMetadata.
 */
#pragma once

#include <string>

#include "Buffer.h"
#include "ConfigParser.h"
#include "ModPlugin.h"
#include "RemoteControl.h"

struct dabmod_b200;

class B200OfdmChain : public ModCodec, public RemoteControllable
{
public:
    /* `format`: "" / "complexf" / "s16" / "u8" / "s8" (what DabModulator hands to
     * FormatConverter); device: CUDA ordinal. */
    /* fixedPoint (or settings.fftEngine == KISS) selects the fixed-point engine: the chain DabModulator builds
     * for FFTEngine::KISS (DabModulator.cpp:144-224), int16 I/Q out, bit-exact with it */
    B200OfdmChain(const mod_settings_t& settings, const std::string& format, int device = 0,
                  bool fixedPoint = false);
    virtual ~B200OfdmChain();
    B200OfdmChain(const B200OfdmChain&) = delete;
    B200OfdmChain& operator=(const B200OfdmChain&) = delete;

    int process(Buffer* const dataIn, Buffer* dataOut) override;
    const char* name() override { return "B200OfdmChain"; }

    /* FormatConverter::get_num_clipped_samples (src/FormatConverter.cpp:186-189) */
    size_t get_num_clipped_samples() const;

    /* RemoteControllable */
    void set_parameter(const std::string& parameter, const std::string& value) override;
    const std::string get_parameter(const std::string& parameter) const override;
    const json::map_t get_all_values() const override;

private:
    dabmod_b200* m_handle = nullptr;
};
