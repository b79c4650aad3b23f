/* dabmod_b200.h -- C ABI of the B200-native COFDM I/Q generation path.
 *
 * One handle = one modulator stream = the hot chain that ODR-DabMod's
 * DabModulator wires between BlockPartitioner and OutputMemory
 * (reference src/DabModulator.cpp:386-417):
 *
 *   QpskSymbolMapper -> FrequencyInterleaver -> DifferentialModulator(+PhaseReference)
 *   -> SignalMultiplexer(+NullSymbol/TII) -> [CicEqualizer] -> OfdmGenerator(+CFR)
 *   -> GainControl -> GuardIntervalInserter -> [FIRFilter] -> [Resampler]
 *   -> [MemlessPoly] -> [FormatConverter]
 *
 * Input of every process call: the byte block BlockPartitioner emits per
 * transmission frame (TF) (reference src/BlockPartitioner.cpp:78-124), i.e.
 * what QpskSymbolMapper::process receives (src/QpskSymbolMapper.cpp:39).
 * Output: the bytes OutputMemory::process copies to the DabModulator caller
 * (src/OutputMemory.cpp:62-83), same layout, same length.
 *
 * Plain C, no exceptions cross this boundary: every call returns 0 on success
 * or a negative DABMOD_B200_E* code, and dabmod_b200_last_error() describes
 * the last failure of the calling thread.  The caller owns all host memory;
 * the library owns device memory, streams and events.  The library needs a
 * CUDA device of compute capability 10.x; there is no CPU fallback.
 */
#ifndef DABMOD_B200_H
#define DABMOD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DABMOD_B200_ABI_VERSION 2

enum {
    DABMOD_B200_OK = 0,
    DABMOD_B200_EINVAL = -1,    /* bad argument / size mismatch (the reference throws std::runtime_error) */
    DABMOD_B200_ECUDA = -2,     /* CUDA runtime error, no usable device */
    DABMOD_B200_ENOMEM = -3,
    DABMOD_B200_EUNSUPPORTED = -4, /* valid for the reference, not implemented here */
    DABMOD_B200_ESTATE = -5,
    DABMOD_B200_EIO = -6           /* the output sink failed (write(2) error) */
};

/* GainMode, reference src/GainControl.h:45 */
enum { DABMOD_B200_GAIN_FIX = 0, DABMOD_B200_GAIN_MAX = 1, DABMOD_B200_GAIN_VAR = 2 };

/* Output sample format, reference src/FormatConverter.cpp:112-165 and
 * DabMod.cpp:250-363 ("complexf" = no FormatConverter node). */
enum { DABMOD_B200_FMT_COMPLEXF = 0, DABMOD_B200_FMT_S16 = 1, DABMOD_B200_FMT_U8 = 2, DABMOD_B200_FMT_S8 = 3 };

/* FFTEngine, reference src/ConfigParser.h:39-43.  FFTW = the float32 chain; KISS = the fixed-point chain of
 * DabModulator.cpp:144-224: complexfix carriers, OfdmGeneratorFixed (vendored KISS FFT, FIXED_POINT=16), no
 * GainControl, GuardIntervalInserter<complexfix>; the output is int16 I/Q ("KISS is already in s16",
 * DabModulator.cpp:279) and bit-exact with the reference.  Like the reference it rejects FIR, resampler and
 * predistortion ("fixed point doesn't support ...", DabModulator.cpp:249,257,265); CFR does not exist there. */
enum { DABMOD_B200_FFT_FLOAT = 0, DABMOD_B200_FFT_KISS_FIXED = 1 };

/* MemlessPoly, reference src/MemlessPoly.cpp:109-110,145-229 */
enum { DABMOD_B200_DPD_NONE = 0, DABMOD_B200_DPD_ODD_POLY = 1, DABMOD_B200_DPD_LUT = 2 };

/* Mirrors the fields of mod_settings_t that parametrise the hot path
 * (reference src/ConfigParser.h:45-96) with the same defaults when zeroed
 * through dabmod_b200_config_init(). */
typedef struct dabmod_b200_config {
    uint32_t abi_version;     /* DABMOD_B200_ABI_VERSION */
    int32_t  device;          /* CUDA device ordinal */
    int32_t  mode;            /* dabMode 1..4 (0 is treated as 1, DabModulator.cpp:75-77) */
    int32_t  gain_mode;       /* gainMode, default VAR */
    uint64_t output_rate;     /* outputRate in Hz, default 2048000 (no Resampler) */
    uint64_t clock_rate;      /* clockRate, 0 = no CicEqualizer (DabModulator.cpp:154-176) */
    float    digital_gain;    /* digitalgain, default 1 */
    float    normalise;       /* normalise, default 1 */
    float    gain_variance;   /* gainmodeVariance, default 4 */
    int32_t  window_overlap;  /* ofdmWindowOverlap, default 0 */
    int32_t  cfr_enable;      /* enableCfr */
    float    cfr_clip;        /* cfrClip */
    float    cfr_errclip;     /* cfrErrorClip */
    int32_t  tii_enable;      /* tiiConfig.enable */
    int32_t  tii_comb;        /* tiiConfig.comb 0..23 */
    int32_t  tii_pattern;     /* tiiConfig.pattern 0..69 */
    int32_t  tii_old_variant; /* tiiConfig.old_variant */
    int32_t  fir_ntaps;       /* 0 = no FIRFilter; taps as FIRFilter::load_filter_taps leaves them */
    const float *fir_taps;
    int32_t  dpd_mode;        /* DABMOD_B200_DPD_* */
    const float *dpd_coefs;   /* ODD_POLY: am0..4, pm0..4 (10 floats); LUT: scalefactor + 32 entries (33 floats) */
    int32_t  format;          /* DABMOD_B200_FMT_* */
    int32_t  max_batch;       /* largest n_tf a *_batch call will pass (sizes device buffers), default 1 */
    int32_t  fft_engine;      /* DABMOD_B200_FFT_*, default FLOAT (fftEngine = FFTW) */
} dabmod_b200_config;

typedef struct dabmod_b200 dabmod_b200;

/* Fills *cfg with the reference defaults (ConfigParser.h:45-96). */
void dabmod_b200_config_init(dabmod_b200_config *cfg);

/* The 45 built-in taps the reference uses for `filtertapsfile=default`
 * (src/FIRFilter.cpp:50-71); returns the count. `taps` may be NULL. */
int dabmod_b200_default_fir_taps(float *taps, int cap);

/* Builds the kernel tables and allocates device buffers for max_batch TFs.
 * Replaces the construction of the blocks in DabModulator.cpp:131-279. */
int dabmod_b200_create(const dabmod_b200_config *cfg, dabmod_b200 **out);
void dabmod_b200_destroy(dabmod_b200 *h);

/* Bytes per TF on either side of the path. in: (L-1)*K/4 (BlockPartitioner);
 * out: samples * {8,4,2,2}; the fixed-point engine always writes int16 pairs (4 bytes per sample). */
size_t dabmod_b200_tf_in_bytes(const dabmod_b200 *h);
size_t dabmod_b200_tf_out_bytes(const dabmod_b200 *h);
size_t dabmod_b200_tf_out_samples(const dabmod_b200 *h);

/* One TF, host buffers: what Node::process -> ModCodec::process(Buffer*, Buffer*)
 * does for the whole chain (src/Flowgraph.cpp:126-145).  `nbytes` must equal
 * dabmod_b200_tf_in_bytes (the reference throws "input size not valid").
 * Unlike the reference's PipelinedModCodec stages there is no 1-call latency:
 * the result of this TF is returned by this call. */
int dabmod_b200_process(dabmod_b200 *h, const uint8_t *bits, size_t nbytes,
                        void *iq_out, size_t cap, size_t *out_bytes);

/* n_tf consecutive TFs of the stream, host buffers (pinned or pageable).
 * Copies are chunked and overlapped with the kernels on internal streams. */
int dabmod_b200_process_batch(dabmod_b200 *h, const uint8_t *bits, size_t n_tf,
                              void *iq_out, size_t cap, size_t *out_bytes);

/* Same, delivered to a file descriptor instead of a buffer: what the reference's OutputFile node does with the
 * chain's output (src/OutputFile.cpp:56-67, fwrite of every Buffer; "fileoutput" in the configuration).  The I/Q
 * goes through a ring of pinned host buffers owned by the handle: slice i is written to `fd` while slice i+1 is
 * on the PCIe link and slice i+2 in the kernels.  Short writes are completed, EINTR retried;
 * any other write error returns DABMOD_B200_EIO (the reference throws).  *out_bytes = bytes written. */
int dabmod_b200_process_batch_to_fd(dabmod_b200 *h, const uint8_t *bits, size_t n_tf, int fd, size_t *out_bytes);

/* Same, device-resident buffers on the handle's device; enqueued on `stream`
 * (a cudaStream_t, NULL = the handle's own stream) and NOT synchronised.
 * Successive calls on one handle are ordered by the library whatever their streams (an event recorded after
 * each enqueue is waited for by the next): the work buffers, the resampler history and the tables belong to
 * the handle.  With CFR enabled the call first waits on the host for the previous one (its per-symbol records
 * are read back before they are overwritten).  The FormatConverter clip count of such a call is read back by
 * dabmod_b200_num_clipped_samples / dabmod_b200_synchronize, which wait for it. */
int dabmod_b200_process_batch_device(dabmod_b200 *h, const uint8_t *d_bits, size_t n_tf,
                                     void *d_iq_out, void *stream);

/* Blocks until everything enqueued for the handle (its own streams and the stream of the last
 * dabmod_b200_process_batch_device call) is done. */
int dabmod_b200_synchronize(dabmod_b200 *h);

/* Forget the stream history: Resampler overlap buffers (src/Resampler.cpp:110-111)
 * and the TII every-second-frame toggle (src/TII.cpp:225-242). */
int dabmod_b200_reset(dabmod_b200 *h);

/* For sharding one stream across GPUs by TF ranges: tells the handle that its
 * next TF is number `tf_index` of the stream (TII parity), and optionally
 * primes the resampler history from the FIR-stage output of the previous TF by
 * re-running that TF (`prev_bits` = its input block, NULL = stream start). */
int dabmod_b200_seek(dabmod_b200 *h, uint64_t tf_index, const uint8_t *prev_bits, size_t nbytes);

/* Remote-control parameters with the reference's names and value syntax:
 *   "digital", "mode" (fix|max|var), "var"      GainControl.cpp:520-603
 *   "windowlen"                                 GuardIntervalInserter.cpp:338-379
 *   "cfr", "clip", "errorclip"                  OfdmGenerator.cpp:376-404
 *   "clip_stats", "papr" (read-only)            OfdmGenerator.cpp:419-453: the same strings the reference
 *                                               prints, from per-symbol records written by the symbol kernel
 *                                               (clip / error-clip ratios and MER over the last 10 frames,
 *                                               PAPR before / after CFR over the last 50 frames' symbols)
 *   "enable", "comb", "pattern", "old_variant"  TII.cpp:339-376 (prefixed "tii." here)
 * plus "taps" (count, then taps, whitespace separated: FIRFilter tapsfile
 * content) and "coefs" (MemlessPoly coefficient-file content).
 * Not in the reference: "profile" (0|1, see dabmod_b200_kernel_time), the kernel selection knobs
 * "sym_kernel" / "fir_kernel" / "res_kernel" (0 = always the general kernel; they never change results) and
 * "sym_chunks" (CTAs per TF of the symbol kernel, 0 = automatic; a tuning knob
 * that never changes results).
 * Takes effect at the next process call. */
int dabmod_b200_set_param(dabmod_b200 *h, const char *name, const char *value);
int dabmod_b200_get_param(dabmod_b200 *h, const char *name, char *buf, size_t cap);

/* FormatConverter::get_num_clipped_samples of the last process call
 * (src/FormatConverter.cpp:176,186-189). */
uint64_t dabmod_b200_num_clipped_samples(dabmod_b200 *h);

/* How many kernels the last process call launched (bench.py reports it). */
uint32_t dabmod_b200_last_launch_count(const dabmod_b200 *h);

/* With set_param("profile", "1") every kernel launch is bracketed by CUDA events
 * on its stream; this returns name and duration of launch `idx` of the last
 * process call (synchronises on its end event).  Profiling inserts event records
 * between kernels and is off by default. */
int dabmod_b200_kernel_time(dabmod_b200 *h, int idx, char *name, size_t cap, float *ms);

/* ---- host-only table introspection (no GPU needed; used by the CPU tests) ----
 * The constant tables the kernels consume, in the reference's own indexing so
 * they can be compared with the reference blocks that build them. */

/* FrequencyInterleaver::m_indices (src/FrequencyInterleaver.cpp:73-92): K entries. */
int dabmod_b200_table_interleaver(int mode, int32_t *idx, int cap);
/* PhaseReference values as quarter turns 0..3 (src/PhaseReference.cpp:152-171): K entries. */
int dabmod_b200_table_phase_ref(int mode, uint8_t *q, int cap);
/* TII::m_Acp (src/TII.cpp:247-337): K flags; returns DABMOD_B200_EUNSUPPORTED for
 * modes where the reference's TII constructor throws (TM III/IV). */
int dabmod_b200_table_tii(int mode, int comb, int pattern, uint8_t *acp, int cap);
/* CicEqualizer::myFilter (src/CicEqualizer.cpp:29-57): K floats. */
int dabmod_b200_table_cic(int n_carriers, float spacing, int ratio, float *filter);
/* Resampler FFT sizes (src/Resampler.cpp:65-76). */
int dabmod_b200_resampler_sizes(uint64_t in_rate, uint64_t out_rate, int resolution,
                                int *fft_in, int *fft_out);

/* Thread-local description of the last error. Never NULL. */
const char *dabmod_b200_last_error(void);

/* The handle's device-resident output buffer (max_batch TFs in the output format); what
 * dabmod_b200_process_batch copies to the host.  For callers that chain device work. */
void *dabmod_b200_device_out(dabmod_b200 *h);

/* Page-locks / releases a host range the caller owns (e.g. the memory of a reference `Buffer`, src/Buffer.cpp:
 * 128-147), so that the copies of dabmod_b200_process_batch / _process_eti_batch into it run at PCIe speed and
 * overlap the kernels.  Optional: pageable memory works, slower.  Unregister before the memory is freed. */
int dabmod_b200_host_register(void *p, size_t bytes);
int dabmod_b200_host_unregister(void *p);

/* ======================================================================================
 * Row N1 of SURVEY.md section 8(f): the channel coding ahead of the path, i.e. the part of
 * DabModulator's graph between EtiReader and QpskSymbolMapper (src/DabModulator.cpp:131-150,
 * 286-383): PrbsGenerator -> ConvEncoder -> PuncturingEncoder [-> TimeInterleaver] for the FIC
 * and every subchannel, FrameMultiplexer, BlockPartitioner.  Input: raw ETI(NI) frames of 6144
 * bytes (what InputFileReader hands to EtiReader::loadEtiData, src/EtiReader.cpp:93-284);
 * output: the per-TF byte blocks dabmod_b200_process* consume.
 * ====================================================================================== */

/* PuncturingRule(length in encoder-output bytes, 32-bit keep mask), src/PuncturingRule.h */
typedef struct dabmod_b200_punct_rule {
    uint32_t length;
    uint32_t pattern;
} dabmod_b200_punct_rule;

/* One coded stream: [0] is the FIC (FicSource), [1..] the subchannels in STC order
 * (SubchannelSource).  The tail rule (3, 0xcccccc) of DabModulator.cpp:316,373 is implied. */
typedef struct dabmod_b200_stream {
    uint32_t framesize;   /* input bytes per ETI frame: FicSource::getFramesize / SubchannelSource::framesize */
    uint32_t out_bytes;   /* FIC: 288 (TM III: 384); subchannel: framesizeCu() * 8 */
    uint32_t start_cu;    /* SubchannelSource::startAddress (ignored for the FIC) */
    uint32_t n_rules;     /* 1..8 */
    dabmod_b200_punct_rule rules[8];   /* get_rules() */
} dabmod_b200_stream;

typedef struct dabmod_b200_coder dabmod_b200_coder;

/* Host-only: what EtiReader + FicSource + SubchannelSource derive from the header of one
 * ETI(NI) frame (EtiReader.cpp:120-190, FicSource.cpp:40-63, SubchannelSource.cpp:66-170,
 * :665-760).  EEP-A/B profiles are derived here; a UEP (short form) subchannel returns
 * DABMOD_B200_EUNSUPPORTED -- pass its rules explicitly (the C++ adapter has them from
 * SubchannelSource::get_rules()).  *mode = transmission mode 1..4 from MID. */
int dabmod_b200_eti_describe(const uint8_t *frame, size_t len, int *mode, dabmod_b200_stream *streams, int cap,
                             int *n_streams);

/* One multiplex configuration on one GPU; max_frames = largest n_frames of a call. */
int dabmod_b200_coder_create(int device, int mode, const dabmod_b200_stream *streams, int n_streams,
                             int max_frames, dabmod_b200_coder **out);
void dabmod_b200_coder_destroy(dabmod_b200_coder *c);
size_t dabmod_b200_coder_tf_bytes(const dabmod_b200_coder *c);
int dabmod_b200_coder_frames_per_tf(const dabmod_b200_coder *c);   /* BlockPartitioner d_cifCount */
/* Byte offset of stream `stream`'s data inside a 6144-byte frame (ETI(NI) layout: SYNC, FC, NST x STC, EOH, FIC,
 * then the subchannels in STC order; EtiReader.cpp:190-249).  The coder reads nothing else of a frame, so a caller
 * that holds parsed sources (FicSource / SubchannelSource buffers) may place their bytes at these offsets.
 * -1 for a bad index. */
int dabmod_b200_coder_stream_offset(const dabmod_b200_coder *c, int stream);

/* n_frames consecutive ETI frames (a multiple of frames_per_tf) -> n_frames / frames_per_tf blocks.
 * Host buffers; synchronous. */
int dabmod_b200_coder_process(dabmod_b200_coder *c, const uint8_t *eti_frames, size_t n_frames, uint8_t *bits_out,
                              size_t cap, size_t *out_bytes);
/* Same with device buffers, enqueued on `stream` (NULL = the coder's own), not synchronised. */
int dabmod_b200_coder_process_device(dabmod_b200_coder *c, const uint8_t *d_eti_frames, size_t n_frames,
                                     uint8_t *d_bits_out, void *stream);
/* Forget / set the time interleaver history (TimeInterleaver.cpp:39-41: 16 frames).  For a
 * stream sharded by frame ranges, prime with the (up to) 15 ETI frames before the shard. */
int dabmod_b200_coder_reset(dabmod_b200_coder *c);
int dabmod_b200_coder_prime(dabmod_b200_coder *c, const uint8_t *eti_frames, size_t n_frames);

/* ETI bytes in, I/Q out: coder and modulator chained on the device (the coded blocks never
 * travel to the host).  Both handles must be on the same device and transmission mode.  Same sliced
 * three-stream pipeline as dabmod_b200_process_batch: ETI frames of slice i+1 cross PCIe while slice i is in
 * the coding + modulator kernels and the I/Q of slice i-1 goes back.  This is the whole of
 * DabModulator::process (src/DabModulator.cpp:131-417) for n_frames / frames_per_tf transmission frames. */
int dabmod_b200_process_eti_batch(dabmod_b200 *h, dabmod_b200_coder *c, const uint8_t *eti_frames, size_t n_frames,
                                  void *iq_out, size_t cap, size_t *out_bytes);
/* Same, delivered to a file descriptor through the modulator's pinned ring (src/OutputFile.cpp:56-67). */
int dabmod_b200_process_eti_batch_to_fd(dabmod_b200 *h, dabmod_b200_coder *c, const uint8_t *eti_frames,
                                        size_t n_frames, int fd, size_t *out_bytes);

/* Positions coder + modulator at transmission frame `tf_index` of a stream, for sharding one ETI stream across
 * GPUs by frame ranges (BASELINE configs[4]).  `eti_before` = the `n_before` ETI frames that precede the shard
 * in the stream; the last min(tf_index * frames_per_tf, 15 + frames_per_tf) of them are used and that many are
 * required: 15 frames rebuild the time interleaver memory (src/TimeInterleaver.cpp:39-41,66-93), the
 * transmission frame before the shard is coded and re-run up to the resampler input to rebuild the Resampler
 * overlap (src/Resampler.cpp:143-145,185-191); tf_index sets the TII every-second-frame toggle
 * (src/TII.cpp:225-242).  The shard's output is then bit-identical to the same frames of an unsharded run. */
int dabmod_b200_seek_eti(dabmod_b200 *h, dabmod_b200_coder *c, uint64_t tf_index, const uint8_t *eti_before,
                         size_t n_before);

/* Thread-local description of the last coder error. Never NULL. */
const char *dabmod_b200_coder_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* DABMOD_B200_H */
