// Channel coding ahead of the COFDM path on the GPU:
// ETI(NI) frames -> energy dispersal -> convolutional encoder -> puncturing ->
// time interleaver -> CIF assembly -> transmission-frame blocks, i.e. the part of
// DabModulator's graph between EtiReader and QpskSymbolMapper
// (reference src/DabModulator.cpp:131-150, 286-383).
//
// The reference runs five sequential byte/bit loops per stream and ETI frame.  Here
// every output bit is computed on its own: after puncturing, output bit o of a stream
// is convolutional-encoder bit c(o) = 4 i + g (a table built once per configuration
// from the puncturing rules), and that bit is the parity of generator g over the seven
// scrambled input bits i-6 .. i.  No state is carried inside a frame, so a batch of ETI
// frames is one flat grid of (frame, stream, output word) items.  The only memory across
// frames is the time interleaver's (15 frames), kept as a ring of punctured frames.
#include "../../include/dabmod_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

constexpr int CIF_BYTES = 864 * 8;
constexpr int ETI_FRAME = 6144;
constexpr int TI_DEPTH = 16;             // TimeInterleaver history, frames
constexpr int MAX_STREAMS = 65;

struct CoderError : std::runtime_error {
    int code;
    CoderError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            throw CoderError(DABMOD_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// per stream, device side
struct StreamDev {
    int in_off;        // byte offset of the stream's data inside an ETI frame
    int framesize;     // input bytes
    int out_bytes;     // punctured bytes per frame
    int punct_off;     // byte offset inside a punctured row
    int map_off;       // offset (in entries) into the bit map
    int start_byte;    // start address * 8 inside the CIF (subchannels)
    int words;         // out_bytes / 4
    int word0;         // first work item (output word) of this stream within a frame
};

struct CodeParams {
    const uint8_t *eti;          // n_frames * 6144
    const StreamDev *streams;
    const uint32_t *map;         // per output bit: (input bit i << 2) | generator, 0xffffffff = padding
    const uint8_t *prbs;         // 6912 bytes
    uint8_t *punct;              // ring of rows, row_bytes each
    int n_streams, words_per_frame, row_bytes, ring_rows, ring_base, n_frames;
};

// PrbsGenerator.cpp:126-188, ConvEncoder.cpp:59-150, PuncturingEncoder.cpp:102-210.
// One warp = one 32-bit output word of one (frame, stream): lane l computes output bit l.
__global__ void __launch_bounds__(256) k_code(const __grid_constant__ CodeParams p)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total = (long long)p.n_frames * p.words_per_frame;
    if (warp >= total) return;
    const int frame = (int)(warp / p.words_per_frame);
    const int item = (int)(warp - (long long)frame * p.words_per_frame);
    // stream of this work item (streams are few: linear search, uniform across the warp)
    int s = 0;
    while (s + 1 < p.n_streams && item >= p.streams[s + 1].word0) s++;
    const StreamDev st = p.streams[s];
    const int w = item - st.word0;
    const uint8_t *in = p.eti + (size_t)frame * ETI_FRAME + st.in_off;

    const uint32_t m = __ldg(p.map + st.map_off + 32 * w + lane);
    unsigned bit = 0;
    if (m != 0xffffffffu) {
        const int i = (int)(m >> 2);                    // newest input bit of the encoder register
        const int g = (int)(m & 3u);
        // scrambled input bits i-6 .. i, MSB first; bits before the frame and the 6 tail bits are 0
        unsigned win = 0;
        const int lo = i - 6;
        const int q0 = lo >> 3;                          // may be -1 (arithmetic shift)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int q = q0 + k;
            unsigned b = 0;
            if (q >= 0 && q < st.framesize) b = (unsigned)__ldg(in + q) ^ (unsigned)__ldg(p.prbs + q);
            win = (win << 8) | b;
        }
        const int sh = 9 - (lo & 7);                     // window = bits [lo, lo+6] of the 16-bit pair
        win = (win >> sh) & 0x7fu;
        // generators 0x5b 0x79 0x65 0x5b act on a register holding the newest bit at bit 6;
        // `win` holds it at bit 0, so the masks are bit-reversed
        const unsigned rpoly = g == 1 ? 0x4fu : g == 2 ? 0x53u : 0x6du;
        bit = __popc(win & rpoly) & 1u;
    }
    const unsigned word = __brev(__ballot_sync(0xffffffffu, bit));   // lane 0 = MSB of byte 0
    if (lane == 0) {
        const int row = (p.ring_base + (TI_DEPTH - 1) + frame) % p.ring_rows;
        uint32_t *dst = reinterpret_cast<uint32_t *>(p.punct + (size_t)row * p.row_bytes + st.punct_off) + w;
        *dst = __byte_perm(word, 0, 0x0123);
    }
}

struct MuxParams {
    const uint8_t *punct;
    const StreamDev *streams;
    const uint8_t *owner;        // 864 entries: stream index of the capacity unit, 0 = filler
    const uint8_t *prbs;
    uint8_t *bits;               // n_tf * tf_bytes
    int row_bytes, ring_rows, ring_base, cif_count, fic_out, tf_bytes;
    long long total;             // n_tf * tf_bytes
};

// TimeInterleaver.cpp:51-96, FrameMultiplexer.cpp:43-91, BlockPartitioner.cpp:78-124.
// One thread = one byte of the transmission-frame block.
__global__ void __launch_bounds__(256) k_mux(const __grid_constant__ MuxParams p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.total) return;
    const int tf = (int)(idx / p.tf_bytes);
    const int o = (int)(idx - (long long)tf * p.tf_bytes);
    const int fic_total = p.cif_count * p.fic_out;
    uint8_t v;
    if (o < fic_total) {
        const int part = o / p.fic_out, j = o - part * p.fic_out;
        const int row = (p.ring_base + (TI_DEPTH - 1) + tf * p.cif_count + part) % p.ring_rows;
        v = p.punct[(size_t)row * p.row_bytes + j];                 // the FIC is stream 0 at offset 0
    }
    else {
        const int c = (o - fic_total) / CIF_BYTES, b = (o - fic_total) - c * CIF_BYTES;
        const int s = p.owner[b >> 3];
        if (s == 0) {
            v = p.prbs[b];
        }
        else {
            const StreamDev st = p.streams[s];
            const int j = b - st.start_byte;
            const int newest = p.ring_base + (TI_DEPTH - 1) + tf * p.cif_count + c;
            // bit 7..0 of byte j come from the frames 0,8,4,12,2,10,6,14 (+1 for odd j) calls back
            const int odd = j & 1;
            unsigned acc = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int d = ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | odd;
                const int row = (newest - d) % p.ring_rows;
                acc |= p.punct[(size_t)row * p.row_bytes + st.punct_off + j] & (0x80u >> k);
            }
            v = (uint8_t)acc;
        }
    }
    p.bits[idx] = v;
}

thread_local std::string g_coder_error;

int guard(const std::function<void()> &fn)
{
    try {
        fn();
        return DABMOD_B200_OK;
    }
    catch (const CoderError &e) {
        g_coder_error = e.what();
        return e.code;
    }
    catch (const std::exception &e) {
        g_coder_error = e.what();
        return DABMOD_B200_EINVAL;
    }
}

// PuncturingEncoder::adjust_item_size, PuncturingEncoder.cpp:60-78
long punct_bits(const dabmod_b200_stream &s)
{
    long bits = 0;
    for (uint32_t r = 0; r < s.n_rules; r++) bits += (long)(s.rules[r].length / 4) * __builtin_popcount(s.rules[r].pattern);
    return bits + __builtin_popcount(0xcccccc);
}

const uint32_t PI_MASK[25] = {0,
    0xc8888888, 0xc888c888, 0xc8c8c888, 0xc8c8c8c8, 0xccc8c8c8, 0xccc8ccc8, 0xccccccc8, 0xcccccccc,
    0xeccccccc, 0xeccceccc, 0xecececcc, 0xecececec, 0xeeececec, 0xeeeceeec, 0xeeeeeeec, 0xeeeeeeee,
    0xfeeeeeee, 0xfeeefeee, 0xfefefeee, 0xfefefefe, 0xfffefefe, 0xfffefffe, 0xfffffffe, 0xffffffff};

} // namespace

struct dabmod_b200_coder {
    int device = 0, mode = 1, cif_count = 4, fic_out = 288, tf_bytes = 28800;
    int n_streams = 0, max_frames = 0, row_bytes = 0, ring_rows = 0, ring_base = 0, words_per_frame = 0;
    std::vector<StreamDev> streams;
    std::mutex mtx;
    cudaStream_t stream = nullptr;
    StreamDev *d_streams = nullptr;
    uint32_t *d_map = nullptr;
    uint8_t *d_prbs = nullptr, *d_owner = nullptr, *d_punct = nullptr, *d_eti = nullptr, *d_bits = nullptr;
};

extern "C" {

const char *dabmod_b200_coder_last_error(void) { return g_coder_error.c_str(); }

int dabmod_b200_eti_describe(const uint8_t *frame, size_t len, int *mode, dabmod_b200_stream *st, int cap,
                             int *n_streams)
{
    return guard([&] {
        if (!frame || !mode || !st || !n_streams) throw CoderError(DABMOD_B200_EINVAL, "null argument");
        if (len < 12) throw CoderError(DABMOD_B200_EINVAL, "ETI frame too short");
        // Eti.h:56-80: FC = FCT | NST:7 FICF:1 | FL_high:3 MID:2 FP:3 | FL_low; STC = SAD_high:2 SCID:6 | SAD_low | STL_high:2 TPL:6 | STL_low
        const unsigned nst = frame[5] & 0x7f, ficf = frame[5] >> 7, mid = (frame[6] >> 3) & 3;
        if (!ficf) throw CoderError(DABMOD_B200_EINVAL, "FIC must be present to modulate!");   // EtiReader.cpp:143-145
        if ((int)nst + 1 > cap || len < 8 + 4 * (size_t)nst) throw CoderError(DABMOD_B200_EINVAL, "stream array too small");
        *mode = mid == 0 ? 4 : (int)mid;
        std::memset(st, 0, sizeof(*st) * (nst + 1));
        // FicSource.cpp:40-63
        st[0].framesize = mid == 3 ? 128 : 96;
        st[0].n_rules = 2;
        st[0].rules[0] = {(mid == 3 ? 29u : 21u) * 16u, 0xeeeeeeeeu};
        st[0].rules[1] = {3u * 16u, 0xeeeeeeecu};
        st[0].out_bytes = (uint32_t)((punct_bits(st[0]) + 7) / 8);
        for (unsigned i = 0; i < nst; i++) {
            const uint8_t *c = frame + 8 + 4 * i;
            dabmod_b200_stream &s = st[1 + i];
            const unsigned tpl = c[2] >> 2;
            s.start_cu = ((c[0] & 3u) << 8) | c[1];
            s.framesize = (((c[2] & 3u) << 8) | c[3]) * 8;
            const unsigned br = s.framesize / 3;            // SubchannelSource::bitrate
            if (!((tpl >> 5) & 1))
                throw CoderError(DABMOD_B200_EUNSUPPORTED,
                                 "subchannel " + std::to_string(i) + " uses a UEP (short form) profile: pass its "
                                 "puncturing rules explicitly (SubchannelSource::get_rules)");
            const unsigned opt = (tpl >> 2) & 7, lvl = (tpl & 3) + 1;
            s.n_rules = 2;
            if (opt == 0) {           // EEP-A, SubchannelSource.cpp:84-121, :690-708
                static const unsigned cu8[4] = {12, 8, 6, 4};
                s.out_bytes = (br / 8) * cu8[lvl - 1] * 8;
                if (lvl == 1) { s.rules[0] = {((6 * br / 8) - 3) * 16, PI_MASK[24]}; s.rules[1] = {3 * 16, PI_MASK[23]}; }
                else if (lvl == 2 && br == 8) { s.rules[0] = {5 * 16, PI_MASK[13]}; s.rules[1] = {1 * 16, PI_MASK[12]}; }
                else if (lvl == 2) { s.rules[0] = {((2 * br / 8) - 3) * 16, PI_MASK[14]}; s.rules[1] = {((4 * br / 8) + 3) * 16, PI_MASK[13]}; }
                else if (lvl == 3) { s.rules[0] = {((6 * br / 8) - 3) * 16, PI_MASK[8]}; s.rules[1] = {3 * 16, PI_MASK[7]}; }
                else { s.rules[0] = {((4 * br / 8) - 3) * 16, PI_MASK[3]}; s.rules[1] = {((2 * br / 8) + 3) * 16, PI_MASK[2]}; }
            }
            else if (opt == 1) {      // EEP-B, SubchannelSource.cpp:122-153, :672-689
                static const unsigned cu32[4] = {27, 21, 18, 15};
                static const int pa[4] = {10, 6, 4, 2}, pb[4] = {9, 5, 3, 1};
                s.out_bytes = (br / 32) * cu32[lvl - 1] * 8;
                s.rules[0] = {((24 * br / 32) - 3) * 16, PI_MASK[pa[lvl - 1]]};
                s.rules[1] = {3 * 16, PI_MASK[pb[lvl - 1]]};
            }
            else throw CoderError(DABMOD_B200_EINVAL, "SubchannelSource unknown protection option!");
        }
        *n_streams = (int)nst + 1;
    });
}

void dabmod_b200_coder_destroy(dabmod_b200_coder *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    cudaFree(c->d_streams); cudaFree(c->d_map); cudaFree(c->d_prbs); cudaFree(c->d_owner);
    cudaFree(c->d_punct); cudaFree(c->d_eti); cudaFree(c->d_bits);
    delete c;
}

int dabmod_b200_coder_create(int device, int mode, const dabmod_b200_stream *st, int n_streams, int max_frames,
                             dabmod_b200_coder **out)
{
    if (out) *out = nullptr;
    dabmod_b200_coder *c = nullptr;
    int rc = guard([&] {
        if (!st || !out) throw CoderError(DABMOD_B200_EINVAL, "null argument");
        if (mode < 1 || mode > 4) throw CoderError(DABMOD_B200_EINVAL, "BlockPartitioner::BlockPartitioner invalid mode");
        if (n_streams < 1 || n_streams > MAX_STREAMS) throw CoderError(DABMOD_B200_EINVAL, "bad stream count");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw CoderError(DABMOD_B200_ECUDA, "no CUDA device (this library has no CPU fallback)");
        if (device < 0 || device >= ndev) throw CoderError(DABMOD_B200_EINVAL, "invalid device ordinal");
        CK(cudaSetDevice(device));
        c = new dabmod_b200_coder();
        c->device = device;
        c->mode = mode;
        c->cif_count = mode == 1 ? 4 : mode == 4 ? 2 : 1;       // BlockPartitioner.cpp:44-73
        c->fic_out = mode == 3 ? 384 : 288;
        c->tf_bytes = c->cif_count * (c->fic_out + CIF_BYTES);
        c->n_streams = n_streams;
        c->max_frames = std::max(max_frames, c->cif_count);

        std::vector<uint32_t> map;
        std::vector<uint8_t> owner(864, 0);
        int in_off = 8 + 4 * (n_streams - 1) + 4, punct_off = 0, word0 = 0;
        for (int s = 0; s < n_streams; s++) {
            const dabmod_b200_stream &d = st[s];
            if (d.n_rules < 1 || d.n_rules > 8) throw CoderError(DABMOD_B200_EINVAL, "stream needs 1..8 puncturing rules");
            if (d.framesize == 0 || d.out_bytes == 0 || (d.out_bytes & 3))
                throw CoderError(DABMOD_B200_EINVAL, "invalid stream size");
            if (in_off + (int)d.framesize + 8 > ETI_FRAME) throw CoderError(DABMOD_B200_EINVAL, "streams exceed the ETI frame");
            if (s == 0 && (int)d.out_bytes != c->fic_out)
                throw CoderError(DABMOD_B200_EINVAL, "BlockPartitioner::process input 0 size not valid!");
            if (s > 0 && d.start_cu * 8 + d.out_bytes > (uint32_t)CIF_BYTES)
                throw CoderError(DABMOD_B200_EINVAL, "subchannel exceeds the CIF");
            // Expand the puncturing rules (PuncturingEncoder.cpp:148-196) into one source index per
            // output bit: the index of the kept convolutional-encoder bit, 4 * input bit + generator.
            const long body = 4L * d.framesize;                  // encoder bytes before the 3 tail bytes
            const size_t base = map.size();
            const long cap_bits = (long)d.out_bytes * 8;
            map.resize(base + (size_t)cap_bits, 0xffffffffu);
            long ob = 0, ic = 0;
            uint32_t r = 0;
            while (ic < body) {
                if (d.rules[r].length == 0 || (d.rules[r].length & 3))
                    throw CoderError(DABMOD_B200_EINVAL, "puncturing rule length must be a positive multiple of 4");
                for (long len = d.rules[r].length; len > 0 && ic < body; len -= 4, ic += 4)
                    for (int k = 0; k < 32; k++)
                        if (d.rules[r].pattern & (0x80000000u >> k)) {
                            if (ob < cap_bits) map[base + ob] = (uint32_t)(ic * 8 + k);
                            ob++;
                        }
                if (++r == d.n_rules) r = 0;
            }
            for (int k = 0; k < 24; k++)                          // tail rule (3, 0xcccccc), DabModulator.cpp:316,373
                if (0xccccccu & (0x800000u >> k)) {
                    if (ob < cap_bits) map[base + ob] = (uint32_t)(ic * 8 + k);
                    ob++;
                }
            // PuncturingEncoder.cpp:120-134: the kept bits must fill the block (UEP: one byte of padding allowed)
            const long need = (ob + 7) / 8;
            if (!(need == (long)d.out_bytes || (s > 0 && need + 1 == (long)d.out_bytes)))
                throw CoderError(DABMOD_B200_EINVAL, "PuncturingEncoder encoder initialisation failed. block_size: " +
                                                         std::to_string(need) + " out_bytes: " + std::to_string(d.out_bytes));
            StreamDev sd{};
            sd.in_off = in_off;
            sd.framesize = (int)d.framesize;
            sd.out_bytes = (int)d.out_bytes;
            sd.punct_off = punct_off;
            sd.map_off = (int)base;
            sd.start_byte = (int)d.start_cu * 8;
            sd.words = (int)d.out_bytes / 4;
            sd.word0 = word0;
            c->streams.push_back(sd);
            if (s > 0)
                for (uint32_t cu = d.start_cu; cu < d.start_cu + d.out_bytes / 8; cu++) owner[cu] = (uint8_t)s;
            in_off += (int)d.framesize;
            punct_off += (int)d.out_bytes;
            word0 += sd.words;
        }
        c->row_bytes = punct_off;
        c->words_per_frame = word0;
        c->ring_rows = (TI_DEPTH - 1) + c->max_frames;
        c->ring_base = 0;

        // PrbsGenerator(., 0x110), PrbsGenerator.cpp:126-188: x^9 + x^5 + 1 from the all-ones state
        std::vector<uint8_t> prbs(CIF_BYTES);
        unsigned reg = 0x1ff;
        for (auto &b : prbs) {
            unsigned v = 0;
            for (int k = 0; k < 8; k++) {
                const unsigned nb = ((reg >> 8) ^ (reg >> 4)) & 1u;
                reg = ((reg << 1) | nb) & 0x1ffu;
                v = (v << 1) | nb;
            }
            b = (uint8_t)v;
        }

        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CK(cudaMalloc((void **)&c->d_streams, sizeof(StreamDev) * c->streams.size()));
        CK(cudaMalloc((void **)&c->d_map, sizeof(uint32_t) * map.size()));
        CK(cudaMalloc((void **)&c->d_prbs, prbs.size()));
        CK(cudaMalloc((void **)&c->d_owner, owner.size()));
        CK(cudaMalloc((void **)&c->d_punct, (size_t)c->ring_rows * c->row_bytes));
        CK(cudaMalloc((void **)&c->d_eti, (size_t)c->max_frames * ETI_FRAME));
        CK(cudaMalloc((void **)&c->d_bits, (size_t)(c->max_frames / c->cif_count) * c->tf_bytes));
        CK(cudaMemcpy(c->d_streams, c->streams.data(), sizeof(StreamDev) * c->streams.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_map, map.data(), sizeof(uint32_t) * map.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_prbs, prbs.data(), prbs.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_owner, owner.data(), owner.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(c->d_punct, 0, (size_t)c->ring_rows * c->row_bytes));
        *out = c;
    });
    if (rc != DABMOD_B200_OK && c) dabmod_b200_coder_destroy(c);
    return rc;
}

size_t dabmod_b200_coder_tf_bytes(const dabmod_b200_coder *c) { return c ? (size_t)c->tf_bytes : 0; }
int dabmod_b200_coder_frames_per_tf(const dabmod_b200_coder *c) { return c ? c->cif_count : 0; }

// Encodes n_frames device-resident ETI frames into the ring (and, unless bits == nullptr, assembles
// the transmission-frame blocks).  Caller holds the lock.
static void coder_enqueue(dabmod_b200_coder *c, const uint8_t *d_eti, size_t n_frames, uint8_t *d_bits, cudaStream_t s)
{
    CodeParams cp{};
    cp.eti = d_eti;
    cp.streams = c->d_streams;
    cp.map = c->d_map;
    cp.prbs = c->d_prbs;
    cp.punct = c->d_punct;
    cp.n_streams = c->n_streams;
    cp.words_per_frame = c->words_per_frame;
    cp.row_bytes = c->row_bytes;
    cp.ring_rows = c->ring_rows;
    cp.ring_base = c->ring_base;
    cp.n_frames = (int)n_frames;
    const long long warps = (long long)n_frames * c->words_per_frame;
    k_code<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(cp);
    CK(cudaGetLastError());
    if (d_bits) {
        MuxParams mp{};
        mp.punct = c->d_punct;
        mp.streams = c->d_streams;
        mp.owner = c->d_owner;
        mp.prbs = c->d_prbs;
        mp.bits = d_bits;
        mp.row_bytes = c->row_bytes;
        mp.ring_rows = c->ring_rows;
        mp.ring_base = c->ring_base + c->ring_rows;      // keeps (newest - d) non-negative before the modulo
        mp.cif_count = c->cif_count;
        mp.fic_out = c->fic_out;
        mp.tf_bytes = c->tf_bytes;
        mp.total = (long long)(n_frames / c->cif_count) * c->tf_bytes;
        k_mux<<<(unsigned)((mp.total + 255) / 256), 256, 0, s>>>(mp);
        CK(cudaGetLastError());
    }
    c->ring_base = (int)((c->ring_base + n_frames) % c->ring_rows);
}

int dabmod_b200_coder_process_device(dabmod_b200_coder *c, const uint8_t *d_eti, size_t n_frames, uint8_t *d_bits,
                                     void *stream)
{
    return guard([&] {
        if (!c || (n_frames && (!d_eti || !d_bits))) throw CoderError(DABMOD_B200_EINVAL, "null argument");
        if (n_frames > (size_t)c->max_frames) throw CoderError(DABMOD_B200_EINVAL, "n_frames exceeds max_frames of the coder");
        if (n_frames % c->cif_count)
            throw CoderError(DABMOD_B200_ESTATE, "a call must carry whole transmission frames (" +
                                                     std::to_string(c->cif_count) + " ETI frames each)");
        if (n_frames == 0) return;
        std::lock_guard<std::mutex> lock(c->mtx);
        CK(cudaSetDevice(c->device));
        coder_enqueue(c, d_eti, n_frames, d_bits, stream ? (cudaStream_t)stream : c->stream);
    });
}

int dabmod_b200_coder_process(dabmod_b200_coder *c, const uint8_t *eti, size_t n_frames, uint8_t *bits, size_t cap,
                              size_t *out_bytes)
{
    if (out_bytes) *out_bytes = 0;
    return guard([&] {
        if (!c || (n_frames && (!eti || !bits))) throw CoderError(DABMOD_B200_EINVAL, "null argument");
        if (n_frames > (size_t)c->max_frames) throw CoderError(DABMOD_B200_EINVAL, "n_frames exceeds max_frames of the coder");
        if (n_frames % c->cif_count)
            throw CoderError(DABMOD_B200_ESTATE, "a call must carry whole transmission frames (" +
                                                     std::to_string(c->cif_count) + " ETI frames each)");
        const size_t nb = (n_frames / c->cif_count) * (size_t)c->tf_bytes;
        if (cap < nb) throw CoderError(DABMOD_B200_EINVAL, "output buffer too small");
        if (n_frames == 0) return;
        std::lock_guard<std::mutex> lock(c->mtx);
        CK(cudaSetDevice(c->device));
        CK(cudaMemcpyAsync(c->d_eti, eti, n_frames * ETI_FRAME, cudaMemcpyHostToDevice, c->stream));
        coder_enqueue(c, c->d_eti, n_frames, c->d_bits, c->stream);
        CK(cudaMemcpyAsync(bits, c->d_bits, nb, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (out_bytes) *out_bytes = nb;
    });
}

int dabmod_b200_coder_prime(dabmod_b200_coder *c, const uint8_t *eti, size_t n_frames)
{
    return guard([&] {
        if (!c || (n_frames && !eti)) throw CoderError(DABMOD_B200_EINVAL, "null argument");
        std::lock_guard<std::mutex> lock(c->mtx);
        CK(cudaSetDevice(c->device));
        CK(cudaMemsetAsync(c->d_punct, 0, (size_t)c->ring_rows * c->row_bytes, c->stream));
        c->ring_base = 0;
        // only the last 15 frames matter (TimeInterleaver.cpp:39-41)
        if (n_frames > (size_t)(TI_DEPTH - 1)) { eti += (n_frames - (TI_DEPTH - 1)) * ETI_FRAME; n_frames = TI_DEPTH - 1; }
        for (size_t done = 0; done < n_frames;) {
            const size_t n = std::min<size_t>(n_frames - done, (size_t)c->max_frames);
            CK(cudaMemcpyAsync(c->d_eti, eti + done * ETI_FRAME, n * ETI_FRAME, cudaMemcpyHostToDevice, c->stream));
            coder_enqueue(c, c->d_eti, n, nullptr, c->stream);
            CK(cudaStreamSynchronize(c->stream));
            done += n;
        }
        CK(cudaStreamSynchronize(c->stream));
    });
}

int dabmod_b200_process_eti_batch(dabmod_b200 *h, dabmod_b200_coder *c, const uint8_t *eti, size_t n_frames,
                                  void *iq_out, size_t cap, size_t *out_bytes)
{
    if (out_bytes) *out_bytes = 0;
    int rc = guard([&] {
        if (!h || !c || (n_frames && (!eti || !iq_out))) throw CoderError(DABMOD_B200_EINVAL, "null argument");
        if (n_frames > (size_t)c->max_frames) throw CoderError(DABMOD_B200_EINVAL, "n_frames exceeds max_frames of the coder");
        if (n_frames % c->cif_count) throw CoderError(DABMOD_B200_ESTATE, "a call must carry whole transmission frames");
        if ((size_t)c->tf_bytes != dabmod_b200_tf_in_bytes(h))
            throw CoderError(DABMOD_B200_EINVAL, "coder and modulator are configured for different transmission modes");
        const size_t n_tf = n_frames / c->cif_count;
        const size_t nb = n_tf * dabmod_b200_tf_out_bytes(h);
        if (cap < nb) throw CoderError(DABMOD_B200_EINVAL, "output buffer too small");
        if (n_frames == 0) return;
        void *d_iq = dabmod_b200_device_out(h);
        if (!d_iq) throw CoderError(DABMOD_B200_ESTATE, "modulator has no device output buffer");
        std::lock_guard<std::mutex> lock(c->mtx);
        CK(cudaSetDevice(c->device));
        CK(cudaMemcpyAsync(c->d_eti, eti, n_frames * ETI_FRAME, cudaMemcpyHostToDevice, c->stream));
        coder_enqueue(c, c->d_eti, n_frames, c->d_bits, c->stream);
        // the coded blocks never leave the device: the modulator kernels follow on the same stream
        if (dabmod_b200_process_batch_device(h, c->d_bits, n_tf, d_iq, c->stream) != DABMOD_B200_OK)
            throw CoderError(DABMOD_B200_EINVAL, dabmod_b200_last_error());
        CK(cudaMemcpyAsync(iq_out, d_iq, nb, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (out_bytes) *out_bytes = nb;
    });
    return rc;
}

int dabmod_b200_coder_reset(dabmod_b200_coder *c)
{
    return guard([&] {
        if (!c) throw CoderError(DABMOD_B200_EINVAL, "null handle");
        std::lock_guard<std::mutex> lock(c->mtx);
        CK(cudaSetDevice(c->device));
        CK(cudaMemsetAsync(c->d_punct, 0, (size_t)c->ring_rows * c->row_bytes, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->ring_base = 0;
    });
}

} // extern "C"
