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
#include "internal.h"

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

using CoderError = dabmod::ApiError;     // one error type across the library, so that codes survive the chain

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            throw CoderError(DABMOD_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

constexpr int MAX_SEGMENTS = 16;
constexpr int CODE_THREADS = 128;

// One application of a puncturing rule: `n_groups` groups of 4 encoder bytes (= 8 input bits)
// starting at group `first_group`, each keeping the `kept` bits selected by `mask`.
struct Segment {
    int first_group, n_groups;
    uint32_t mask;
    int kept;          // popcount(mask)
    int out_bit;       // output bit offset of the segment's first kept bit
};

// per stream, device side
struct StreamDev {
    int in_off;        // byte offset of the stream's data inside an ETI frame
    int framesize;     // input bytes = number of full groups
    int out_bytes;     // punctured bytes per frame
    int punct_off;     // byte offset inside a punctured row
    int start_byte;    // start address * 8 inside the CIF (subchannels)
    int n_segments;
    int tail_out_bit;  // output bit offset of the tail rule (3, 0xcccccc)
    Segment seg[MAX_SEGMENTS];
};

struct CodeParams {
    const uint8_t *eti;          // n_frames * 6144
    const StreamDev *streams;
    const uint8_t *prbs;         // 6912 bytes
    uint8_t *punct;              // ring of rows, row_bytes each
    int n_streams, row_bytes, ring_rows, ring_base, n_frames;
};

// bit t of the byte c -> bit 4t of the result
__device__ __forceinline__ uint32_t spread_byte(uint32_t c)
{
    c = (c | (c << 12)) & 0x000f000fu;
    c = (c | (c << 6)) & 0x03030303u;
    c = (c | (c << 3)) & 0x11111111u;
    return c;
}

// The 32 encoder output bits of the 8 input bits in the low byte of `w` (bits 13..8 of `w` = the six
// input bits before them): ConvEncoder.cpp:88-108.  Generators 0x5b 0x79 0x65 0x5b act on a register
// that holds the newest bit at bit 6, i.e. on input ages {0,2,3,5,6}, {0,1,2,3,6}, {0,1,4,6}.
// Result: first input bit in the top nibble, generator 0 in the nibble's MSB.
__device__ __forceinline__ uint32_t conv8(uint32_t w)
{
    const uint32_t c0 = (w ^ (w >> 2) ^ (w >> 3) ^ (w >> 5) ^ (w >> 6)) & 0xffu;
    const uint32_t c1 = (w ^ (w >> 1) ^ (w >> 2) ^ (w >> 3) ^ (w >> 6)) & 0xffu;
    const uint32_t c2 = (w ^ (w >> 1) ^ (w >> 4) ^ (w >> 6)) & 0xffu;
    const uint32_t s0 = spread_byte(c0);
    return (s0 << 3) | (spread_byte(c1) << 2) | (spread_byte(c2) << 1) | s0;
}

// the bits of `v` selected by `mask`, packed MSB first (PuncturingEncoder.cpp:152-166)
__device__ __forceinline__ uint32_t extract_bits(uint32_t v, uint32_t mask)
{
    uint32_t out = 0;
    while (mask) {
        const int b = 31 - __clz(mask);
        out = (out << 1) | ((v >> b) & 1u);
        mask &= ~(1u << b);
    }
    return out;
}

// ORs the `n` low bits of `v` into the big-endian bit string `buf` at bit offset `o`
__device__ __forceinline__ void put_bits(uint32_t *buf, int o, uint32_t v, int n, int cap_bits)
{
    if (n == 0 || o >= cap_bits) return;
    const int j = o >> 5, r = o & 31;
    if (r + n <= 32) {
        atomicOr(buf + j, v << (32 - r - n));
    }
    else {
        atomicOr(buf + j, v >> (r + n - 32));
        if (((j + 1) << 5) < cap_bits) atomicOr(buf + j + 1, v << (64 - r - n));
    }
}

// PrbsGenerator.cpp:126-188, ConvEncoder.cpp:59-150, PuncturingEncoder.cpp:102-210.
// One CTA = one (frame, stream); one thread = one group of 8 input bits: 32 encoder bits by
// shift-and-XOR, the kept ones packed and OR-ed into the row at the position the rule table gives.
__global__ void __launch_bounds__(CODE_THREADS) k_code(const __grid_constant__ CodeParams p)
{
    __shared__ uint32_t row[CIF_BYTES / 4];
    __shared__ uint8_t scr[ETI_FRAME];
    const int frame = blockIdx.x / p.n_streams, s = blockIdx.x - frame * p.n_streams;
    const StreamDev &st = p.streams[s];
    const int n = st.framesize, words = st.out_bytes / 4, cap_bits = st.out_bytes * 8;
    const uint8_t *in = p.eti + (size_t)frame * ETI_FRAME + st.in_off;
    for (int i = threadIdx.x; i < n; i += CODE_THREADS) scr[i] = __ldg(in + i) ^ __ldg(p.prbs + i);
    for (int i = threadIdx.x; i < words; i += CODE_THREADS) row[i] = 0;
    __syncthreads();
    for (int g = threadIdx.x; g <= n; g += CODE_THREADS) {
        const uint32_t prev = g > 0 ? scr[g - 1] : 0u;
        if (g < n) {
            const uint32_t c = conv8((prev << 8) | scr[g]);
            int k = 0;
            while (k + 1 < st.n_segments && g >= st.seg[k + 1].first_group) k++;
            const Segment sg = st.seg[k];
            put_bits(row, sg.out_bit + (g - sg.first_group) * sg.kept, extract_bits(c, sg.mask), sg.kept, cap_bits);
        }
        else {
            // six zero tail bits -> 24 encoder bits, tail rule 0xcccccc keeps 12
            const uint32_t c = conv8(prev << 8) >> 8;
            put_bits(row, st.tail_out_bit, extract_bits(c, 0xccccccu), 12, cap_bits);
        }
    }
    __syncthreads();
    const int r = (p.ring_base + (TI_DEPTH - 1) + frame) % p.ring_rows;
    uint32_t *dst = reinterpret_cast<uint32_t *>(p.punct + (size_t)r * p.row_bytes + st.punct_off);
    for (int i = threadIdx.x; i < words; i += CODE_THREADS) dst[i] = __byte_perm(row[i], 0, 0x0123);
}

struct MuxParams {
    const uint8_t *punct;
    const StreamDev *streams;
    const uint8_t *owner;        // 864 entries: stream index of the capacity unit, 0 = filler
    const uint8_t *prbs;
    uint8_t *bits;               // n_tf * tf_bytes
    int row_bytes, ring_rows, ring_base, cif_count, fic_out, tf_bytes;
    long long total_words;       // n_tf * tf_bytes / 4
};

// TimeInterleaver.cpp:51-96, FrameMultiplexer.cpp:43-91, BlockPartitioner.cpp:78-124.
// One thread = four bytes of the transmission-frame block (all sizes and offsets are multiples of 8).
__global__ void __launch_bounds__(256) k_mux(const __grid_constant__ MuxParams p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.total_words) return;
    const int wpt = p.tf_bytes / 4;
    const int tf = (int)(idx / wpt);
    const int o = (int)(idx - (long long)tf * wpt) * 4;
    const int fic_total = p.cif_count * p.fic_out;
    uint32_t v;
    if (o < fic_total) {
        const int part = o / p.fic_out, j = o - part * p.fic_out;
        const int row = (p.ring_base + (TI_DEPTH - 1) + tf * p.cif_count + part) % p.ring_rows;
        v = *reinterpret_cast<const uint32_t *>(p.punct + (size_t)row * p.row_bytes + j);   // FIC = stream 0 at offset 0
    }
    else {
        const int c = (o - fic_total) / CIF_BYTES, b = (o - fic_total) - c * CIF_BYTES;
        const int s = p.owner[b >> 3];
        if (s == 0) {
            v = *reinterpret_cast<const uint32_t *>(p.prbs + b);
        }
        else {
            const StreamDev &st = p.streams[s];
            const int j = b - st.start_byte;             // multiple of 4: bytes j, j+2 even, j+1, j+3 odd
            const int newest = p.ring_base + (TI_DEPTH - 1) + tf * p.cif_count + c;
            // bit 7..0 of a byte come from the frames 0,8,4,12,2,10,6,14 (+1 for odd bytes) calls back
            v = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int d = ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1);
                const int r0 = (newest - d) % p.ring_rows, r1 = (newest - d - 1) % p.ring_rows;
                const uint32_t e = *reinterpret_cast<const uint32_t *>(p.punct + (size_t)r0 * p.row_bytes + st.punct_off + j);
                const uint32_t q = *reinterpret_cast<const uint32_t *>(p.punct + (size_t)r1 * p.row_bytes + st.punct_off + j);
                v |= (e & (0x00800080u >> k)) | (q & (0x80008000u >> k));
            }
        }
    }
    reinterpret_cast<uint32_t *>(p.bits)[idx] = v;
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
    int n_streams = 0, max_frames = 0, row_bytes = 0, ring_rows = 0, ring_base = 0;
    std::vector<StreamDev> streams;
    std::mutex mtx;
    cudaStream_t stream = nullptr;
    StreamDev *d_streams = nullptr;
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
        // InputFileReader.cpp:84 accepts a frame by its FSYNC word; EtiReader then consumes STC, EOH, the FIC and
        // every stream (EtiReader.cpp:190-249) -- a frame shorter than that is a truncated read.
        const uint32_t sync = (uint32_t)frame[1] | ((uint32_t)frame[2] << 8) | ((uint32_t)frame[3] << 16);
        if (sync != 0xb63a07u && sync != 0x49c5f8u) throw CoderError(DABMOD_B200_EINVAL, "ETI frame without FSYNC");
        size_t need = 8 + 4 * (size_t)nst + 4 + st[0].framesize;
        for (unsigned i = 0; i < nst; i++) need += st[1 + i].framesize;
        if (len < need)
            throw CoderError(DABMOD_B200_EINVAL, "ETI frame too short for its header: " + std::to_string(len) + " < " +
                                                     std::to_string(need));
    });
}

void dabmod_b200_coder_destroy(dabmod_b200_coder *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    cudaFree(c->d_streams); cudaFree(c->d_prbs); cudaFree(c->d_owner);
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

        std::vector<uint8_t> owner(864, 0);
        int in_off = 8 + 4 * (n_streams - 1) + 4, punct_off = 0;
        for (int s = 0; s < n_streams; s++) {
            const dabmod_b200_stream &d = st[s];
            if (d.n_rules < 1 || d.n_rules > 8) throw CoderError(DABMOD_B200_EINVAL, "stream needs 1..8 puncturing rules");
            if (d.framesize == 0 || d.out_bytes == 0 || (d.out_bytes & 7) || d.out_bytes > (uint32_t)CIF_BYTES)
                throw CoderError(DABMOD_B200_EINVAL, "invalid stream size");
            if (in_off + (int)d.framesize + 8 > ETI_FRAME) throw CoderError(DABMOD_B200_EINVAL, "streams exceed the ETI frame");
            if (s == 0 && (int)d.out_bytes != c->fic_out)
                throw CoderError(DABMOD_B200_EINVAL, "BlockPartitioner::process input 0 size not valid!");
            if (s > 0 && d.start_cu * 8 + d.out_bytes > (uint32_t)CIF_BYTES)
                throw CoderError(DABMOD_B200_EINVAL, "subchannel exceeds the CIF");
            // The puncturing rules (PuncturingEncoder.cpp:148-196) as a table of segments, each covering
            // length/4 groups of 4 encoder bytes.  The reference sizes its input block from the rules
            // (adjust_item_size, :55-78) and throws "wrong input size" (:137-140) unless they cover the
            // ConvEncoder output exactly: every rule is applied once, over its full length.
            StreamDev sd{};
            const long groups = d.framesize;                     // 4 encoder bytes per input byte
            long g = 0, ob = 0;
            for (uint32_t r = 0; r < d.n_rules; r++) {
                if (d.rules[r].length == 0 || (d.rules[r].length & 3))
                    throw CoderError(DABMOD_B200_EINVAL, "puncturing rule length must be a positive multiple of 4");
                Segment &sg = sd.seg[sd.n_segments++];
                sg.first_group = (int)g;
                sg.n_groups = (int)(d.rules[r].length / 4);
                sg.mask = d.rules[r].pattern;
                sg.kept = __builtin_popcount(sg.mask);
                sg.out_bit = (int)ob;
                g += sg.n_groups;
                ob += (long)sg.n_groups * sg.kept;
            }
            if (g != groups)
                throw CoderError(DABMOD_B200_EINVAL, "PuncturingEncoder::process wrong input size: the rules of stream " +
                                                         std::to_string(s) + " cover " + std::to_string(4 * g + 3) +
                                                         " encoder bytes, the ConvEncoder emits " + std::to_string(4 * groups + 3));
            sd.tail_out_bit = (int)ob;
            ob += 12;                                            // tail rule (3, 0xcccccc), DabModulator.cpp:316,373
            // PuncturingEncoder.cpp:120-134: the kept bits must fill the block (UEP: one byte of padding allowed)
            const long need = (ob + 7) / 8;
            if (!(need == (long)d.out_bytes || (s > 0 && need + 1 == (long)d.out_bytes)))
                throw CoderError(DABMOD_B200_EINVAL, "PuncturingEncoder encoder initialisation failed. block_size: " +
                                                         std::to_string(need) + " out_bytes: " + std::to_string(d.out_bytes));
            sd.in_off = in_off;
            sd.framesize = (int)d.framesize;
            sd.out_bytes = (int)d.out_bytes;
            sd.punct_off = punct_off;
            sd.start_byte = (int)d.start_cu * 8;
            c->streams.push_back(sd);
            if (s > 0)
                for (uint32_t cu = d.start_cu; cu < d.start_cu + d.out_bytes / 8; cu++) owner[cu] = (uint8_t)s;
            in_off += (int)d.framesize;
            punct_off += (int)d.out_bytes;
        }
        c->row_bytes = punct_off;
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
        CK(cudaMalloc((void **)&c->d_prbs, prbs.size()));
        CK(cudaMalloc((void **)&c->d_owner, owner.size()));
        CK(cudaMalloc((void **)&c->d_punct, (size_t)c->ring_rows * c->row_bytes));
        CK(cudaMalloc((void **)&c->d_eti, (size_t)c->max_frames * ETI_FRAME));
        CK(cudaMalloc((void **)&c->d_bits, (size_t)(c->max_frames / c->cif_count) * c->tf_bytes));
        CK(cudaMemcpy(c->d_streams, c->streams.data(), sizeof(StreamDev) * c->streams.size(), cudaMemcpyHostToDevice));
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
int dabmod_b200_coder_stream_offset(const dabmod_b200_coder *c, int stream)
{
    return c && stream >= 0 && stream < (int)c->streams.size() ? c->streams[stream].in_off : -1;
}

// Encodes n_frames device-resident ETI frames into the ring (and, unless bits == nullptr, assembles
// the transmission-frame blocks).  Caller holds the lock.
static void coder_enqueue(dabmod_b200_coder *c, const uint8_t *d_eti, size_t n_frames, uint8_t *d_bits, cudaStream_t s)
{
    CodeParams cp{};
    cp.eti = d_eti;
    cp.streams = c->d_streams;
    cp.prbs = c->d_prbs;
    cp.punct = c->d_punct;
    cp.n_streams = c->n_streams;
    cp.row_bytes = c->row_bytes;
    cp.ring_rows = c->ring_rows;
    cp.ring_base = c->ring_base;
    cp.n_frames = (int)n_frames;
    k_code<<<(unsigned)(n_frames * c->n_streams), CODE_THREADS, 0, s>>>(cp);
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
        mp.total_words = (long long)(n_frames / c->cif_count) * c->tf_bytes / 4;
        k_mux<<<(unsigned)((mp.total_words + 255) / 256), 256, 0, s>>>(mp);
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

} // extern "C"

namespace {
// The coder as the front of the modulator's sliced pipeline: ETI frames cross PCIe on the copy stream, the
// coding kernels run on the modulator's compute stream right before the symbol kernels of the same slice,
// the coded blocks never leave the device.  Caller holds the coder's lock.
dabmod::PipeFront eti_front(dabmod_b200_coder *c, const uint8_t *eti)
{
    dabmod::PipeFront f;
    const size_t cif = (size_t)c->cif_count;
    f.upload = [c, eti, cif](size_t t0, size_t nt, cudaStream_t s_in) {
        CK(cudaMemcpyAsync(c->d_eti + t0 * cif * ETI_FRAME, eti + t0 * cif * ETI_FRAME, nt * cif * ETI_FRAME,
                           cudaMemcpyHostToDevice, s_in));
    };
    f.encode = [c, cif](size_t t0, size_t nt, cudaStream_t s) -> const uint8_t * {
        uint8_t *blocks = c->d_bits + t0 * (size_t)c->tf_bytes;
        coder_enqueue(c, c->d_eti + t0 * cif * ETI_FRAME, nt * cif, blocks, s);
        return blocks;
    };
    return f;
}

void check_chain(dabmod_b200 *h, dabmod_b200_coder *c, size_t n_frames)
{
    if (n_frames > (size_t)c->max_frames) throw CoderError(DABMOD_B200_EINVAL, "n_frames exceeds max_frames of the coder");
    if (n_frames % c->cif_count) throw CoderError(DABMOD_B200_ESTATE, "a call must carry whole transmission frames");
    if ((size_t)c->tf_bytes != dabmod_b200_tf_in_bytes(h))
        throw CoderError(DABMOD_B200_EINVAL, "coder and modulator are configured for different transmission modes");
    if (c->device != dabmod::device_of(h)) throw CoderError(DABMOD_B200_EINVAL, "coder and modulator are on different devices");
}
} // namespace

extern "C" {

int dabmod_b200_process_eti_batch(dabmod_b200 *h, dabmod_b200_coder *c, const uint8_t *eti, size_t n_frames,
                                  void *iq_out, size_t cap, size_t *out_bytes)
{
    if (out_bytes) *out_bytes = 0;
    return guard([&] {
        if (!h || !c || (n_frames && (!eti || !iq_out))) throw CoderError(DABMOD_B200_EINVAL, "null argument");
        check_chain(h, c, n_frames);
        std::lock_guard<std::mutex> lock(c->mtx);
        dabmod::PipeSink sink;
        sink.host_out = iq_out;
        sink.cap = cap;
        dabmod::run_pipeline(h, n_frames / c->cif_count, eti_front(c, eti), sink, out_bytes);
    });
}

int dabmod_b200_process_eti_batch_to_fd(dabmod_b200 *h, dabmod_b200_coder *c, const uint8_t *eti, size_t n_frames,
                                        int fd, size_t *out_bytes)
{
    if (out_bytes) *out_bytes = 0;
    return guard([&] {
        if (!h || !c || (n_frames && !eti)) throw CoderError(DABMOD_B200_EINVAL, "null argument");
        check_chain(h, c, n_frames);
        std::lock_guard<std::mutex> lock(c->mtx);
        dabmod::PipeSink sink;
        sink.to_fd = true;
        sink.fd = fd;
        dabmod::run_pipeline(h, n_frames / c->cif_count, eti_front(c, eti), sink, out_bytes);
    });
}

int dabmod_b200_seek_eti(dabmod_b200 *h, dabmod_b200_coder *c, uint64_t tf_index, const uint8_t *eti_before,
                         size_t n_before)
{
    return guard([&] {
        if (!h || !c || (n_before && !eti_before)) throw CoderError(DABMOD_B200_EINVAL, "null argument");
        check_chain(h, c, 0);
        const size_t cif = (size_t)c->cif_count;
        const size_t want = (size_t)std::min<uint64_t>(tf_index * cif, (uint64_t)(TI_DEPTH - 1) + cif);
        if (n_before < want)
            throw CoderError(DABMOD_B200_EINVAL, "seek_eti: " + std::to_string(want) + " ETI frames before the shard are "
                                                 "needed (time interleaver depth + one transmission frame), got " +
                                                 std::to_string(n_before));
        std::lock_guard<std::mutex> lock(c->mtx);
        CK(cudaSetDevice(c->device));
        CK(cudaMemsetAsync(c->d_punct, 0, (size_t)c->ring_rows * c->row_bytes, c->stream));
        c->ring_base = 0;
        if (tf_index == 0) {
            CK(cudaStreamSynchronize(c->stream));
            dabmod::seek_device(h, 0, nullptr);
            return;
        }
        // the frames that matter: 15 of time-interleaver history, then the transmission frame before the shard
        eti_before += (n_before - want) * ETI_FRAME;
        const size_t n_hist = want - cif;
        for (size_t done = 0; done < n_hist;) {
            const size_t n = std::min<size_t>(n_hist - done, (size_t)c->max_frames);
            CK(cudaMemcpyAsync(c->d_eti, eti_before + done * ETI_FRAME, n * ETI_FRAME, cudaMemcpyHostToDevice, c->stream));
            coder_enqueue(c, c->d_eti, n, nullptr, c->stream);
            CK(cudaStreamSynchronize(c->stream));       // d_eti is reused by the next chunk
            done += n;
        }
        CK(cudaMemcpyAsync(c->d_eti, eti_before + n_hist * ETI_FRAME, cif * ETI_FRAME, cudaMemcpyHostToDevice, c->stream));
        coder_enqueue(c, c->d_eti, cif, c->d_bits, c->stream);
        CK(cudaStreamSynchronize(c->stream));
        dabmod::seek_device(h, tf_index, c->d_bits);    // re-runs that frame up to the resampler input
    });
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
