/* TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the channel coding that
 * feeds the COFDM hot path: ETI(NI) frame ->
 * energy dispersal -> convolutional encoder -> puncturing -> time interleaver ->
 * CIF assembly -> transmission-frame block (what QpskSymbolMapper receives).
 * Not product code: only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * may load it.
 *
 * Parity status: PINNED (bit-exact) against the unmodified reference blocks
 * driven by oracle/ref_coder_harness.cpp (tests/test_coder.py).
 *
 * Written from the behaviour of the reference files cited at each function
 * (paths relative to /root/reference/src).  Plain sequential bit loops, one
 * reference block per function, exactly like the reference's Flowgraph.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DABC_EXPORT __attribute__((visibility("default")))
#define DABC_MAX_STREAMS 65      /* FIC + 64 subchannels (NST is 7 bits, 64 is the ETSI limit) */
#define DABC_MAX_RULES 8
#define DABC_CIF_BYTES (864 * 8)
#define DABC_ETI_FRAME 6144

typedef struct { uint32_t length, pattern; } dabc_rule;

typedef struct {
    uint32_t framesize;       /* input bytes per ETI frame */
    uint32_t out_bytes;       /* FIC: punctured size; subchannel: framesizeCu * 8 */
    uint32_t start_cu;        /* start address in capacity units (subchannels) */
    uint32_t n_rules;
    dabc_rule rules[DABC_MAX_RULES];
} dabc_stream;

/* ------------------------------------------------------------------------ */
/* PrbsGenerator(framesize, 0x110), PrbsGenerator.cpp:126-188: 9-bit register */
/* preset to all ones; new bit = reg[8] ^ reg[4]; bytes MSB first.            */
/* ------------------------------------------------------------------------ */
DABC_EXPORT void dabc_prbs(int n, uint8_t *out)
{
    unsigned reg = 0x1ff;
    for (int i = 0; i < n; i++) {
        unsigned byte = 0;
        for (int b = 0; b < 8; b++) {
            const unsigned nb = ((reg >> 8) ^ (reg >> 4)) & 1u;
            reg = ((reg << 1) | nb) & 0x1ffu;
            byte = (byte << 1) | nb;
        }
        out[i] = (uint8_t)byte;
    }
}

/* ------------------------------------------------------------------------ */
/* ConvEncoder::process, ConvEncoder.cpp:59-150: K = 7, rate 1/4, generators  */
/* 0x5b 0x79 0x65 0x5b on a register fed at bit 6; 6 zero tail bits.          */
/* in: n bytes, out: 4 n + 3 bytes.                                           */
/* ------------------------------------------------------------------------ */
static unsigned parity7(unsigned v)
{
    v ^= v >> 4; v ^= v >> 2; v ^= v >> 1;
    return v & 1u;
}

DABC_EXPORT void dabc_conv(const uint8_t *in, int n, uint8_t *out)
{
    static const unsigned poly[4] = {0x5b, 0x79, 0x65, 0x5b};
    unsigned mem = 0;
    long o = 0;           /* output bit index */
    memset(out, 0, (size_t)4 * n + 3);
    for (long i = 0; i < 8L * n + 6; i++) {
        const unsigned bit = i < 8L * n ? (in[i >> 3] >> (7 - (i & 7))) & 1u : 0u;
        mem = (mem >> 1) | (bit << 6);
        for (int g = 0; g < 4; g++, o++)
            out[o >> 3] |= (uint8_t)(parity7(mem & poly[g]) << (7 - (o & 7)));
    }
}

/* ------------------------------------------------------------------------ */
/* PuncturingEncoder::process, PuncturingEncoder.cpp:102-210.  Rules are      */
/* consumed in order and cycled, each covers `length` input bytes in groups   */
/* of 4 with a 32-bit keep mask; the last 3 input bytes use the 24-bit tail   */
/* mask; kept bits packed MSB first, zero padded to out_bytes.                */
/* Returns the number of kept bits.                                           */
/* ------------------------------------------------------------------------ */
DABC_EXPORT long dabc_puncture(const uint8_t *in, long n_in, const dabc_rule *rules, int n_rules,
                               uint32_t tail_len, uint32_t tail_pattern, uint8_t *out, long out_bytes)
{
    long ob = 0;
    long ic = 0;
    const long body = n_in - (long)tail_len;
    int r = 0;
    memset(out, 0, (size_t)out_bytes);
    while (ic < body) {
        for (long len = rules[r].length; len > 0; len -= 4) {
            uint32_t mask = 0x80000000u;
            for (int i = 0; i < 4; i++) {
                const uint8_t d = in[ic++];
                for (int j = 0; j < 8; j++, mask >>= 1)
                    if (rules[r].pattern & mask) {
                        if (ob < 8 * out_bytes) out[ob >> 3] |= (uint8_t)(((d >> (7 - j)) & 1u) << (7 - (ob & 7)));
                        ob++;
                    }
            }
        }
        if (++r == n_rules) r = 0;
    }
    {
        uint32_t mask = 0x800000u;
        for (uint32_t i = 0; i < tail_len; i++) {
            const uint8_t d = in[ic++];
            for (int j = 0; j < 8; j++, mask >>= 1)
                if (tail_pattern & mask) {
                    if (ob < 8 * out_bytes) out[ob >> 3] |= (uint8_t)(((d >> (7 - j)) & 1u) << (7 - (ob & 7)));
                    ob++;
                }
        }
    }
    return ob;
}

/* ------------------------------------------------------------------------ */
/* Stream description from the ETI header: EtiReader.cpp:93-284 (FC, STC),    */
/* FicSource.cpp:40-63, SubchannelSource.cpp:66-170 (EEP rules), :672-760     */
/* (capacity units).  UEP (short form) subchannels are not derived here:      */
/* returns -2 for them (the tests pass the reference's rules explicitly).     */
/* Returns the number of streams (1 + NST), mode in *mode (MID, 0 -> 4).      */
/* ------------------------------------------------------------------------ */
static const uint32_t PI_MASK[25] = {0,
    0xc8888888, 0xc888c888, 0xc8c8c888, 0xc8c8c8c8, 0xccc8c8c8, 0xccc8ccc8, 0xccccccc8, 0xcccccccc,
    0xeccccccc, 0xeccceccc, 0xecececcc, 0xecececec, 0xeeececec, 0xeeeceeec, 0xeeeeeeec, 0xeeeeeeee,
    0xfeeeeeee, 0xfeeefeee, 0xfefefeee, 0xfefefefe, 0xfffefefe, 0xfffefffe, 0xfffffffe, 0xffffffff};

static long punct_out_bytes(const dabc_stream *s)
{
    /* PuncturingEncoder::adjust_item_size, PuncturingEncoder.cpp:60-78 */
    long bits = 0;
    for (uint32_t r = 0; r < s->n_rules; r++)
        bits += (long)(s->rules[r].length / 4) * __builtin_popcount(s->rules[r].pattern);
    bits += __builtin_popcount(0xcccccc);
    return (bits + 7) / 8;
}

DABC_EXPORT int dabc_describe(const uint8_t *frame, int *mode, dabc_stream *st, int cap)
{
    const unsigned nst = frame[5] & 0x7f, ficf = frame[5] >> 7, mid = (frame[6] >> 3) & 3;
    if (!ficf) return -1;
    if ((int)nst + 1 > cap) return -1;
    *mode = mid == 0 ? 4 : (int)mid;
    memset(st, 0, sizeof(dabc_stream) * (nst + 1));
    st[0].framesize = mid == 3 ? 128 : 96;
    st[0].n_rules = 2;
    st[0].rules[0] = (dabc_rule){(mid == 3 ? 29u : 21u) * 16u, 0xeeeeeeeeu};
    st[0].rules[1] = (dabc_rule){3u * 16u, 0xeeeeeeecu};
    st[0].out_bytes = (uint32_t)punct_out_bytes(&st[0]);
    for (unsigned i = 0; i < nst; i++) {
        const uint8_t *c = frame + 8 + 4 * i;
        const unsigned sad = ((c[0] & 3u) << 8) | c[1];
        const unsigned tpl = c[2] >> 2;
        const unsigned stl = ((c[2] & 3u) << 8) | c[3];
        dabc_stream *s = &st[1 + i];
        s->framesize = stl * 8;
        s->start_cu = sad;
        const unsigned br = s->framesize / 3;          /* kbit/s */
        if (!((tpl >> 5) & 1)) return -2;              /* short form = UEP tables */
        const unsigned opt = (tpl >> 2) & 7, lvl = (tpl & 3) + 1;
        s->n_rules = 2;
        if (opt == 0) {                                /* EEP-A */
            static const unsigned cu_per_8k[4] = {12, 8, 6, 4};
            s->out_bytes = (br / 8) * cu_per_8k[lvl - 1] * 8;
            switch (lvl) {
                case 1: s->rules[0] = (dabc_rule){((6 * br / 8) - 3) * 16, PI_MASK[24]};
                        s->rules[1] = (dabc_rule){3 * 16, PI_MASK[23]}; break;
                case 2:
                    if (br == 8) {
                        s->rules[0] = (dabc_rule){5 * 16, PI_MASK[13]};
                        s->rules[1] = (dabc_rule){1 * 16, PI_MASK[12]};
                    }
                    else {
                        s->rules[0] = (dabc_rule){((2 * br / 8) - 3) * 16, PI_MASK[14]};
                        s->rules[1] = (dabc_rule){((4 * br / 8) + 3) * 16, PI_MASK[13]};
                    }
                    break;
                case 3: s->rules[0] = (dabc_rule){((6 * br / 8) - 3) * 16, PI_MASK[8]};
                        s->rules[1] = (dabc_rule){3 * 16, PI_MASK[7]}; break;
                default: s->rules[0] = (dabc_rule){((4 * br / 8) - 3) * 16, PI_MASK[3]};
                         s->rules[1] = (dabc_rule){((2 * br / 8) + 3) * 16, PI_MASK[2]}; break;
            }
        }
        else if (opt == 1) {                           /* EEP-B */
            static const unsigned cu_per_32k[4] = {27, 21, 18, 15};
            static const int pi_a[4] = {10, 6, 4, 2}, pi_b[4] = {9, 5, 3, 1};
            s->out_bytes = (br / 32) * cu_per_32k[lvl - 1] * 8;
            s->rules[0] = (dabc_rule){((24 * br / 32) - 3) * 16, PI_MASK[pi_a[lvl - 1]]};
            s->rules[1] = (dabc_rule){3 * 16, PI_MASK[pi_b[lvl - 1]]};
        }
        else return -2;
    }
    return (int)nst + 1;
}

/* ------------------------------------------------------------------------ */
/* The chain of DabModulator.cpp:131-150,286-383 for one multiplex            */
/* configuration.  TimeInterleaver.cpp:51-96 (16-frame history per            */
/* subchannel), FrameMultiplexer.cpp:43-91 (PRBS filler + subchannels at      */
/* startAddress*8), BlockPartitioner.cpp:78-124 (FIC parts then CIFs).        */
/* ------------------------------------------------------------------------ */
typedef struct {
    int mode, n_streams, cif_count, fic_out, tf_bytes;
    dabc_stream st[DABC_MAX_STREAMS];
    uint8_t *hist[DABC_MAX_STREAMS];   /* 16 x out_bytes ring per subchannel, [slot][byte] */
    int hist_pos;                      /* slot of the newest frame */
    int cif_nb;                        /* position inside the transmission frame */
    uint8_t filler[DABC_CIF_BYTES];
    uint8_t *tf;
} dabc_coder;

DABC_EXPORT void dabc_coder_free(dabc_coder *c)
{
    if (!c) return;
    for (int i = 0; i < DABC_MAX_STREAMS; i++) free(c->hist[i]);
    free(c->tf);
    free(c);
}

DABC_EXPORT dabc_coder *dabc_coder_new(int mode, int n_streams, const dabc_stream *st)
{
    if (n_streams < 1 || n_streams > DABC_MAX_STREAMS || mode < 1 || mode > 4) return NULL;
    dabc_coder *c = calloc(1, sizeof(*c));
    c->mode = mode;
    c->n_streams = n_streams;
    memcpy(c->st, st, sizeof(dabc_stream) * n_streams);
    c->cif_count = mode == 1 ? 4 : mode == 4 ? 2 : 1;
    c->fic_out = mode == 3 ? 384 : 288;
    c->tf_bytes = c->cif_count * (c->fic_out + DABC_CIF_BYTES);
    if ((int)st[0].out_bytes != c->fic_out) { dabc_coder_free(c); return NULL; }
    for (int i = 1; i < n_streams; i++) {
        if ((st[i].out_bytes & 1) || (st[i].start_cu * 8 + st[i].out_bytes > DABC_CIF_BYTES)) { dabc_coder_free(c); return NULL; }
        c->hist[i] = calloc(16, st[i].out_bytes ? st[i].out_bytes : 1);
    }
    dabc_prbs(DABC_CIF_BYTES, c->filler);
    c->tf = calloc(1, c->tf_bytes);
    return c;
}

DABC_EXPORT int dabc_coder_tf_bytes(const dabc_coder *c) { return c->tf_bytes; }

/* One raw ETI(NI) frame (6144 bytes).  Returns tf_bytes and fills `out` when this frame
 * completes a transmission frame, else 0. */
DABC_EXPORT int dabc_coder_feed(dabc_coder *c, const uint8_t *frame, uint8_t *out)
{
    static const int delay_even[8] = {0, 8, 4, 12, 2, 10, 6, 14};   /* bit 7 .. bit 0 */
    const int nst = c->n_streams - 1;
    const uint8_t *data = frame + 8 + 4 * nst + 4;       /* SYNC, FC, STC[nst], EOH */
    uint8_t *fic_dst = c->tf + (size_t)c->cif_nb * c->fic_out;
    uint8_t *cif = c->tf + (size_t)c->cif_count * c->fic_out + (size_t)c->cif_nb * DABC_CIF_BYTES;
    memcpy(cif, c->filler, DABC_CIF_BYTES);
    c->hist_pos = (c->hist_pos + 15) & 15;               /* new newest slot (push_front) */
    for (int s = 0; s < c->n_streams; s++) {
        const dabc_stream *st = &c->st[s];
        const int n = (int)st->framesize;
        uint8_t *scr = malloc((size_t)n + 1), *prbs = malloc((size_t)n + 1), *enc = malloc((size_t)4 * n + 3);
        dabc_prbs(n, prbs);
        for (int i = 0; i < n; i++) scr[i] = data[i] ^ prbs[i];
        dabc_conv(scr, n, enc);
        if (s == 0) {
            dabc_puncture(enc, 4L * n + 3, st->rules, (int)st->n_rules, 3, 0xcccccc, fic_dst, c->fic_out);
        }
        else {
            const int ob = (int)st->out_bytes;
            uint8_t *newest = c->hist[s] + (size_t)c->hist_pos * ob;
            dabc_puncture(enc, 4L * n + 3, st->rules, (int)st->n_rules, 3, 0xcccccc, newest, ob);
            uint8_t *dst = cif + (size_t)st->start_cu * 8;
            for (int j = 0; j < ob; j++) {
                unsigned v = 0;
                for (int b = 0; b < 8; b++) {
                    const int d = delay_even[b] + (j & 1);
                    const uint8_t *h = c->hist[s] + (size_t)((c->hist_pos + d) & 15) * ob;
                    v |= h[j] & (0x80u >> b);
                }
                dst[j] = (uint8_t)v;
            }
        }
        free(scr); free(prbs); free(enc);
        data += n;
    }
    if (++c->cif_nb == c->cif_count) {
        c->cif_nb = 0;
        memcpy(out, c->tf, c->tf_bytes);
        return c->tf_bytes;
    }
    return 0;
}
