/* TEST INFRASTRUCTURE -- not product code (see ref_harness.cpp).
 *
 * The channel coding ahead of the hot path, driven through the
 * UNMODIFIED reference: raw ETI(NI) frames go into the reference's own EtiReader,
 * and the graph below is wired exactly as DabModulator::process does
 * (src/DabModulator.cpp:131-150, 286-383):
 *
 *   FicSource -> PrbsGenerator -> ConvEncoder -> PuncturingEncoder ------------------+
 *   SubchannelSource_i -> PrbsGenerator -> ConvEncoder -> PuncturingEncoder          |
 *        -> TimeInterleaver --+                                                      +-> BlockPartitioner -> capture
 *   PrbsGenerator(864*8) -----+-> FrameMultiplexer ----------------------------------+
 *
 * The only node that is not reference code is Capture (stands in for QpskSymbolMapper).
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <stdexcept>
#include <vector>

#include "Buffer.h"
#include "ModPlugin.h"
#include "Flowgraph.h"
#include "EtiReader.h"
#include "FicSource.h"
#include "SubchannelSource.h"
#include "PrbsGenerator.h"
#include "ConvEncoder.h"
#include "PuncturingEncoder.h"
#include "TimeInterleaver.h"
#include "FrameMultiplexer.h"
#include "BlockPartitioner.h"

namespace {

class Capture : public ModOutput {
public:
    std::vector<uint8_t> data;
    bool fresh = false;
    int process(Buffer *dataIn) override
    {
        const uint8_t *p = reinterpret_cast<const uint8_t *>(dataIn->getData());
        data.assign(p, p + dataIn->getLength());
        fresh = true;
        return (int)dataIn->getLength();
    }
    const char *name() override { return "Capture"; }
};

struct CoderHarness {
    double tist_offset = 0.0;
    EtiReader reader;
    std::unique_ptr<Flowgraph> fg;
    std::shared_ptr<Capture> sink;
    std::string err;
    CoderHarness() : reader(tist_offset) {}

    void build()
    {
        const unsigned mode = reader.getMode() == 0 ? 4 : reader.getMode();   /* MID 0 = TM IV (Eti.h) */
        fg.reset(new Flowgraph(false));
        auto cifPrbs = std::make_shared<PrbsGenerator>(864 * 8, 0x110);
        auto cifMux = std::make_shared<FrameMultiplexer>(reader);
        auto cifPart = std::make_shared<BlockPartitioner>(mode);
        sink = std::make_shared<Capture>();
        fg->connect(cifPrbs, cifMux);

        std::shared_ptr<FicSource> fic(reader.getFic());
        const size_t ficSizeIn = fic->getFramesize();
        auto ficPrbs = std::make_shared<PrbsGenerator>(ficSizeIn, 0x110);
        auto ficConv = std::make_shared<ConvEncoder>(ficSizeIn);
        auto ficPunc = std::make_shared<PuncturingEncoder>();
        for (const auto &rule : fic->get_rules()) ficPunc->append_rule(rule);
        ficPunc->append_tail_rule(PuncturingRule(3, 0xcccccc));
        fg->connect(fic, ficPrbs);
        fg->connect(ficPrbs, ficConv);
        fg->connect(ficConv, ficPunc);
        fg->connect(ficPunc, cifPart);

        for (const auto &sub : reader.getSubchannels()) {
            const size_t sizeIn = sub->framesize();
            const size_t sizeOut = sub->framesizeCu() * 8;
            auto prbs = std::make_shared<PrbsGenerator>(sizeIn, 0x110);
            auto conv = std::make_shared<ConvEncoder>(sizeIn);
            auto punc = std::make_shared<PuncturingEncoder>(sub->framesizeCu());
            for (const auto &rule : sub->get_rules()) punc->append_rule(rule);
            punc->append_tail_rule(PuncturingRule(3, 0xcccccc));
            auto til = std::make_shared<TimeInterleaver>(sizeOut);
            fg->connect(sub, prbs);
            fg->connect(prbs, conv);
            fg->connect(conv, punc);
            fg->connect(punc, til);
            fg->connect(til, cifMux);
        }
        fg->connect(cifMux, cifPart);
        fg->connect(cifPart, sink);
    }
};

thread_local std::string g_err;

} // namespace

extern "C" {

void *refc_create(void)
{
    try {
        return new CoderHarness();
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void refc_destroy(void *h) { delete static_cast<CoderHarness *>(h); }

const char *refc_last_error(void) { return g_err.c_str(); }

/* One raw ETI(NI) frame in.  Returns the number of BlockPartitioner bytes written to `out`
 * (0 while the transmission frame is incomplete), or -1 on error. */
long refc_feed(void *hv, const uint8_t *frame, size_t len, uint8_t *out, size_t cap)
{
    CoderHarness *h = static_cast<CoderHarness *>(hv);
    try {
        Buffer in(len, frame);
        size_t used = 0;
        while (used < len) {
            Buffer rest(len - used, frame + used);
            const int n = h->reader.loadEtiData(rest);
            if (n <= 0) break;
            used += (size_t)n;
        }
        if (!h->fg) h->build();
        h->sink->fresh = false;
        h->fg->run();
        if (!h->sink->fresh) return 0;
        if (h->sink->data.size() > cap) throw std::runtime_error("output buffer too small");
        std::memcpy(out, h->sink->data.data(), h->sink->data.size());
        return (long)h->sink->data.size();
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

/* What the reference derived from the ETI headers: for stream 0 = FIC and 1.. = subchannels:
 * framesize (input bytes), out_bytes, start address (CU), number of rules; rules as
 * (length, pattern) pairs into `rules` (cap pairs).  Returns the number of streams. */
int refc_describe(void *hv, uint32_t *framesize, uint32_t *out_bytes, uint32_t *start, uint32_t *n_rules,
                  uint32_t *rules, int cap_streams, int cap_rules)
{
    CoderHarness *h = static_cast<CoderHarness *>(hv);
    try {
        int s = 0, r = 0;
        auto put = [&](size_t fs, size_t ob, size_t st, const std::vector<PuncturingRule> &rl) {
            if (s >= cap_streams) throw std::runtime_error("too many streams");
            framesize[s] = (uint32_t)fs; out_bytes[s] = (uint32_t)ob; start[s] = (uint32_t)st;
            n_rules[s] = (uint32_t)rl.size();
            for (const auto &x : rl) {
                if (r >= cap_rules) throw std::runtime_error("too many rules");
                rules[2 * r] = (uint32_t)x.length(); rules[2 * r + 1] = x.pattern(); r++;
            }
            s++;
        };
        std::shared_ptr<FicSource> fic(h->reader.getFic());
        PuncturingEncoder pe;
        for (const auto &rule : fic->get_rules()) pe.append_rule(rule);
        pe.append_tail_rule(PuncturingRule(3, 0xcccccc));
        put(fic->getFramesize(), pe.getOutputSize(), 0, fic->get_rules());
        for (const auto &sub : h->reader.getSubchannels())
            put(sub->framesize(), sub->framesizeCu() * 8, sub->startAddress(), sub->get_rules());
        return s;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

} // extern "C"
