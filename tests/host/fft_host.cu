// Host build of the register butterflies (fft.cuh, fft_reg.cuh) for the CPU test.
#define DABMOD_FN __host__ __device__ inline
#include "../../odr-dabmod_b200/csrc/fft_reg.cuh"

using namespace dabmod;

template <int R>
static void run(float2 *v, int inv)
{
    float2 a[R];
    for (int i = 0; i < R; i++) a[i] = v[i];
    if (inv) {
        if (R == 4) fft4<true>(a[0], a[1], a[2], a[3]);
        if (R == 8) fft8<true>(a);
        if (R == 16) fft16<true>(a);
    }
    else {
        if (R == 4) fft4<false>(a[0], a[1], a[2], a[3]);
        if (R == 8) fft8<false>(a);
        if (R == 16) fft16<false>(a);
    }
    for (int i = 0; i < R; i++) v[i] = a[i];
}

extern "C" int fft_host(int n, int inv, float *io)
{
    float2 *v = reinterpret_cast<float2 *>(io);
    switch (n) {
        case 4: run<4>(v, inv); return 0;
        case 8: run<8>(v, inv); return 0;
        case 16: run<16>(v, inv); return 0;
        case 32: {
            float2 a[32];
            for (int i = 0; i < 32; i++) a[i] = v[i];
            if (inv) fft32<true>(a); else fft32<false>(a);
            for (int i = 0; i < 32; i++) v[i] = a[i];
            return 0;
        }
        case 64: {
            float2 a[64];
            for (int i = 0; i < 64; i++) a[i] = v[i];
            if (inv) fft64<true>(a); else fft64<false>(a);
            for (int i = 0; i < 64; i++) v[i] = a[i];
            return 0;
        }
        case 5: {
            float2 a[5];
            for (int i = 0; i < 5; i++) a[i] = v[i];
            if (inv) dft5<true>(a[0], a[1], a[2], a[3], a[4]); else dft5<false>(a[0], a[1], a[2], a[3], a[4]);
            for (int i = 0; i < 5; i++) v[i] = a[i];
            return 0;
        }
        case 20: {
            float2 a[20];
            for (int i = 0; i < 20; i++) a[i] = v[i];
            if (inv) fft20<true>(a); else fft20<false>(a);
            for (int i = 0; i < 20; i++) v[i] = a[i];
            return 0;
        }
        case -10: {  // fft10_lo: the first five outputs of a 10-point transform
            float2 a[10];
            for (int i = 0; i < 10; i++) a[i] = v[i];
            if (inv) fft10_lo<true>(a); else fft10_lo<false>(a);
            for (int i = 0; i < 5; i++) v[i] = a[i];
            return 0;
        }
        case -8: {   // fft8r, the by-reference variant
            float2 a[8];
            for (int i = 0; i < 8; i++) a[i] = v[i];
            if (inv) fft8r<true>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
            else fft8r<false>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
            for (int i = 0; i < 8; i++) v[i] = a[i];
            return 0;
        }
    }
    return -1;
}

// ---- k_resample_q's per-thread stages (resample_q.cuh), one hop emulated thread by thread ----
#define RQ_HOST_ONLY 1
#include "../../odr-dabmod_b200/csrc/resample_q.cuh"
#include <vector>
#include <cmath>

// F: 4096 spectrum bins (already scaled); tw_out: No roots e^{+j 2 pi i / No}; out: P*2000 samples
extern "C" int rq_hop_host(int P, const float *F_, const float *tw_out_, float *out_)
{
    const float2 *Fin = reinterpret_cast<const float2 *>(F_);
    const float2 *tw_out = reinterpret_cast<const float2 *>(tw_out_);
    float2 *out = reinterpret_cast<float2 *>(out_);
    const int no = P * RQ_Q;
    std::vector<float2> itw2(5 * 20), itw3(4 * 400), buf(RQ_BUF), fold(RQ_NFOLD);
    for (int i = 0; i < 5 * 20; i++) itw2[i] = tw_out[(((i % 20) << (i / 20)) % 400) * 10 * P];
    for (int i = 0; i < 4 * 400; i++) itw3[i] = tw_out[(((i % 400) << (i / 400)) % 4000) * P];
    for (int t = 0; t < RQ_NFOLD; t++) fold[t] = Fin[2048 + t];
    std::vector<float2> regs(256 * 20);
    for (int rho = 0; rho < P; rho++) {
        for (auto &b : buf) b = make_float2(NAN, NAN);   // an unwritten slot poisons the result
        RqPhase ph;
        ph.c = tw_out[256 * rho];
        ph.d = tw_out[(no - 160 * rho) % no];
        ph.e = tw_out[96 * rho];
        for (int t = 0; t < 256; t++) {
            float2 F[16];
            for (int r = 0; r < 16; r++) F[r] = Fin[t + 256 * r];
            const float2 fp7 = t >= 160 ? fold[t - 160] : make_float2(0.f, 0.f);
            if (rho == 0) rq_spread<true>(F, fp7, fold[RQ_NFOLD - 1], t, tw_out[0], ph, buf.data());
            else rq_spread<false>(F, fp7, fold[RQ_NFOLD - 1], t, tw_out[t * rho], ph, buf.data());
        }
        for (int t = 0; t < RQ_ACTIVE; t++) {
            float2 v[20];
            rq_pass1_load(buf.data(), t, v);
            for (int r = 0; r < 20; r++) regs[t * 20 + r] = v[r];
        }
        for (int t = 0; t < RQ_ACTIVE; t++) {
            float2 v[20];
            for (int r = 0; r < 20; r++) v[r] = regs[t * 20 + r];
            rq_pass1_store(buf.data(), t, v);
        }
        for (int t = 0; t < RQ_ACTIVE; t++) {
            float2 v[20];
            rq_pass2_load(buf.data(), itw2.data(), t, v);
            for (int r = 0; r < 20; r++) regs[t * 20 + r] = v[r];
        }
        for (int t = 0; t < RQ_ACTIVE; t++) {
            float2 v[20];
            for (int r = 0; r < 20; r++) v[r] = regs[t * 20 + r];
            rq_pass2_store(buf.data(), t, v);
        }
        for (int b = 0; b < 400; b++) {
            float2 x[10];
            rq_pass3(buf.data(), itw3.data(), b, x);
            for (int r = 0; r < 5; r++) buf[RQ_S2 * r + b] = x[r];   // in place, like the last phase
        }
        for (int m = 0; m < RQ_KEEP; m++) {
            out[(size_t)m * P + rho] = buf[rq_result_slot(m)];
        }
    }
    return 0;
}
