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
