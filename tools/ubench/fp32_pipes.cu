// Microbenchmark: issue throughput of scalar vs packed FP32 ops on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { float2 v = make_float2(a, b); return *reinterpret_cast<u64 *>(&v); }

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float s, int iters, const float4 cst)
{
    float a[16];
    u64 p[8];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = pk(a[2 * i], a[2 * i + 1]);
    const u64 ps = pk(s, s * 0.5f), pc = pk(0.25f, 0.75f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) a[i] = fmaf(a[i], s, 0.5f * s);              // FFMA reg,reg,reg
            if (MODE == 1) a[i] = fmaf(a[i], cst.x, a[(i + 1) & 15]);   // FFMA reg,const,reg
            if (MODE == 2) a[i] = a[i] + s;                             // FADD
            if (MODE == 5) a[i] = fmaf(a[i], 1.0009765625f, 0.5f);      // FFMA imm
            if (MODE == 8) { if (i & 1) a[i] = fmaf(a[i], s, 0.5f * s); else a[i] = fminf(a[i], s) ; }  // FFMA + FMNMX
            if (MODE == 9) { if (i & 1) a[i] = fmaf(a[i], s, 0.5f * s); else a[i] = __int_as_float(__float_as_int(a[i]) ^ (i + it)); }  // FFMA + LOP3
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps), "l"(pc));
            if (MODE == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
            if (MODE == 6) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
            if (MODE == 7) {   // alternate FFMA2 / FADD2 on independent chains
                if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps), "l"(pc));
                else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) r += a[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { float2 v = *reinterpret_cast<float2 *>(&p[i]); r += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, int flops_per_op, int ops_per_iter)
{
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float *out;
    const int blocks = sms * 8, iters = 20000;
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const float4 c = make_float4(1.0001f, 0, 0, 0);
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 100, c);
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(out, 1.0001f, iters, c);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)blocks * 256 * iters * ops_per_iter;
    const double per_clk_sm = ops / (ms * 1e-3) / sms / (khz * 1e3);
    printf("%-22s %8.3f ms  %7.1f thread-ops/clk/SM (at %d MHz nominal)  %7.2f TFLOP/s\n", name, ms, per_clk_sm,
           khz / 1000, ops * flops_per_op / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main()
{
    run<0>("FFMA reg", 2, 16);
    run<1>("FFMA const-operand", 2, 16);
    run<5>("FFMA imm", 2, 16);
    run<2>("FADD reg", 1, 16);
    run<3>("FFMA2 (f32x2)", 4, 8);
    run<4>("FADD2 (f32x2)", 2, 8);
    run<6>("FMUL2 (f32x2)", 2, 8);
    run<7>("FFMA2+FADD2 alternating", 3, 8);
    run<8>("FFMA+FMNMX alternating", 1, 16);
    run<9>("FFMA+LOP3 alternating", 1, 16);
    return 0;
}
