// Throughput of the special-function ops the tanh epilogues can be built from (per SM per clock), B200.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters) {
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = 0.001f * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
            if (OP == 3) { unsigned u = __float_as_uint(x[j]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); x[j] = __uint_as_float(u); }
            if (OP == 4) x[j] = fmaf(x[j], 1.0001f, 0.5f);
            if (OP == 5) { unsigned u = __float_as_uint(x[j]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u)); x[j] = __uint_as_float(u); }
        }
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, float* d) {
    const int iters = 4096, blocks = 148 * 4, threads = 512;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<blocks, threads>>>(d, 16);
    cudaEventRecord(a);
    k<OP><<<blocks, threads>>>(d, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * threads * iters * 8;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-22s %.1f Gop/s  = %.2f lane-ops / clk / SM (at %d MHz)\n", name, ops / ms * 1e-6, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 4 * 512 * 4);
    run<0>("tanh.approx.f32", d); run<1>("ex2.approx.f32", d); run<2>("rcp.approx.f32", d); run<3>("tanh.approx.f16x2 (x2)", d);
    run<5>("tanh.approx.bf16x2 (x2)", d); run<4>("fma.f32", d);
    return 0;
}
