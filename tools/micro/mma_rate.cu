// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 operands, M = 128, K = 16 per instruction) for the operand layouts and
// N the persistent kernels use.  One CTA; operands are whatever shared memory holds (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I online-neural-cdes_b200/csrc -I include -o /tmp/mma_rate tools/micro/mma_rate.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include "ncde_b200.h"
#include "common.cuh"
#include "field_tc.cuh"
using namespace ncde;

// MODE 0: one thread of warp 0 issues everything (descriptors built once, advanced in an unrolled loop: the form the kernels use)
// MODE 1: two threads (warps 0 and 1) issue half each into different TMEM columns
// MODE 2: warp 0 converged, every MMA under elect.sync
__device__ __forceinline__ bool elect1() {
    uint32_t p;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(p));
    return p != 0;
}
template <int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int a_mn, int b_mn, int N, int n_mma, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ long long tstart[2], tend[2];
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int n_issuers = MODE == 1 ? 2 : 1;
    const bool issuer = MODE == 2 ? warp == 0 : (lane == 0 && warp < n_issuers);
    // MODE 3: the A tile is rewritten with generic-proxy stores before every timed batch
    long long best = 1ll << 60;
    for (int r = 0; r < reps; ++r) {
        if (MODE == 3) {
            for (int i = tid; i < 64 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u + r, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
            fence_async_smem();
        }
        __syncthreads();
        if ((MODE == 4 || MODE == 5 || MODE == 6) && warp >= 1) {
            // background shared-memory traffic from three warps into a scratch region above the operand tiles, for roughly as long as the batch lasts
            const uint32_t scratch = smem_u32(smem + 140 * 1024) + (uint32_t)(tid - 32) * 16u;
            uint32_t acc = 0;
            const int iters = MODE == 6 ? n_mma / 4 : n_mma * 4;
            for (int it = 0; it < iters; ++it) {
                if (MODE == 4 || MODE == 6) asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(scratch + (uint32_t)(it & 7) * 1536u), "r"(acc + it));
                else { uint32_t v0, v1, v2, v3; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(scratch + (uint32_t)(it & 7) * 1536u)); acc += v0 ^ v3; }
                if (MODE == 6) __nanosleep(200);
            }
            if (acc == 0x12345678u) out[3] = acc;
        }
        if (issuer) {
            const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 64 * 1024);
            const uint32_t idesc = make_idesc(128, N, a_mn, b_mn);
            const uint64_t ad0 = a_mn ? make_sdesc(a_s, 128u * 128u, 1024) : make_sdesc(a_s, 16, 1024);
            const uint64_t bd0 = b_mn ? make_sdesc(b_s, 128u * 128u, 1024) : make_sdesc(b_s, 16, 1024);
            const uint32_t d = tmem_base + (uint32_t)warp * 256u;
            const long long t0 = clock64();
            for (int i = 0; i < n_mma / n_issuers; i += 8) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t ao = a_mn ? (uint32_t)ks * 2048u : (uint32_t)(ks >> 2) * 128u * 128u + (uint32_t)(ks & 3) * 32u;
                    const uint32_t bo = b_mn ? (uint32_t)ks * 2048u : (uint32_t)(ks >> 2) * (uint32_t)N * 128u + (uint32_t)(ks & 3) * 32u;
                    if (MODE != 2 || elect1()) umma_bf16(d, ad0 + (ao >> 4), bd0 + (bo >> 4), idesc, (i + ks) > 0 ? 1u : 0u);
                }
            }
            const long long t_issue = clock64();
            if (MODE != 2 || elect1()) umma_commit(bar + warp);
            while (!mbar_try_wait(bar + warp, (uint32_t)r & 1u)) {}
            const long long t1 = clock64();
            if (lane == 0) { tstart[warp] = t0; tend[warp] = t1; if (warp == 0 && r == reps - 1) out[1] = t_issue - t0; }
        }
        __syncthreads();
        if (tid == 0) {
            long long a0 = tstart[0], a1 = tend[0];
            if (n_issuers == 2) { a0 = min(a0, tstart[1]); a1 = max(a1, tend[1]); }
            if (a1 - a0 < best) best = a1 - a0;
        }
    }
    if (tid == 0) out[0] = best;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int MODE>
static void run(long long* d, const char* name) {
    cudaFuncSetAttribute(mma_rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int cfgs[][3] = {{0, 0, 112}, {0, 0, 208}, {0, 0, 256}, {0, 1, 128}, {1, 1, 112}, {1, 1, 64}, {1, 1, 48}, {1, 1, 208}};
    for (auto& c : cfgs) {
        const int n_mma = 96;
        mma_rate_kernel<MODE><<<1, 128, 200 * 1024>>>(c[0], c[1], c[2], n_mma, 20, d);
        long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%-10s A %s  B %s  N=%3d  n=%2d : %6lld clk, %.1f clk/MMA (issue loop %.1f; math floor %.0f)  %s\n", name, c[0] ? "MN" : "K ", c[1] ? "MN" : "K ",
               c[2], n_mma, h[0], (double)h[0] / n_mma, (double)h[1] / n_mma * (MODE == 1 ? 2 : 1), c[2] / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    run<0>(d, "one-thread");
    run<4>(d, "bg st.shared");
    run<5>(d, "bg ld.shared");
    run<6>(d, "bg sparse st");
    return 0;
}
