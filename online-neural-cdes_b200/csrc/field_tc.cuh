// Final-layer kernels on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a.
//
// bf16 operands, fp32 accumulate.  Operand tiles live in shared memory in the canonical 128-byte-swizzled layout:
// a tile of R rows x KP bf16 columns is split into 64-column blocks, each block is R rows of 128 bytes, and inside
// every 8-row group (1024 B) the 16-byte chunk c of row r is stored at chunk position c ^ (r & 7).  The same bytes
// serve as a K-major operand (rows = M/N index, K contiguous) and as an MN-major operand (rows = K index, M/N
// contiguous), which is what lets the backward kernel keep ONE copy of the weights, the activations and G.
//
// Tiles are filled with ordinary 16-byte stores (the operands are produced or converted in the kernel anyway),
// followed by fence.proxy.async so the tensor-core (async proxy) reads see them.
#pragma once
#include <cuda.h>        // CUtensorMap (types only; the encoder is resolved at run time, nothing links against libcuda)
#include <cuda_bf16.h>

#include "solve_kernels.cuh"

namespace ncde {

constexpr int kTcEpiThreads = 256;             // 8 epilogue warps: thread = TMEM lane (batch row), two warp-groups split the columns
constexpr int kTcThreads = kTcEpiThreads + 32; // + one producer warp: TMA loads and every tcgen05.mma are issued by its lane 0
constexpr int kTcM = 128;  // batch rows per MMA tile (= TMEM lanes)
constexpr int kTcKP = 128; // K of the final layer padded to two 64-column swizzle blocks

struct TcFieldArgs {
    int B, Bp, H, Cp, Hg, n_hg, Npad, KP, DF, Bt;
    int CpB;                   // floats per dX/dt row in shared memory (Cp padded so that CpB/4 is odd: conflict-free LDS.128)
    int rec_a, rec_x;          // record (RK stage) index inside the activation / dX tensor maps
    const __nv_bfloat16* Wbf;  // [n_hg][Npad][KP]
    const float* b3;           // [n_hg][Npad]
    const __nv_bfloat16* abf;  // [Bp][KP] final-layer input, bf16 row-major            (read through TMA map A)
    const float* dXT;          // [Bp][Cp] dX/dt, ROW-major on the tensor-core path      (read through TMA map X)
    float* koutT;              // [H][Bp]            forward
    const float* gkT;          // [H][Bp]            backward
    float* P;                  // [n_hg][B][DFP]     backward
    float* dW3acc;             // [n_bt][Np][DFP]
    float* db3acc;             // [n_bt][Np]
    int DFP;
    const AdaptCtrl* ctrl;     // adaptive solver: skip all work once ctrl->done
    // all-tensor-core fixed-grid path: the forward epilogue also emits the NEXT stage's input y + increment(k...) as a bf16
    // row-major record [Bp][128] (the operand tile of the tensor-core hidden layers)
    __nv_bfloat16* zs_out;     // null: nothing to emit
    const float* yT;           // [H][Bp]
    const float* kT[NCDE_MAX_STAGES];   // stage derivatives incl. the one this launch writes (koutT)
    int next_combine;
    float dt;
    // last stage of a step (all-tensor-core path): the epilogue also advances the state, y_{n+1} = y_n + RK combination
    // (fixed_grid.py:6-29, rk_common.py:114; same operation order as advance_kernel), and emits it
    float* adv_ynewT;          // [H][Bp] or null: nothing to advance
    __nv_bfloat16* adv_ybf;    // bf16 record [Bp][128] of y_{n+1} (input of the next step's first stage) or null
    float* adv_emit;           // (B, H) row-major slice of z_out that receives y_{n+1}, or null
    int adv_method;
    // backward of the all-tensor-core path: dL/dk of this stage is formed on the fly from the gradient of the step result and
    // the stage-input gradients of the LATER stages (rk_common.py:106-114 transposed), so no gk arrays are read-modify-written:
    //     gk[b,h] = gcoef * gy1[b,h] + sum_q dzcoef[q] * dz_q[b,h]
    int prefetch;              // fixed-grid solves: the saved records this launch reads (dX/dt; in the backward pass also the
                               // activations) are older than the predecessor kernel, so their first tiles — and in the backward
                               // kernel the first recompute MMA — are issued ahead of griddepcontrol.wait, under the predecessor
    int gk_early;              // form gk (global loads of gy / dz) BEFORE waiting for the accumulators instead of after (NCDE_GK_EARLY=1;
                               // experimental: 17 % of the backward kernel's stall samples sit on those loads, profiles/README.md)
    int p_transposed;          // P is written as bf16 P^T[g][k][Bp] (coalesced; consumed by p_reduce) instead of fp32 [g][b][k]
    const float* gy1T;         // [H][Bp] or null (then gkT is read)
    float gcoef;
    int n_dz;
    const float* dzT[NCDE_MAX_STAGES];
    float dzcoef[NCDE_MAX_STAGES];
};

// TMA descriptors of one launch: W = packed final-layer weights {k, n} box {64, Npad}; A = bf16 activations {k, b, rec}
// box {64, 128, 1}; X = dX/dt {c, b, rec} box {CpB, 128, 1}.  A and W use the 128-byte swizzle (the UMMA canonical layout).
struct TcMaps { CUtensorMap W, A, X; };

__host__ __device__ inline uint32_t tc_tmem_cols(int need) {
    return need <= 32 ? 32u : (need <= 64 ? 64u : (need <= 128 ? 128u : (need <= 256 ? 256u : 512u)));
}

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA tile loads (cp.async.bulk.tensor): completion is signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// tcgen05.commit: the mbarrier receives one arrival when every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// issue-only variants: the registers are valid after tmem_wait_ld(...) on the same array
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
// tcgen05.wait::ld, tied to the destination registers so that no use of them can be scheduled above the wait
template <int W>
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[W]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < W; i += 8)
        asm volatile("" : "+r"(r[i]), "+r"(r[i + 1]), "+r"(r[i + 2]), "+r"(r[i + 3]), "+r"(r[i + 4]), "+r"(r[i + 5]), "+r"(r[i + 6]),
                          "+r"(r[i + 7]));
}
template <int W>
__device__ __forceinline__ void tmem_ldw_issue(uint32_t taddr, uint32_t (&r)[W]);
template <>
__device__ __forceinline__ void tmem_ldw_issue<8>(uint32_t taddr, uint32_t (&r)[8]) { tmem_ld8_issue(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ldw_issue<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32_issue(taddr, r); }

template <int W>
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, float (&v)[W]);
template <>
__device__ __forceinline__ void tmem_ldw<8>(uint32_t taddr, float (&v)[8]) { tmem_ld8(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ldw<32>(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// tanh(x) and 1 - tanh(x)^2 from one ex2.approx and one rcp.approx (both ~1 ulp): with s = e^{-2|x|},
// tanh|x| = (1 - s) / (1 + s) and sech^2 = 4 s / (1 + s)^2 — no cancellation when |tanh| -> 1, unlike 1 - t*t on top
// of tanh.approx (2^-11 error).
__device__ __forceinline__ void tanh_sech2(float x, float& t, float& sech2) {
    const float ax = fabsf(x);
    float s, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(-2.885390081777927f * ax));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + s));
    t = copysignf((1.f - s) * r, x);
    sech2 = 4.f * s * r * r;
}
// The pre-activation of the bf16 path carries ~1e-2 absolute error from operand rounding, so MUFU.TANH (2^-11) is
// noise here and costs one SFU op instead of two plus ten FP ops.  NCDE_TC_EXACT_TANH=1 builds the ex2/rcp variant.
#ifndef NCDE_TC_EXACT_TANH
#define NCDE_TC_EXACT_TANH 0
#endif
__device__ __forceinline__ float tanh_fast(float x) {
#if NCDE_TC_EXACT_TANH
    float t, q;
    tanh_sech2(x, t, q);
    return t;
#else
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}
__device__ __forceinline__ float sech2_fast(float x) {
#if NCDE_TC_EXACT_TANH
    float t, q;
    tanh_sech2(x, t, q);
    return q;
#else
    const float t = tanh_fast(x);
    return fmaf(-t, t, 1.f);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// descriptors (cute/arch/mma_sm100_desc.hpp bit layouts)
// ---------------------------------------------------------------------------------------------------------------
// instruction descriptor, kind::f16: D = F32 (bit 4), A = B = BF16 (bits 7, 10), a_major bit 15, b_major bit 16,
// N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor, SWIZZLE_128B, version 1 (Blackwell).  Offsets in bytes.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
// byte offset of the 16-byte chunk `chunk16` (index along the contiguous dimension, 8 bf16 per chunk) of row `row`
// inside a swizzled tile with `rows` rows
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk16, int rows) {
    const int blk = chunk16 >> 3, c = chunk16 & 7;
    return (uint32_t)blk * (uint32_t)rows * 128u + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u +
           (uint32_t)((c ^ (row & 7)) << 4);
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing for the tensor-core path: bf16, per h-group [Npad][KP] row-major (K contiguous), zero padded
// ---------------------------------------------------------------------------------------------------------------
__global__ void pack_final_bf16_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                       __nv_bfloat16* __restrict__ Wbf, float* __restrict__ b3, int H, int C, int Cp,
                                       int Hg, int n_hg, int Npad, int KP, int DF) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)n_hg * Npad * KP;
    if (idx < total) {
        const int k = (int)(idx % KP);
        const int nl = (int)((idx / KP) % Npad);
        const int g = (int)(idx / ((int64_t)KP * Npad));
        const int hl = nl / Cp, c = nl % Cp, h = g * Hg + hl;
        float v = 0.f;
        if (hl < Hg && h < H && c < C && k < DF) v = W[((int64_t)h * C + c) * DF + k];
        Wbf[idx] = __float2bfloat16(v);
    }
    if (idx < (int64_t)n_hg * Npad) {
        const int nl = (int)(idx % Npad), g = (int)(idx / Npad);
        const int hl = nl / Cp, c = nl % Cp, h = g * Hg + hl;
        b3[idx] = (bias && hl < Hg && h < H && c < C) ? bias[(int64_t)h * C + c] : 0.f;
    }
}

// D[128 x N] (+)= A[128 x K] . B[N x K]^T, both operands K-major swizzled tiles; one thread issues
__device__ __forceinline__ void issue_gemm_kmajor(uint32_t d_tmem, uint32_t a_saddr, int a_rows, uint32_t b_saddr, int b_rows,
                                                  int N, int K, bool accumulate_first) {
    const uint32_t idesc = make_idesc(kTcM, N, 0, 0);
    for (int k = 0; k < K / 16; ++k) {
        const uint32_t koff_a = (uint32_t)(k >> 2) * (uint32_t)a_rows * 128u + (uint32_t)(k & 3) * 32u;
        const uint32_t koff_b = (uint32_t)(k >> 2) * (uint32_t)b_rows * 128u + (uint32_t)(k & 3) * 32u;
        umma_bf16(d_tmem, make_sdesc(a_saddr + koff_a, 16, 1024), make_sdesc(b_saddr + koff_b, 16, 1024), idesc,
                  (k > 0 || accumulate_first) ? 1u : 0u);
    }
}

// One W-column chunk of an epilogue for TMEM lane `row`: the accumulator values (tcgen05.ld in flight until
// tmem_wait_ld), the biases and this row's dX/dt (both LDS.128 from 16-byte aligned shared addresses).  Chunks are software
// pipelined: wait for chunk j, issue chunk j+1, then compute chunk j — tcgen05.wait::ld waits for every outstanding load of
// the thread, so the next load is issued right after the wait and flies during the arithmetic.
template <int W>
struct TcChunk {
    uint32_t r[W];
    uint32_t b3_s, dx_s;   // shared-memory addresses of the chunk's biases and dX/dt values (read right before use: only the
                           // TMEM registers are double-buffered, 2 x 32 of the 168 the launch bounds allow)
};
// chunk j (32 columns) of the unit that starts at column colbase + c_begin of this row; expects lane_addr, colbase, c_begin,
// b3_s and dx_s in scope (a macro, not a lambda: a lambda taking the chunk by reference pushes it to local memory)
#define issue32(k, j) chunk_issue<32>(k, lane_addr + (uint32_t)(colbase + c_begin + 32 * (j)), b3_s + 4u * (colbase + c_begin + 32 * (j)), \
                                      dx_s + 4u * (c_begin + 32 * (j)))
template <int W>
__device__ __forceinline__ void chunk_issue(TcChunk<W>& k, uint32_t taddr, uint32_t b3_s, uint32_t dx_s) {
    tmem_ldw_issue<W>(taddr, k.r);
    k.b3_s = b3_s; k.dx_s = dx_s;
}
// sum_j tanh(D[row][col0 + j] + b3[j]) * dX[row][j] over the chunk (after tmem_wait_ld)
template <int W>
__device__ __forceinline__ float fwd_finish(const TcChunk<W>& k) {
    float4 bb[W / 4], dd[W / 4];
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
        bb[q] = lds128(k.b3_s + 16u * q); dd[q] = lds128(k.dx_s + 16u * q);
    }
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
        acc0 = fmaf(tanh_fast(__uint_as_float(k.r[4 * q + 0]) + bb[q].x), dd[q].x, acc0);
        acc1 = fmaf(tanh_fast(__uint_as_float(k.r[4 * q + 1]) + bb[q].y), dd[q].y, acc1);
        acc2 = fmaf(tanh_fast(__uint_as_float(k.r[4 * q + 2]) + bb[q].z), dd[q].z, acc2);
        acc3 = fmaf(tanh_fast(__uint_as_float(k.r[4 * q + 3]) + bb[q].w), dd[q].w, acc3);
    }
    return (acc0 + acc1) + (acc2 + acc3);
}

// ---------------------------------------------------------------------------------------------------------------
// forward:  k[b, h] = sum_c tanh( a[b,:] . W3[(h,c),:] + b3[(h,c)] ) * dX[b, c]
// CTA (g, bt): the W3 slice of h-group g stays in shared memory; the 128-row tiles of the batch stream through a
// warp-specialised pipeline.  Warp 8 (one elected lane) is the producer: TMA loads of the bf16 activation tile and of the
// dX/dt tile (both double-buffered) and the tcgen05.mma of tile i+1 into the second TMEM accumulator, all while warps
// 0-7 run the epilogue of tile i (the CUDA-core-bound part: one MUFU.TANH per element).  Hand-offs are mbarriers only:
//   full_a[b] / full_x[b]  TMA -> MMA / epilogue        mma[b]  MMA complete -> epilogue, activation buffer free
//   done[b]                epilogue finished (8 warp arrivals) -> accumulator b and dX buffer b free
// Epilogue: thread = one TMEM lane (batch row); the two warp-groups split the columns.
// ---------------------------------------------------------------------------------------------------------------
// what the forward epilogue does with one finished k[b, h]: store it (feature-major, coalesced over the lanes = rows), emit the
// next stage's input record, or — after the last stage of a step — advance the state and emit it
__device__ __forceinline__ void tc_fwd_finish_element(const TcFieldArgs& a, int h, int64_t b, float k) {
    const size_t off = (size_t)h * a.Bp + b;
    a.koutT[off] = k;
    if (a.zs_out)   // this thread's store above is visible to its own loads
        a.zs_out[(size_t)b * 128 + h] = __float2bfloat16(__fadd_rn(a.yT[off], stage_increment(a.next_combine, a.dt, a.kT, nullptr, off)));
    if (a.adv_ynewT) {
        const float y = a.yT[off];
        float yn;
        if (a.adv_method == NCDE_RK4_38) {   // (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
            float s = __fadd_rn(a.kT[0][off], __fmul_rn(3.f, __fadd_rn(a.kT[1][off], a.kT[2][off])));
            s = __fadd_rn(s, k);
            yn = __fadd_rn(y, __fmul_rn(__fmul_rn(s, a.dt), 0.125f));
        } else {                             // Euler: y0 + dt * f0
            yn = __fadd_rn(y, __fmul_rn(a.dt, k));
        }
        a.adv_ynewT[off] = yn;
        if (a.adv_ybf) a.adv_ybf[(size_t)b * 128 + h] = __float2bfloat16(yn);
        if (a.adv_emit) a.adv_emit[(size_t)b * a.H + h] = yn;
    }
}

struct TcFwdSmem {   // byte offsets from the 1024-aligned base
    uint32_t Ws, As, dXs, b3s, part, bars, total;
};
__host__ __device__ inline TcFwdSmem tc_fwd_layout(int Npad, int CpB) {
    TcFwdSmem L;
    uint32_t o = 0;
    L.Ws = o; o += (uint32_t)Npad * kTcKP * 2;
    L.As = o; o += 2u * kTcM * kTcKP * 2;
    L.dXs = o; o += 2u * kTcM * (uint32_t)CpB * 4;
    L.b3s = o; o += (uint32_t)Npad * 4;
    L.part = o; o += kTcM * 4;
    o = (o + 15u) & ~15u;
    L.bars = o; o += 16 * 8;
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(kTcThreads, 1) tc_field_fwd_kernel(const __grid_constant__ TcFieldArgs a,
                                                                     const __grid_constant__ TcMaps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int Npad = a.Npad;
    const TcFwdSmem L = tc_fwd_layout(Npad, a.CpB);
    const uint32_t a_bytes = kTcM * kTcKP * 2, x_bytes = kTcM * (uint32_t)a.CpB * 4;
    uint8_t* Ws = smem + L.Ws;
    uint8_t* As = smem + L.As;
    uint8_t* dXs = smem + L.dXs;
    float* b3s = reinterpret_cast<float*>(smem + L.b3s);
    float* part = reinterpret_cast<float*>(smem + L.part);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* full_a = bars;          // [2]
    uint64_t* full_x = bars + 2;      // [2]
    uint64_t* mma_bar = bars + 4;     // [2]
    uint64_t* done = bars + 6;        // [2]
    uint64_t* w_bar = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x, bt = blockIdx.y;
    const uint32_t acc_stride = tc_tmem_cols(Npad);     // columns per accumulator (power of two >= 32)

    if (warp == 0) tmem_alloc(tmem_slot, 2 * acc_stride);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(full_a + i, 1); mbar_init(full_x + i, 1); mbar_init(mma_bar + i, 1); mbar_init(done + i, 8); }
        mbar_init(w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < Npad; i += kTcThreads) b3s[i] = a.b3[(size_t)g * Npad + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool producer = warp == 8 && lane == 0;
    if (producer) {
        // the packed weights were written long before the predecessor kernel started: safe to fetch ahead of the dependency
        tma_prefetch_desc(&maps.A);
        tma_prefetch_desc(&maps.X);
        mbar_expect_tx(w_bar, (uint32_t)Npad * kTcKP * 2);
        tma_load_3d(Ws, &maps.W, w_bar, 0, g * Npad, 0);
        tma_load_3d(Ws + (size_t)Npad * 128, &maps.W, w_bar, 64, g * Npad, 0);
    }
    const int64_t row_begin = (int64_t)bt * a.Bt;
    const bool pre = a.prefetch != 0 && a.ctrl == nullptr;
    if (producer && pre) {
        // dX/dt of the first two tiles: written by dx_all before the stage loop began
        for (int i = 0; i < 2; ++i) {
            const int64_t b0 = row_begin + (int64_t)i * kTcM;
            if (b0 < a.B && b0 < row_begin + a.Bt) {
                mbar_expect_tx(full_x + i, x_bytes);
                tma_load_3d(dXs + (size_t)i * x_bytes, &maps.X, full_x + i, 0, (int)b0, a.rec_x);
            }
        }
    }
    pdl_trigger();
    pdl_wait();  // the activation tiles come from hidden_fwd

    int64_t row_end = min((int64_t)a.B, row_begin + a.Bt);
    if (a.ctrl && a.ctrl->done) row_end = row_begin;
    const int nt = row_end > row_begin ? (int)((row_end - row_begin + kTcM - 1) / kTcM) : 0;

    if (warp == 8) {
        if (lane == 0 && nt > 0) {
            auto load = [&](int i) {
                const int b = i & 1;
                const int b0 = (int)(row_begin + (int64_t)i * kTcM);
                mbar_expect_tx(full_a + b, a_bytes);
                tma_load_3d(As + (size_t)b * a_bytes, &maps.A, full_a + b, 0, b0, a.rec_a);
                tma_load_3d(As + (size_t)b * a_bytes + a_bytes / 2, &maps.A, full_a + b, 64, b0, a.rec_a);
                if (pre && i < 2) return;   // already in flight
                mbar_expect_tx(full_x + b, x_bytes);
                tma_load_3d(dXs + (size_t)b * x_bytes, &maps.X, full_x + b, 0, b0, a.rec_x);
            };
            auto issue = [&](int i) {
                const int b = i & 1;
                mbar_wait(full_a + b, (uint32_t)(i >> 1) & 1u);
                tc_fence_after();
                issue_gemm_kmajor(tmem_base + (uint32_t)b * acc_stride, smem_u32(As + (size_t)b * a_bytes), kTcM, smem_u32(Ws), Npad, Npad,
                                  kTcKP, false);
                umma_commit(mma_bar + b);
            };
            load(0);
            if (nt > 1) load(1);
            mbar_wait(w_bar, 0);
            issue(0);
            for (int i = 0; i < nt; ++i) {
                if (i + 1 < nt) issue(i + 1);      // accumulator (i+1)&1 is free: load(i+1) was issued after done(i-1)
                if (i + 2 < nt) {
                    mbar_wait(mma_bar + (i & 1), (uint32_t)(i >> 1) & 1u);   // MMA(i) complete: activation buffer free
                    mbar_wait(done + (i & 1), (uint32_t)(i >> 1) & 1u);      // epilogue(i) complete: dX buffer and accumulator free
                    load(i + 2);
                }
            }
        }
    } else {
        const int wg = warp >> 2;                    // warp-group 0/1
        const int row = (warp & 3) * 32 + lane;      // TMEM lane == row inside the tile
        // share of this warp-group: a set of whole h's when Hg >= 2, else half of the channels of the single h
        int h_begin, h_end, c_begin, c_end;
        if (a.Hg >= 2) { h_begin = wg == 0 ? 0 : a.Hg / 2; h_end = wg == 0 ? a.Hg / 2 : a.Hg; c_begin = 0; c_end = a.Cp; }
        else { const int half = ((a.Cp / 2 + 7) / 8) * 8; h_begin = 0; h_end = 1; c_begin = wg == 0 ? 0 : half; c_end = wg == 0 ? half : a.Cp; }
        const uint32_t b3_s = smem_u32(b3s);
        for (int i = 0; i < nt; ++i) {
            const int b = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            const int64_t b0 = row_begin + (int64_t)i * kTcM;
            mbar_wait(full_x + b, ph);
            mbar_wait(mma_bar + b, ph);
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + (uint32_t)b * acc_stride + ((uint32_t)((warp & 3) * 32) << 16);
            const uint32_t dx_s = smem_u32(dXs) + (uint32_t)b * x_bytes + (uint32_t)row * (uint32_t)a.CpB * 4u;
            for (int hl = h_begin; hl < h_end; ++hl) {
                float acc = 0.f;
                const int colbase = hl * a.Cp;
                const int n32 = (c_end - c_begin) / 32;
                {
                    TcChunk<32> A, Bk;
                    if (n32 > 0) issue32(A, 0);
                    for (int j = 0; j < n32; j += 2) {
                        tmem_wait_ld<32>(A.r);
                        if (j + 1 < n32) issue32(Bk, j + 1);
                        acc += fwd_finish<32>(A);
                        if (j + 1 < n32) {
                            tmem_wait_ld<32>(Bk.r);
                            if (j + 2 < n32) issue32(A, j + 2);
                            acc += fwd_finish<32>(Bk);
                        }
                    }
                }
                for (int c0 = c_begin + 32 * n32; c0 + 8 <= c_end; c0 += 8) {
                    TcChunk<8> T;
                    chunk_issue<8>(T, lane_addr + (uint32_t)(colbase + c0), b3_s + 4u * (colbase + c0), dx_s + 4u * c0);
                    tmem_wait_ld<8>(T.r);
                    acc += fwd_finish<8>(T);
                }
                if (a.Hg >= 2) {
                    const int h = g * a.Hg + hl;
                    if (h < a.H && b0 + row < a.B) tc_fwd_finish_element(a, h, b0 + row, acc);
                } else {
                    if (wg == 0) part[row] = acc;
                    named_bar_sync(1, kTcEpiThreads);
                    if (wg == 1 && g < a.H && b0 + row < a.B) tc_fwd_finish_element(a, g, b0 + row, acc + part[row]);
                    named_bar_sync(1, kTcEpiThreads);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(done + b);
        }
    }
    tc_fence_before();
    __syncwarp();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 2 * acc_stride);
}

static inline size_t tc_fwd_smem_bytes(int Npad, int CpB) { return 1024 + tc_fwd_layout(Npad, CpB).total; }

// ---------------------------------------------------------------------------------------------------------------
// backward of one RK stage:
//   MMA1   pre[m][n]  = As . Ws^T                      (A, B K-major)                 -> TMEM cols [0, Npad)
//   epi 1  G = gk * dX * (1 - tanh^2(pre + b3)) -> bf16 tile Gs[m][n]
//   dgrad  P[m][k]    = Gs . Ws        (A = Gs K-major over n, B = Ws MN-major)       -> TMEM cols [0, KP)
//   wgrad  dW^T[k][n] += As^T . Gs     (A = As MN-major, B = Gs MN-major)             -> TMEM cols [256, 256+Npad)
//   bias   db3[n]     += sum_m Gs[m][n]   column sums of the bf16 tile by the CUDA cores WHILE dgrad/wgrad run
//   epi 2  P -> global partial of dL/d(act) for this h-group (waits for dgrad only; wgrad overlaps it)
// dW^T stays in TMEM across all row tiles of the CTA and is added to the global accumulator once at the end.
// Same warp specialisation as the forward kernel: warp 8 issues the TMA loads (the next dX/dt tile as soon as epilogue 1 is
// done, the next activation tile as soon as wgrad is: it lands during epilogue 2) and every MMA; warps 0-7 run the
// epilogues.  Barriers:
//   full_a, full_x      TMA landed          pre_bar  MMA1 complete        g_ready  G tile written (8 warp arrivals)
//   dg_bar  dgrad complete                  done2    epilogue 2 finished (8 warp arrivals)      fin_bar  last wgrad complete
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kTcDwCol = 256;

// One W-column chunk of epilogue 1 for TMEM lane `row` (after tmem_wait_ld): G = gk * dX * sech^2(pre + b3) -> bf16 into the
// swizzled G tile (16-byte stores, n0 is a multiple of 8).
template <int W>
__device__ __forceinline__ void bwd_finish(const TcChunk<W>& k, float gk, uint32_t gs_s, int row, int n0) {
    float v[W];
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
        const float4 bb = lds128(k.b3_s + 16u * q), dd = lds128(k.dx_s + 16u * q);
        // gk is zero for padded rows and dX/dt rows beyond the batch are zero-filled by TMA: no NaN can enter the G tile
        v[4 * q + 0] = gk * dd.x * sech2_fast(__uint_as_float(k.r[4 * q + 0]) + bb.x);
        v[4 * q + 1] = gk * dd.y * sech2_fast(__uint_as_float(k.r[4 * q + 1]) + bb.y);
        v[4 * q + 2] = gk * dd.z * sech2_fast(__uint_as_float(k.r[4 * q + 2]) + bb.z);
        v[4 * q + 3] = gk * dd.w * sech2_fast(__uint_as_float(k.r[4 * q + 3]) + bb.w);
    }
#pragma unroll
    for (int j8 = 0; j8 < W / 8; ++j8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j8 * 8 + 2 * j], v[j8 * 8 + 2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(gs_s + sw128_off(row, (n0 >> 3) + j8, kTcM)), "r"(pk[0]), "r"(pk[1]),
                     "r"(pk[2]), "r"(pk[3]) : "memory");
    }
}

struct TcBwdSmem {
    uint32_t Ws, As, Gs, dXs, b3s, bsum, bars, total;
};
__host__ __device__ inline TcBwdSmem tc_bwd_layout(int Npad, int CpB, int EW) {
    TcBwdSmem L;
    const uint32_t NP64 = ((uint32_t)Npad + 63u) & ~63u;
    uint32_t o = 0;
    L.Ws = o; o += (uint32_t)Npad * kTcKP * 2;
    L.As = o; o += kTcM * kTcKP * 2;
    L.Gs = o; o += kTcM * NP64 * 2;
    L.dXs = o; o += kTcM * (uint32_t)CpB * 4;
    L.b3s = o; o += (uint32_t)Npad * 4;
    L.bsum = o; o += (uint32_t)EW * (uint32_t)Npad * 4;
    o = (o + 15u) & ~15u;
    L.bars = o; o += 16 * 8;
    L.total = o;
    return L;
}

// EW = number of epilogue warps (8 or 16; with 16, four warps share each TMEM lane quarter and split the columns).  Measured on
// cfg 5: 8 warps 36.8 us per launch, 16 warps 40.6 us — the epilogue is not occupancy-bound, so 8 is the default (NCDE_BWD_EW=16
// selects the other instantiation for A/B runs).
template <int EW, bool GKE = false>   // GKE: form gk before waiting for the accumulators (experimental, see TcFieldArgs::gk_early)
__global__ void __launch_bounds__(EW * 32 + 32, 1) tc_field_bwd_kernel(const __grid_constant__ TcFieldArgs a,
                                                                       const __grid_constant__ TcMaps maps) {
    constexpr int kEpi = EW * 32, kAll = EW * 32 + 32, kCg = EW / 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int Npad = a.Npad;
    constexpr int KP = kTcKP;
    const TcBwdSmem L = tc_bwd_layout(Npad, a.CpB, EW);
    const uint32_t a_bytes = kTcM * kTcKP * 2, x_bytes = kTcM * (uint32_t)a.CpB * 4;
    uint8_t* Ws = smem + L.Ws;                            // [Npad][KP]  bf16 swizzled (rows n, contiguous k)
    uint8_t* As = smem + L.As;                            // [128][KP]   (rows m, contiguous k)
    uint8_t* Gs = smem + L.Gs;                            // [128][NP64] (rows m, contiguous n)
    uint8_t* dXs = smem + L.dXs;                          // [128][CpB] fp32
    float* b3s = reinterpret_cast<float*>(smem + L.b3s);  // [Npad]
    float* bsum = reinterpret_cast<float*>(smem + L.bsum);  // [EW row slices][Npad]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* full_a = bars;
    uint64_t* wg_bar = bars + 1;
    uint64_t* full_x = bars + 2;
    uint64_t* pre_bar = bars + 3;
    uint64_t* g_ready = bars + 4;
    uint64_t* dg_bar = bars + 5;
    uint64_t* done2 = bars + 6;
    uint64_t* fin_bar = bars + 7;
    uint64_t* w_bar = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x, bt = blockIdx.y;

    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        mbar_init(full_a, 1); mbar_init(wg_bar, 1); mbar_init(full_x, 1); mbar_init(pre_bar, 1); mbar_init(g_ready, EW);
        mbar_init(dg_bar, 1); mbar_init(done2, EW); mbar_init(fin_bar, 1); mbar_init(w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < Npad; i += kAll) b3s[i] = a.b3[(size_t)g * Npad + i];
    const int S = a.Hg * a.Cp;                          // valid columns (multiple of 8); [S, Npad) is zero padding
    if (S < Npad && tid < kTcM) *reinterpret_cast<uint4*>(Gs + sw128_off(tid, S >> 3, kTcM)) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool producer = warp == EW && lane == 0;
    if (producer) {
        tma_prefetch_desc(&maps.A);
        tma_prefetch_desc(&maps.X);
        mbar_expect_tx(w_bar, (uint32_t)Npad * kTcKP * 2);
        tma_load_3d(Ws, &maps.W, w_bar, 0, g * Npad, 0);
        tma_load_3d(Ws + (size_t)Npad * 128, &maps.W, w_bar, 64, g * Npad, 0);
    }
    const int64_t row_begin = (int64_t)bt * a.Bt;
    const bool pre = a.prefetch != 0 && a.ctrl == nullptr && row_begin < a.B;
    if (producer && pre) {
        // first tile: activations and dX/dt are saved records of the forward pass, the recompute MMA needs nothing else
        mbar_expect_tx(full_a, a_bytes);
        tma_load_3d(As, &maps.A, full_a, 0, (int)row_begin, a.rec_a);
        tma_load_3d(As + a_bytes / 2, &maps.A, full_a, 64, (int)row_begin, a.rec_a);
        mbar_expect_tx(full_x, x_bytes);
        tma_load_3d(dXs, &maps.X, full_x, 0, (int)row_begin, a.rec_x);
        mbar_wait(w_bar, 0);
        mbar_wait(full_a, 0);
        tc_fence_after();
        issue_gemm_kmajor(tmem_base, smem_u32(As), kTcM, smem_u32(Ws), Npad, Npad, KP, false);
        umma_commit(pre_bar);
    }
    pdl_trigger();
    pdl_wait();  // gk comes from the previous kernels

    int64_t row_end = min((int64_t)a.B, row_begin + a.Bt);
    if (a.ctrl && a.ctrl->done) row_end = row_begin;
    const int nt = row_end > row_begin ? (int)((row_end - row_begin + kTcM - 1) / kTcM) : 0;

    if (warp == EW) {
        if (lane == 0 && nt > 0) {
            auto load_A = [&](int i) {
                const int b0 = (int)(row_begin + (int64_t)i * kTcM);
                mbar_expect_tx(full_a, a_bytes);
                tma_load_3d(As, &maps.A, full_a, 0, b0, a.rec_a);
                tma_load_3d(As + a_bytes / 2, &maps.A, full_a, 64, b0, a.rec_a);
            };
            auto load_X = [&](int i) {
                mbar_expect_tx(full_x, x_bytes);
                tma_load_3d(dXs, &maps.X, full_x, 0, (int)(row_begin + (int64_t)i * kTcM), a.rec_x);
            };
            if (!pre) {
                load_A(0);
                load_X(0);
                mbar_wait(w_bar, 0);
            }
            for (int i = 0; i < nt; ++i) {
                const uint32_t ph = (uint32_t)i & 1u;
                const uint32_t As_i = smem_u32(As);
                if (!(pre && i == 0)) {   // tile 0 of a prefetching launch was issued ahead of the dependency
                    mbar_wait(full_a, ph);
                    if (i > 0) mbar_wait(done2, ph ^ 1u);    // epilogue 2 of tile i-1 has read P out of the accumulator
                    tc_fence_after();
                    issue_gemm_kmajor(tmem_base, As_i, kTcM, smem_u32(Ws), Npad, Npad, KP, false);
                    umma_commit(pre_bar);
                }
                mbar_wait(g_ready, ph);                      // G tile of tile i written by all 8 warps; dX/dt tile no longer needed
                tc_fence_after();
                if (i + 1 < nt) load_X(i + 1);
                // dgrad: D[128 x KP] = Gs (K-major over n) . Ws (MN-major: N = k contiguous, K = n rows)
                {
                    const uint32_t idesc = make_idesc(kTcM, KP, 0, 1);
                    for (int ks = 0; ks < Npad / 16; ++ks) {
                        const uint32_t a_off = (uint32_t)(ks >> 2) * (uint32_t)kTcM * 128u + (uint32_t)(ks & 3) * 32u;
                        const uint32_t b_off = (uint32_t)ks * 2048u;
                        umma_bf16(tmem_base, make_sdesc(smem_u32(Gs) + a_off, 16, 1024),
                                  make_sdesc(smem_u32(Ws) + b_off, (uint32_t)Npad * 128u, 1024), idesc, ks > 0 ? 1u : 0u);
                    }
                }
                umma_commit(dg_bar);
                // wgrad: D[KP x Npad] += As^T (MN-major: M = k contiguous, K = m rows) . Gs (MN-major: N = n contiguous)
                {
                    const uint32_t idesc = make_idesc(KP, Npad, 1, 1);
                    for (int ks = 0; ks < kTcM / 16; ++ks) {
                        const uint32_t off = (uint32_t)ks * 2048u;
                        umma_bf16(tmem_base + kTcDwCol, make_sdesc(As_i + off, (uint32_t)kTcM * 128u, 1024),
                                  make_sdesc(smem_u32(Gs) + off, (uint32_t)kTcM * 128u, 1024), idesc, (ks > 0 || i > 0) ? 1u : 0u);
                    }
                }
                if (i + 1 < nt) {
                    // the activation tile is free once wgrad(i) has completed; the next one lands during epilogue 2
                    umma_commit(wg_bar);
                    mbar_wait(wg_bar, ph);
                    load_A(i + 1);
                }
            }
            umma_commit(fin_bar);
        }
    } else {
        const int cg = warp >> 2;                    // column group: the kCg warps of a TMEM lane quarter split the columns
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        // columns are handed out as units (h, channel part): P parts per h when there are fewer h's than column groups
        const int P = a.Hg >= kCg ? 1 : kCg / a.Hg;
        const int U = a.Hg * P;
        const int u_begin = cg * U / kCg, u_end = (cg + 1) * U / kCg;
        const int cw = (((a.Cp + P - 1) / P) + 7) & ~7;
        // column split of the final dW^T read-out (16-column chunks)
        const int col_begin = (cg * Npad / kCg) & ~15;
        const int col_end = cg == kCg - 1 ? Npad : (((cg + 1) * Npad / kCg) & ~15);
        // bias gradient: thread (chunk of 8 columns = lane, slice of 128 / EW rows = warp) keeps its partial column sums in registers
        const int n_chunk8 = Npad >> 3;                     // <= 30
        float bacc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bacc[j] = 0.f;
        const uint32_t b3_s = smem_u32(b3s), gs_s = smem_u32(Gs);
        const uint32_t dx_s = smem_u32(dXs) + (uint32_t)row * (uint32_t)a.CpB * 4u;

        for (int i = 0; i < nt; ++i) {
            const uint32_t ph = (uint32_t)i & 1u;
            const int64_t b0 = row_begin + (int64_t)i * kTcM;
            const int64_t b = b0 + row;
            const bool row_ok = b < a.B;
            // gk of this thread's (at most 4) units, loaded while the recompute MMA of the tile is still in flight
            float gk_pre[4] = {0.f, 0.f, 0.f, 0.f};
            if (GKE && a.gy1T) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const int u = u_begin + k4;
                    const int h = g * a.Hg + u / P;
                    if (u < u_end && row_ok && h < a.H) {
                        const size_t off = (size_t)h * a.Bp + b;
                        float gk = a.gcoef * __ldg(a.gy1T + off);
                        for (int q = 0; q < a.n_dz; ++q) gk = fmaf(a.dzcoef[q], __ldg(a.dzT[q] + off), gk);
                        gk_pre[k4] = gk;
                    }
                }
            }
            mbar_wait(full_x, ph);
            mbar_wait(pre_bar, ph);
            tc_fence_after();
            // ---- epilogue 1: G tile; per h, branch-free over the channels ----
            for (int u = u_begin; u < u_end; ++u) {
                const int hl = u / P, c_begin = (u % P) * cw, c_end = min(a.Cp, c_begin + cw);
                const int h = g * a.Hg + hl;
                float gk = 0.f;
                if (GKE && a.gy1T && u - u_begin < 4) {
                    const int k4 = u - u_begin;
                    gk = k4 == 0 ? gk_pre[0] : (k4 == 1 ? gk_pre[1] : (k4 == 2 ? gk_pre[2] : gk_pre[3]));
                } else if (row_ok && h < a.H) {
                    const size_t off = (size_t)h * a.Bp + b;
                    if (a.gy1T) {
                        gk = a.gcoef * __ldg(a.gy1T + off);
                        for (int q = 0; q < a.n_dz; ++q) gk = fmaf(a.dzcoef[q], __ldg(a.dzT[q] + off), gk);
                    } else {
                        gk = __ldg(a.gkT + off);
                    }
                }
                const int colbase = hl * a.Cp;
                const int n32 = (c_end - c_begin) / 32;
                {
                    TcChunk<32> A, Bk;
                    if (n32 > 0) issue32(A, 0);
                    for (int j = 0; j < n32; j += 2) {
                        tmem_wait_ld<32>(A.r);
                        if (j + 1 < n32) issue32(Bk, j + 1);
                        bwd_finish<32>(A, gk, gs_s, row, colbase + c_begin + 32 * j);
                        if (j + 1 < n32) {
                            tmem_wait_ld<32>(Bk.r);
                            if (j + 2 < n32) issue32(A, j + 2);
                            bwd_finish<32>(Bk, gk, gs_s, row, colbase + c_begin + 32 * (j + 1));
                        }
                    }
                }
                for (int c0 = c_begin + 32 * n32; c0 + 8 <= c_end; c0 += 8) {
                    TcChunk<8> T;
                    chunk_issue<8>(T, lane_addr + (uint32_t)(colbase + c0), b3_s + 4u * (colbase + c0), dx_s + 4u * c0);
                    tmem_wait_ld<8>(T.r);
                    bwd_finish<8>(T, gk, gs_s, row, colbase + c0);
                }
            }
            fence_async_smem();     // generic-proxy writes of G (and reads of dX/dt) before the async-proxy MMA / TMA
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(g_ready);
            // bias gradient from the bf16 G tile while the tensor core works: every warp needs ALL rows -> epilogue-wide barrier
            named_bar_sync(1, kEpi);
            if (lane < n_chunk8) {
#pragma unroll 4
                for (int r = warp * (kTcM / EW); r < (warp + 1) * (kTcM / EW); ++r) {
                    uint32_t w4[4];
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w4[0]), "=r"(w4[1]), "=r"(w4[2]), "=r"(w4[3])
                                 : "r"(gs_s + sw128_off(r, lane, kTcM)));
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[j]));
                        bacc[2 * j] += f.x;
                        bacc[2 * j + 1] += f.y;
                    }
                }
            }
            mbar_wait(dg_bar, ph);
            tc_fence_after();
            // ---- epilogue 2: partial input gradient of this h-group ----
            {
                const int kb = cg * (KP / kCg), ke = kb + KP / kCg;      // 64 (8 warps) or 32 (16 warps) columns
                if (a.p_transposed) {
                    // P^T[g][k][b] in bf16 (the consumer rounds the 64-group sum to bf16 anyway; the partials are summed in fp32):
                    // the 32 lanes of a warp (= consecutive rows) write 64 contiguous bytes per column; row pitch Bp keeps
                    // p_reduce's 16-byte loads aligned
                    __nv_bfloat16* pcol = reinterpret_cast<__nv_bfloat16*>(a.P) + ((size_t)g * 128 + kb) * a.Bp + (size_t)b;
                    uint32_t r0[32], r1[32];
                    tmem_ld32_issue(lane_addr + (uint32_t)kb, r0);
                    tmem_wait_ld<32>(r0);
                    if (ke - kb > 32) tmem_ld32_issue(lane_addr + (uint32_t)kb + 32u, r1);
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) pcol[(size_t)j * a.Bp] = __float2bfloat16(__uint_as_float(r0[j]));
                    }
                    if (ke - kb > 32) {
                        tmem_wait_ld<32>(r1);
                        if (row_ok) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) pcol[(size_t)(32 + j) * a.Bp] = __float2bfloat16(__uint_as_float(r1[j]));
                        }
                    }
                } else {
                    float* prow = a.P + ((size_t)g * a.B + (size_t)b) * a.DFP;
                    for (int k0 = kb; k0 < ke; k0 += 16) {
                        float v[16];
                        tmem_ld16(lane_addr + (uint32_t)k0, v);
                        if (row_ok) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                *reinterpret_cast<float4*>(prow + k0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(done2);
        }
        // ---- dW^T (TMEM lanes = k, columns = n) -> global accumulator [bt][g*Npad + n][k];  bias gradient ----
        if (nt > 0) {
            if (lane < n_chunk8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) bsum[warp * Npad + lane * 8 + j] = bacc[j];
            }
            mbar_wait(fin_bar, 0);
            tc_fence_after();
            named_bar_sync(1, kEpi);
            const int k = row;
            if (k < KP) {
                for (int n0 = col_begin; n0 < col_end; n0 += 16) {
                    float v[16];
                    tmem_ld16(lane_addr + kTcDwCol + (uint32_t)n0, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // exactly one thread ever adds to this element in this launch: a reduction without return value is
                        // deterministic and does not stall on the read
                        float* p = a.dW3acc + (((size_t)bt * a.n_hg + g) * Npad + n0 + j) * a.DFP + k;
                        atomicAdd(p, v[j]);
                    }
                }
            }
            for (int n = tid; n < Npad; n += kEpi) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < EW; ++w) s += bsum[w * Npad + n];
                atomicAdd(a.db3acc + ((size_t)bt * a.n_hg + g) * Npad + n, s);
            }
        }
    }
    tc_fence_before();
    __syncwarp();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static inline size_t tc_bwd_smem_bytes(int Npad, int CpB, int EW) { return 1024 + tc_bwd_layout(Npad, CpB, EW).total; }

}  // namespace ncde
