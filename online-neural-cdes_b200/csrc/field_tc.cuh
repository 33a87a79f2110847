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
#include <cuda_bf16.h>

#include "solve_kernels.cuh"

namespace ncde {

constexpr int kTcThreads = 256;
constexpr int kTcM = 128;  // batch rows per MMA tile (= TMEM lanes)

struct TcFieldArgs {
    int B, Bp, H, Cp, Hg, n_hg, Npad, KP, DF, Bt;
    const __nv_bfloat16* Wbf;  // [n_hg][Npad][KP]
    const float* b3;           // [n_hg][Npad]
    const __nv_bfloat16* abf;  // [Bp][KP] final-layer input, bf16 row-major
    const float* dXT;          // [Cp][Bp]
    float* koutT;              // [H][Bp]            forward
    const float* gkT;          // [H][Bp]            backward
    float* P;                  // [n_hg][B][DFP]     backward
    float* dW3acc;             // [n_bt][Np][DFP]
    float* db3acc;             // [n_bt][Np]
    int DFP;
    const AdaptCtrl* ctrl;     // adaptive solver: skip all work once ctrl->done
};

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

// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes == 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// cooperative asynchronous copy of a [rows][KP] bf16 row-major global tile into the swizzled shared layout; rows >=
// valid_rows are zero-filled (batch padding must not reach the weight-gradient reduction: 0 * garbage could be NaN)
__device__ __forceinline__ void load_tile_sw128(uint8_t* smem_tile, const __nv_bfloat16* __restrict__ g, int rows, int KP,
                                                int valid_rows, int tid, int nthreads) {
    const int cpr = KP / 8;  // 16-byte chunks per row
    for (int idx = tid; idx < rows * cpr; idx += nthreads) {
        const int r = idx / cpr, ch = idx % cpr;
        const bool ok = r < valid_rows;
        cp_async16(smem_tile + sw128_off(r, ch, rows), g + (size_t)(ok ? r : 0) * KP + ch * 8, ok ? 16 : 0);
    }
}

// dXs[c][r] = dXT[c][b0 + r] for the 128 rows of a tile
__device__ __forceinline__ void load_dx_tile(float* dXs, const float* __restrict__ dXT, int Cp, int Bp, int64_t b0, int tid,
                                             int nthreads) {
    for (int idx = tid; idx < Cp * (kTcM / 4); idx += nthreads) {
        const int c = idx / (kTcM / 4), q = idx % (kTcM / 4);
        cp_async16(dXs + (size_t)idx * 4, dXT + (size_t)c * Bp + b0 + q * 4, 16);
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

// sum_j tanh(D[row][col0 + j] + b3[j]) * dX[j][row] over W consecutive channels; dx points at dXs[c0][row]
template <int W>
__device__ __forceinline__ float fwd_chunk(uint32_t taddr, const float* __restrict__ b3, const float* __restrict__ dx) {
    float v[W];
    tmem_ldw<W>(taddr, v);
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int j = 0; j < W; j += 2) {
        acc0 = fmaf(tanh_fast(v[j] + b3[j]), dx[j * kTcM], acc0);
        acc1 = fmaf(tanh_fast(v[j + 1] + b3[j + 1]), dx[(j + 1) * kTcM], acc1);
    }
    return acc0 + acc1;
}

// ---------------------------------------------------------------------------------------------------------------
// forward:  k[b, h] = sum_c tanh( a[b,:] . W3[(h,c),:] + b3[(h,c)] ) * dX[b, c]
// CTA (g, bt): W slice of h-group g resident in shared memory; 128-row tiles of the batch stream through.
// Epilogue: thread = one TMEM lane (batch row); the two warp-groups split the columns.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1) tc_field_fwd_kernel(const __grid_constant__ TcFieldArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve (every operand tile 1024-byte aligned)
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int Npad = a.Npad, KP = a.KP;
    uint8_t* Ws = smem;                                 // [Npad][KP] bf16 swizzled
    uint8_t* As = Ws + (size_t)Npad * KP * 2;           // [128][KP]
    float* b3s = reinterpret_cast<float*>(As + (size_t)kTcM * KP * 2);  // [Npad]
    float* part = b3s + Npad;                           // [2][Hg][128] per-warp-group partial sums
    float* dXs = part + 2 * a.Hg * kTcM;                // [Cp][128] dX/dt of the current row tile
    uint64_t* mbar = reinterpret_cast<uint64_t*>(dXs + (size_t)a.Cp * kTcM);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x, bt = blockIdx.y;
    // power of two >= Npad + 16: the last 16-column epilogue load may start up to 15 columns before Npad
    const uint32_t ncols = tc_tmem_cols(Npad + 16);

    if (warp == 0) tmem_alloc(tmem_slot, ncols);
    if (tid == 0) mbar_init(mbar, 1);
    load_tile_sw128(Ws, a.Wbf + (size_t)g * Npad * KP, Npad, KP, Npad, tid, kTcThreads);
    for (int i = tid; i < Npad; i += kTcThreads) b3s[i] = a.b3[(size_t)g * Npad + i];
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();  // the activation tiles come from hidden_fwd

    const int wg = warp >> 2;                    // warp-group 0/1
    const int row = (warp & 3) * 32 + lane;      // TMEM lane == row inside the tile
    // share of this warp-group: a set of whole h's when Hg >= 2, else half of the channels of the single h
    int h_begin, h_end, c_begin, c_end;
    if (a.Hg >= 2) { h_begin = wg == 0 ? 0 : a.Hg / 2; h_end = wg == 0 ? a.Hg / 2 : a.Hg; c_begin = 0; c_end = a.Cp; }
    else { const int half = ((a.Cp / 2 + 7) / 8) * 8; h_begin = 0; h_end = 1; c_begin = wg == 0 ? 0 : half; c_end = wg == 0 ? half : a.Cp; }

    uint32_t phase = 0;
    const int64_t row_begin = (int64_t)bt * a.Bt;
    const int64_t row_end = (a.ctrl && a.ctrl->done) ? row_begin : min((int64_t)a.Bp, row_begin + a.Bt);
    for (int64_t b0 = row_begin; b0 < row_end && b0 < a.B; b0 += kTcM) {
        // asynchronous fills: activation tile (MMA operand) and dX/dt of these 128 rows (epilogue operand)
        load_tile_sw128(As, a.abf + (size_t)b0 * KP, kTcM, KP, (int)min((int64_t)kTcM, (int64_t)a.B - b0), tid, kTcThreads);
        load_dx_tile(dXs, a.dXT, a.Cp, a.Bp, b0, tid, kTcThreads);
        cp_async_wait_all();
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor(tmem_base, smem_u32(As), kTcM, smem_u32(Ws), Npad, Npad, KP, false);
            umma_commit(mbar);
        }
        mbar_wait(mbar, phase);
        phase ^= 1;
        tc_fence_after();

        // epilogue: per h of this warp-group, a branch-free sweep over its channels (static smem offsets -> ILP)
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int hl = h_begin; hl < h_end; ++hl) {
            float acc = 0.f;
            const int colbase = hl * a.Cp;
            int c0 = c_begin;
            for (; c0 + 32 <= c_end; c0 += 32) acc += fwd_chunk<32>(lane_addr + (uint32_t)(colbase + c0), b3s + colbase + c0, dXs + c0 * kTcM + row);
            for (; c0 + 8 <= c_end; c0 += 8) acc += fwd_chunk<8>(lane_addr + (uint32_t)(colbase + c0), b3s + colbase + c0, dXs + c0 * kTcM + row);
            part[(wg * a.Hg + hl) * kTcM + row] = acc;
        }
        tc_fence_before();
        __syncthreads();
        // combine and write k^T[h][b]
        for (int idx = tid; idx < a.Hg * kTcM; idx += kTcThreads) {
            const int h_l = idx / kTcM, m = idx % kTcM;
            float s;
            if (a.Hg >= 2) s = part[((h_l < a.Hg / 2 ? 0 : 1) * a.Hg + h_l) * kTcM + m];
            else s = part[m] + part[kTcM + m];
            const int h = g * a.Hg + h_l;
            if (h < a.H && b0 + m < a.B) a.koutT[(size_t)h * a.Bp + b0 + m] = s;
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    cp_async_wait_all();  // the weight tile copy must have landed before the CTA may exit
    if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

static inline size_t tc_fwd_smem_bytes(int Npad, int KP, int Hg, int Cp) {
    return 1024 + (size_t)Npad * KP * 2 + (size_t)kTcM * KP * 2 + (size_t)Npad * 4 + (size_t)2 * Hg * kTcM * 4 +
           (size_t)Cp * kTcM * 4 + 64;
}

// ---------------------------------------------------------------------------------------------------------------
// backward of one RK stage:
//   MMA1   pre[m][n]  = As . Ws^T                      (A, B K-major)                 -> TMEM cols [0, Npad)
//   epi 1  G = gk * dX * (1 - tanh^2(pre + b3)) -> bf16 tile Gs[m][n];  per-warp column sums for db3
//   dgrad  P[m][k]    = Gs . Ws        (A = Gs K-major over n, B = Ws MN-major)       -> TMEM cols [0, KP)
//   wgrad  dW^T[k][n] += As^T . Gs     (A = As MN-major, B = Gs MN-major)             -> TMEM cols [256, 256+Npad)
//   epi 2  P -> global partial of dL/d(act) for this h-group
// dW^T stays in TMEM across all row tiles of the CTA and is added to the global accumulator once at the end.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kTcDwCol = 256;

// Column sums of a W-column x 32-row (lane) block by recursive halving: after log2(W) exchange steps every lane holds
// one column summed over a lane subset, the remaining levels are plain xor-reductions.  W + log2(32/W) shuffles
// instead of 5 W.  dst points at the first of the W columns of this warp's accumulator row.
template <int W>
__device__ __forceinline__ void colsum_butterfly(float (&v)[W], int lane, float* dst) {
    const uint32_t full = 0xffffffffu;
    int col = 0;
    int bit = 16;
#pragma unroll
    for (int width = W; width > 1; width >>= 1) {
        const int half = width >> 1;
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            const float send = up ? v[j] : v[j + half];
            const float keep = up ? v[j + half] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, bit);
        }
        col += up ? half : 0;
        bit >>= 1;
    }
    int low = 0;
#pragma unroll
    for (; bit >= 1; bit >>= 1) { v[0] += __shfl_xor_sync(full, v[0], bit); low |= bit; }
    if ((lane & low) == 0) dst[col] += v[0];
}

// One W-column chunk of epilogue 1 for TMEM lane `row`: G = gk * dX * sech^2(pre + b3) -> bf16 into the swizzled G
// tile (16-byte stores, n0 is a multiple of 8) and column sums for the bias gradient.
template <int W>
__device__ __forceinline__ void bwd_chunk(uint32_t taddr, const float* __restrict__ b3, const float* __restrict__ dx,
                                          float gk, bool row_ok, uint8_t* Gs, int row, int n0, float* bsum_row, int lane) {
    float v[W];
    tmem_ldw<W>(taddr, v);
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const float gv = gk * dx[j * kTcM] * sech2_fast(v[j] + b3[j]);
        v[j] = row_ok ? gv : 0.f;   // select, not multiply: padded rows hold uninitialised dX
    }
#pragma unroll
    for (int j8 = 0; j8 < W / 8; ++j8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j8 * 8 + 2 * j], v[j8 * 8 + 2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(Gs + sw128_off(row, (n0 >> 3) + j8, kTcM)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    if constexpr (W == 32) {
        float lo[16], hi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { lo[j] = v[j]; hi[j] = v[16 + j]; }
        colsum_butterfly<16>(lo, lane, bsum_row + n0);
        colsum_butterfly<16>(hi, lane, bsum_row + n0 + 16);
    } else {
        colsum_butterfly<W>(v, lane, bsum_row + n0);
    }
}

__global__ void __launch_bounds__(kTcThreads, 1) tc_field_bwd_kernel(const __grid_constant__ TcFieldArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int Npad = a.Npad, KP = a.KP;
    const int NP64 = (Npad + 63) & ~63;
    uint8_t* Ws = smem;                                   // [Npad][KP]  bf16 swizzled (rows n, contiguous k)
    uint8_t* As = Ws + (size_t)Npad * KP * 2;             // [128][KP]   (rows m, contiguous k)
    uint8_t* Gs = As + (size_t)kTcM * KP * 2;             // [128][NP64] (rows m, contiguous n)
    float* b3s = reinterpret_cast<float*>(Gs + (size_t)kTcM * NP64 * 2);  // [Npad]
    float* bsum = b3s + Npad;                             // [8 warps][Npad]
    float* dXs = bsum + 8 * Npad;                         // [Cp][128]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(dXs + (size_t)a.Cp * kTcM);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x, bt = blockIdx.y;

    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 0) mbar_init(mbar, 1);
    load_tile_sw128(Ws, a.Wbf + (size_t)g * Npad * KP, Npad, KP, Npad, tid, kTcThreads);
    for (int i = tid; i < Npad; i += kTcThreads) b3s[i] = a.b3[(size_t)g * Npad + i];
    for (int i = tid; i < 8 * Npad; i += kTcThreads) bsum[i] = 0.f;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();  // gk comes from the previous kernels

    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int S = a.Hg * a.Cp;                          // valid columns (multiple of 8); [S, Npad) is zero padding
    int h_begin, h_end, c_begin, c_end;
    if (a.Hg >= 2) { h_begin = wg == 0 ? 0 : a.Hg / 2; h_end = wg == 0 ? a.Hg / 2 : a.Hg; c_begin = 0; c_end = a.Cp; }
    else { const int half = ((a.Cp / 2 + 7) / 8) * 8; h_begin = 0; h_end = 1; c_begin = wg == 0 ? 0 : half; c_end = wg == 0 ? half : a.Cp; }
    // column split of the final dW^T read-out (16-column chunks)
    const int split = (Npad / 2) & ~15;
    const int col_begin = wg == 0 ? 0 : split;
    const int col_end = wg == 0 ? split : Npad;
    if (S < Npad && tid < kTcM) *reinterpret_cast<uint4*>(Gs + sw128_off(tid, S >> 3, kTcM)) = make_uint4(0, 0, 0, 0);

    uint32_t phase = 0;
    bool first_tile = true;
    const int64_t row_begin = (int64_t)bt * a.Bt;
    const int64_t row_end = (a.ctrl && a.ctrl->done) ? row_begin : min((int64_t)a.Bp, row_begin + a.Bt);
    for (int64_t b0 = row_begin; b0 < row_end && b0 < a.B; b0 += kTcM) {
        load_tile_sw128(As, a.abf + (size_t)b0 * KP, kTcM, KP, (int)min((int64_t)kTcM, (int64_t)a.B - b0), tid, kTcThreads);
        load_dx_tile(dXs, a.dXT, a.Cp, a.Bp, b0, tid, kTcThreads);
        cp_async_wait_all();
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor(tmem_base, smem_u32(As), kTcM, smem_u32(Ws), Npad, Npad, KP, false);
            umma_commit(mbar);
        }
        mbar_wait(mbar, phase);
        phase ^= 1;
        tc_fence_after();

        // ---- epilogue 1: G tile + bias-gradient column sums; per h, branch-free over the channels ----
        const int64_t b = b0 + row;
        const bool row_ok = b < a.B;
        for (int hl = h_begin; hl < h_end; ++hl) {
            const int h = g * a.Hg + hl;
            const float gk = (row_ok && h < a.H) ? __ldg(a.gkT + (size_t)h * a.Bp + b) : 0.f;
            const int colbase = hl * a.Cp;
            int c0 = c_begin;
            for (; c0 + 32 <= c_end; c0 += 32)
                bwd_chunk<32>(lane_addr + (uint32_t)(colbase + c0), b3s + colbase + c0, dXs + c0 * kTcM + row, gk, row_ok, Gs, row,
                              colbase + c0, bsum + warp * Npad, lane);
            for (; c0 + 8 <= c_end; c0 += 8)
                bwd_chunk<8>(lane_addr + (uint32_t)(colbase + c0), b3s + colbase + c0, dXs + c0 * kTcM + row, gk, row_ok, Gs, row,
                             colbase + c0, bsum + warp * Npad, lane);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
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
            // wgrad: D[KP x Npad] += As^T (MN-major: M = k contiguous, K = m rows) . Gs (MN-major: N = n contiguous)
            {
                const uint32_t idesc = make_idesc(KP, Npad, 1, 1);
                for (int ks = 0; ks < kTcM / 16; ++ks) {
                    const uint32_t off = (uint32_t)ks * 2048u;
                    umma_bf16(tmem_base + kTcDwCol, make_sdesc(smem_u32(As) + off, (uint32_t)kTcM * 128u, 1024),
                              make_sdesc(smem_u32(Gs) + off, (uint32_t)kTcM * 128u, 1024), idesc,
                              (ks > 0 || !first_tile) ? 1u : 0u);
                }
            }
            umma_commit(mbar);
        }
        first_tile = false;
        mbar_wait(mbar, phase);
        phase ^= 1;
        tc_fence_after();

        // ---- epilogue 2: partial input gradient of this h-group ----
        {
            const int kb = wg * (KP / 2), ke = kb + KP / 2;
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
        tc_fence_before();
        __syncthreads();
    }
    // ---- dW^T (TMEM lanes = k, columns = n) -> global accumulator [bt][g*Npad + n][k];  bias gradient ----
    if (!first_tile) {
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
        for (int n = tid; n < Npad; n += kTcThreads) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += bsum[w * Npad + n];
            atomicAdd(a.db3acc + ((size_t)bt * a.n_hg + g) * Npad + n, s);
        }
    }
    tc_fence_before();
    __syncthreads();
    cp_async_wait_all();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static inline size_t tc_bwd_smem_bytes(int Npad, int KP, int Hg, int Cp) {
    (void)Hg;
    const int NP64 = (Npad + 63) & ~63;
    return 1024 + (size_t)Npad * KP * 2 + (size_t)kTcM * KP * 2 + (size_t)kTcM * NP64 * 2 + (size_t)Npad * 4 +
           (size_t)8 * Npad * 4 + (size_t)Cp * kTcM * 4 + 64;
}

}  // namespace ncde
