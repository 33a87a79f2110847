// Hidden layers of the vector field on the tensor cores (bf16 operands, fp32 accumulate), fixed-grid bf16 path.
//
// Every activation of the vector-field MLP lives in global memory as a bf16 ROW-major tile [B][128] (feature axis padded
// with zeros to 128): the same bytes are a K-major A operand for the forward GEMM, a K-major A operand for the
// input-gradient GEMM and an MN-major operand for the weight-gradient GEMM (K = batch), so no layout conversion happens
// anywhere between the kernels.  Weights are packed once per call as bf16 [layer][128 out][128 in].
//
// CTA = one 128-row batch tile (the hidden layers are 0.1 GFLOP per stage: this kernel is latency-, not throughput-bound;
// its CTAs start early on SMs the final-layer kernel leaves idle and prefetch the weights under programmatic dependent
// launch).  The layer chain runs out of shared memory: MMA l -> TMEM -> epilogue (bias, activation, bf16) -> operand tile
// of MMA l+1, which is also written to the saved record by a TMA store.
#pragma once
#include "field_tc.cuh"

namespace ncde {

constexpr int kTcHidMaxLayers = NCDE_MAX_LAYERS;

struct TcHiddenMaps {
    CUtensorMap W;                           // {128 in, 128 out, n_layers} bf16, box {64, 128, 1}, 128B swizzle
    CUtensorMap act[kTcHidMaxLayers + 1];    // act[l]: input of layer l (act[F] = input of the final layer); {128, B, recs}
};

struct TcHiddenArgs {
    int B, F, rec;
    int act[kTcHidMaxLayers];                // activation of hidden layer l
    const float* bias;                       // [F][128] fp32, zero padded
    const AdaptCtrl* ctrl;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

constexpr uint32_t kTcHidTile = kTcM * kTcKP * 2;   // 32 KB: one [128][128] bf16 operand tile

struct TcHidSmem { uint32_t Wt, At, bias, bars, total; };
__host__ __device__ inline TcHidSmem tc_hid_layout() {
    TcHidSmem L;
    uint32_t o = 0;
    L.Wt = o; o += 2 * kTcHidTile;
    L.At = o; o += 2 * kTcHidTile;
    L.bias = o; o += kTcHidMaxLayers * 128 * 4;
    L.bars = o; o += 16 * 8;
    L.total = o;
    return L;
}
static inline size_t tc_hid_smem_bytes() { return 1024 + tc_hid_layout().total; }

// pack hidden layer l: Wh[l][o][i] = bf16(W[o][i]) zero padded to 128 x 128; bh[l][o] = bias[o]
__global__ void pack_hidden_bf16_kernel(const float* __restrict__ W, const float* __restrict__ bias, __nv_bfloat16* __restrict__ Wh,
                                        float* __restrict__ bh, int Dout, int Din) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 128 * 128) {
        const int o = idx >> 7, i = idx & 127;
        Wh[idx] = __float2bfloat16((o < Dout && i < Din) ? W[(size_t)o * Din + i] : 0.f);
    }
    if (idx < 128) bh[idx] = (bias && idx < Dout) ? bias[idx] : 0.f;
}

// z0 (B, H) fp32 row-major -> bf16 row-major [B][128] (zero padded): input record of the first stage
__global__ void to_bf16_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int H) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (int64_t)B * 128) {
        const int h = (int)(idx & 127);
        const int64_t b = idx >> 7;
        dst[idx] = __float2bfloat16(h < H ? src[b * H + h] : 0.f);
    }
}

__global__ void __launch_bounds__(kTcThreads, 1) tc_hidden_fwd_kernel(const __grid_constant__ TcHiddenArgs a,
                                                                      const __grid_constant__ TcHiddenMaps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const TcHidSmem L = tc_hid_layout();
    uint8_t* Wt = smem + L.Wt;
    uint8_t* At = smem + L.At;
    float* bias_s = reinterpret_cast<float*>(smem + L.bias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* w_full = bars;          // [2]
    uint64_t* a_full = bars + 2;      // TMA: input tile of layer 0
    uint64_t* mma_bar = bars + 3;
    uint64_t* a_ready = bars + 4;     // epilogue -> producer: operand tile of the next layer written (8 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b0 = blockIdx.x * kTcM;
    const int F = a.F;

    if (warp == 0) tmem_alloc(tmem_slot, 128);
    if (tid == 0) {
        mbar_init(w_full, 1); mbar_init(w_full + 1, 1); mbar_init(a_full, 1); mbar_init(mma_bar, 1); mbar_init(a_ready, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < F * 128; i += kTcThreads) bias_s[i] = a.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool producer = warp == 8 && lane == 0;
    auto load_W = [&](int l) {
        uint8_t* dst = Wt + (size_t)(l & 1) * kTcHidTile;
        mbar_expect_tx(w_full + (l & 1), kTcHidTile);
        tma_load_3d(dst, &maps.W, w_full + (l & 1), 0, 0, l);
        tma_load_3d(dst + kTcHidTile / 2, &maps.W, w_full + (l & 1), 64, 0, l);
    };
    if (producer) {
        // the packed weights are older than the predecessor kernel: fetch them ahead of the dependency
        tma_prefetch_desc(&maps.act[0]);
        if (F > 0) load_W(0);
        if (F > 1) load_W(1);
    }
    pdl_trigger();
    pdl_wait();   // the stage input comes from the previous kernel
    if (a.ctrl && a.ctrl->done) {   // uniform
        if (producer) {   // the weight tiles already in flight must land before the CTA may exit
            if (F > 0) mbar_wait(w_full, 0);
            if (F > 1) mbar_wait(w_full + 1, 0);
        }
        __syncthreads();
        if (warp == 0) tmem_dealloc(tmem_base, 128);
        return;
    }

    if (warp == 8) {
        if (lane == 0 && F > 0) {
            mbar_expect_tx(a_full, kTcHidTile);
            tma_load_3d(At, &maps.act[0], a_full, 0, b0, a.rec);
            tma_load_3d(At + kTcHidTile / 2, &maps.act[0], a_full, 64, b0, a.rec);
            for (int l = 0; l < F; ++l) {
                if (l == 0) mbar_wait(a_full, 0);
                else {
                    mbar_wait(a_ready, (uint32_t)(l - 1) & 1u);     // output of layer l-1 = operand tile l & 1, fenced by its writers
                    tma_store_3d(&maps.act[l], At + (size_t)(l & 1) * kTcHidTile, 0, b0, a.rec);
                    tma_store_3d(&maps.act[l], At + (size_t)(l & 1) * kTcHidTile + kTcHidTile / 2, 64, b0, a.rec);
                    bulk_commit();
                }
                mbar_wait(w_full + (l & 1), (uint32_t)(l >> 1) & 1u);
                tc_fence_after();
                issue_gemm_kmajor(tmem_base, smem_u32(At + (size_t)(l & 1) * kTcHidTile), kTcM, smem_u32(Wt + (size_t)(l & 1) * kTcHidTile),
                                  128, 128, kTcKP, false);
                // the epilogue of this layer overwrites the tile the PREVIOUS store read from: that store must be done reading
                bulk_wait_read<1>();
                umma_commit(mma_bar);
                if (l + 2 < F) {
                    mbar_wait(mma_bar, (uint32_t)l & 1u);           // weight buffer l & 1 free
                    load_W(l + 2);
                }
            }
            mbar_wait(a_ready, (uint32_t)(F - 1) & 1u);
            tma_store_3d(&maps.act[F], At + (size_t)(F & 1) * kTcHidTile, 0, b0, a.rec);
            tma_store_3d(&maps.act[F], At + (size_t)(F & 1) * kTcHidTile + kTcHidTile / 2, 64, b0, a.rec);
            bulk_commit();
            bulk_wait<0>();   // the records must be complete (and shared memory no longer read) when the CTA exits
        }
    } else {
        const int wg = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int l = 0; l < F; ++l) {
            mbar_wait(mma_bar, (uint32_t)l & 1u);
            tc_fence_after();
            const uint32_t dst = smem_u32(At + (size_t)((l + 1) & 1) * kTcHidTile);
            const uint32_t bias_a = smem_u32(bias_s + l * 128);
            const int act = a.act[l];
#pragma unroll
            for (int c0 = wg * 64; c0 < wg * 64 + 64; c0 += 32) {
                uint32_t r[32];
                tmem_ld32_issue(lane_addr + (uint32_t)c0, r);
                float4 bb[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) bb[q] = lds128(bias_a + 4u * c0 + 16u * q);
                tmem_wait_ld<32>(r);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bq = bb[j8 * 2 + (j >> 2)];
                        const float bj = (j & 3) == 0 ? bq.x : ((j & 3) == 1 ? bq.y : ((j & 3) == 2 ? bq.z : bq.w));
                        v[j] = apply_act(__uint_as_float(r[j8 * 8 + j]) + bj, act);
                    }
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        pk[j] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + sw128_off(row, (c0 >> 3) + j8, kTcM)), "r"(pk[0]),
                                 "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                }
            }
            fence_async_smem();     // generic-proxy writes before the async-proxy MMA / TMA store
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
    }
    tc_fence_before();
    __syncwarp();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}


// ---------------------------------------------------------------------------------------------------------------
// Backward of the hidden layers.
//
// p_reduce: dL/d(a_F) = sum over the h-groups of the final-layer kernel's partials P^T[g][k][b], times act'(a_F), rounded to
// a bf16 row-major record — the first operand tile of the tensor-core chain.  Batch-split over every SM (the 64-way sum is
// 33 MB of L2 reads per stage: this part must not sit on 8 SMs).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad_bf(float out, int act) {
    if (act == NCDE_ACT_RELU) return out > 0.f ? 1.f : 0.f;
    if (act == NCDE_ACT_TANH) return 1.f - out * out;
    return 1.f;
}

struct PReduceArgs {
    int B, Bp, n_hg, DFP, act;     // act: activation of the last hidden layer
    const float* P;                // bf16 P^T [n_hg][128 k][Bp] (a float buffer reinterpreted)
    const __nv_bfloat16* aF;       // [Bp][128] output of the last hidden layer (saved record)
    __nv_bfloat16* dpre;           // [Bp][128]
};

// CTA (x, y) = 128 rows x 8 columns, 128 threads: thread = (column = tid / 16, 8 consecutive rows), so every load is 16 bytes
// (8 bf16 partials) and a half-warp reads 256 contiguous bytes of P^T[g][k][.]; the 64 groups are summed in fp32 registers, the
// 128 x 8 tile is transposed through shared memory and leaves as one 16-byte bf16 store per row.
// one 128-row x 8-column slice of the reduction by threads 0..127 of a CTA; `tile` is [128][9] floats of shared memory;
// `sync128` synchronises exactly these 128 threads
template <typename Sync>
__device__ __forceinline__ void p_reduce_slice(const PReduceArgs& a, int tile_idx, int col_block, int tid, float (*tile)[9], Sync sync128) {
    const int col = tid >> 4, rg = tid & 15;
    const int64_t b0 = (int64_t)tile_idx * 128;
    const int k0 = col_block * 8;
    const int64_t bq = b0 + rg * 8;        // < Bp: the row pitch of P^T covers the padded batch, padded rows are never used
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    if (bq < a.B) {
        const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.P) + (size_t)(k0 + col) * a.Bp + bq);
        const size_t gs = (size_t)128 * a.Bp / 8;
#pragma unroll 16
        for (int g = 0; g < a.n_hg; ++g) {
            const uint4 v = __ldg(p + (size_t)g * gs);
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[j]));
                s[2 * j] += f.x;
                s[2 * j + 1] += f.y;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) tile[rg * 8 + j][col] = s[j];
    sync128();
    if (b0 + tid < a.B) {
        const int64_t b = b0 + tid;
        const uint4 av = *reinterpret_cast<const uint4*>(a.aF + (size_t)b * 128 + k0);
        const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[j]));
            __nv_bfloat162 h2 = __floats2bfloat162_rn(tile[tid][2 * j] * act_grad_bf(a2.x, a.act), tile[tid][2 * j + 1] * act_grad_bf(a2.y, a.act));
            o[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(a.dpre + (size_t)b * 128 + k0) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// stand-alone launch (kept for A/B runs: NCDE_NO_FUSED_PREDUCE=1)
__global__ void __launch_bounds__(128) p_reduce_kernel(const __grid_constant__ PReduceArgs a) {
    __shared__ float tile[128][9];
    pdl_trigger();
    pdl_wait();
    p_reduce_slice(a, blockIdx.x, blockIdx.y, threadIdx.x, tile, [] { __syncthreads(); });
}

// ---------------------------------------------------------------------------------------------------------------
// tc_hidden_bwd: CTA = one 128-row batch tile; the input-gradient chain of the hidden layers, l = F-1 .. 0, with dpre_l
// (gradient w.r.t. the pre-activation of layer l) as a bf16 operand tile in shared memory:
//   dgrad   da_l[b][i] = dpre_l . W_l              (A = dpre_l K-major, B = W_l MN-major)        -> TMEM cols [0, 128)
//   epi     l > 0: dpre_{l-1} = da_l * act'(a_l) -> bf16 operand tile of the next level, also TMA-stored to the dpre record
//           l = 0: dz = da_0 -> fp32 feature-major dL/d(stage input); the RK adjoint combination happens where it is consumed
// The weight gradients are NOT formed here: dpre_l of every stage is kept as a record (p_reduce writes the top one) and one
// split-K kernel at the end of the backward pass contracts them with the saved activations (tc_hidden_wgrad), which takes
// the wgrad MMAs, the bias column sums and 32 k reductions per CTA off the sequential stage chain.
// Warp 8 is the producer (TMA + MMA).
// ---------------------------------------------------------------------------------------------------------------
struct TcHiddenBwdMaps {
    CUtensorMap W;                        // as in the forward kernel
    CUtensorMap act[kTcHidMaxLayers];     // act[l]: saved input of layer l
    CUtensorMap dpre;                     // {128, B, n_rec * F}: record (rec * F + l) = dpre_l of stage rec
};

struct TcHiddenBwdArgs {
    int B, Bp, H, F, rec;
    int act[kTcHidMaxLayers];             // activation of layer l (act[l-1] is the one a_l went through)
    float* dz_out;                        // [H][Bp] fp32: dL/d(stage input) of this stage
    // fused 64-way reduction (grid.y = 16 column blocks per batch tile; block y = 0 of each tile then runs the GEMM chain)
    // last backward stage of a step: instead of dz this launch writes the gradient of the step's start state,
    //   gy <- gy + dz + sum of the other stages' dz [+ gout_scale * grad_out of the output that sits at this time point]
    float* gy_io;                         // [H][Bp] or null
    const float* dz_other[NCDE_MAX_STAGES];
    int n_other;
    const float* gout;                    // (B, H) row-major or null
    float gout_scale;
    PReduceArgs pr;                       // pr.P == null: the top dpre record was written by a separate p_reduce launch
    int* tile_count;                      // per batch tile: column blocks finished so far (monotonic over the stages of a pass)
    int count_target;                     // 16 x (stages processed so far, this one included)
};

struct TcHidBwdSmem { uint32_t Wt, At, Dt, bars, total; };
__host__ __device__ inline TcHidBwdSmem tc_hid_bwd_layout() {
    TcHidBwdSmem L;
    uint32_t o = 0;
    L.Wt = o; o += 2 * kTcHidTile;
    L.At = o; o += 2 * kTcHidTile;
    L.Dt = o; o += 2 * kTcHidTile;
    L.bars = o; o += 16 * 8;
    L.total = o;
    return L;
}
static inline size_t tc_hid_bwd_smem_bytes() { return 1024 + tc_hid_bwd_layout().total; }

__global__ void __launch_bounds__(kTcThreads, 1) tc_hidden_bwd_kernel(const __grid_constant__ TcHiddenBwdArgs a,
                                                                      const __grid_constant__ TcHiddenBwdMaps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const TcHidBwdSmem L = tc_hid_bwd_layout();
    uint8_t* Wt = smem + L.Wt;     // weights of level l in buffer l & 1
    uint8_t* At = smem + L.At;     // a_l in buffer l & 1 (levels l > 0 only: act' needs it)
    uint8_t* Dt = smem + L.Dt;     // dpre_l in buffer l & 1
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* w_full = bars;          // [2]
    uint64_t* a_full = bars + 2;      // [2]
    uint64_t* d_full = bars + 4;      // TMA: dpre_{F-1}
    uint64_t* dg_bar = bars + 5;      // dgrad of the level complete
    uint64_t* dp_ready = bars + 6;    // epilogue -> producer: dpre_{l-1} tile written (8 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b0 = blockIdx.x * kTcM;
    const int F = a.F;
    const bool fused = a.pr.P != nullptr;
    const bool leader = blockIdx.y == 0;        // the one CTA per batch tile that runs the GEMM chain
    __shared__ float red_tile[128][9];

    if (leader) {
        if (warp == 0) tmem_alloc(tmem_slot, 128);
        if (tid == 0) {
            for (int i = 0; i < 2; ++i) { mbar_init(w_full + i, 1); mbar_init(a_full + i, 1); }
            mbar_init(d_full, 1); mbar_init(dg_bar, 1); mbar_init(dp_ready, 8);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = leader ? *tmem_slot : 0u;
    auto load_W = [&](int l) {
        uint8_t* dst = Wt + (size_t)(l & 1) * kTcHidTile;
        mbar_expect_tx(w_full + (l & 1), kTcHidTile);
        tma_load_3d(dst, &maps.W, w_full + (l & 1), 0, 0, l);
        tma_load_3d(dst + kTcHidTile / 2, &maps.W, w_full + (l & 1), 64, 0, l);
    };
    auto load_A = [&](int l) {   // saved records: written by the forward pass, older than every kernel of the backward pass
        if (l == 0) return;      // level 0 has no activation derivative to apply
        uint8_t* dst = At + (size_t)(l & 1) * kTcHidTile;
        mbar_expect_tx(a_full + (l & 1), kTcHidTile);
        tma_load_3d(dst, &maps.act[l], a_full + (l & 1), 0, b0, a.rec);
        tma_load_3d(dst + kTcHidTile / 2, &maps.act[l], a_full + (l & 1), 64, b0, a.rec);
    };
    const bool producer = leader && warp == 8 && lane == 0;
    if (producer) {
        tma_prefetch_desc(&maps.dpre);
        load_W(F - 1); load_A(F - 1);
        if (F > 1) { load_W(F - 2); load_A(F - 2); }
    }
    pdl_trigger();
    pdl_wait();   // the partials P come from the final-layer kernel (or dpre_{F-1} from a separate p_reduce launch)

    if (fused) {
        // phase 1, every CTA: one 128 x 8 slice of  dpre_{F-1} = (sum_g P_g) * act'(a_F)  -> the stage's dpre record, then a
        // release on the tile's counter; the leader's producer thread acquires it before its TMA reads the assembled tile
        if (tid < 128) {
            p_reduce_slice(a.pr, blockIdx.x, blockIdx.y, tid, red_tile, [] { named_bar_sync(2, 128); });
            __threadfence();
            named_bar_sync(2, 128);
            if (tid == 0) atomicAdd(a.tile_count + blockIdx.x, 1);
        }
        if (!leader) return;
    }

    if (warp == 8) {
        if (lane == 0) {
            if (fused) {
                int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(a.tile_count + blockIdx.x) : "memory");
                } while (seen < a.count_target);
                asm volatile("fence.proxy.async;" ::: "memory");   // the TMA (async proxy) read comes after the generic-proxy acquire
            }
            uint8_t* d_top = Dt + (size_t)((F - 1) & 1) * kTcHidTile;
            mbar_expect_tx(d_full, kTcHidTile);
            tma_load_3d(d_top, &maps.dpre, d_full, 0, b0, a.rec * F + F - 1);
            tma_load_3d(d_top + kTcHidTile / 2, &maps.dpre, d_full, 64, b0, a.rec * F + F - 1);
            uint32_t use_w[2] = {0, 0};
            for (int l = F - 1, n = 0; l >= 0; --l, ++n) {
                const int bf = l & 1;
                if (n == 0) mbar_wait(d_full, 0);
                else {
                    mbar_wait(dp_ready, (uint32_t)(n - 1) & 1u);   // dpre_l written (and fenced); every warp is done with a_{l+1}
                    // keep dpre_l for the weight-gradient kernel
                    tma_store_3d(&maps.dpre, Dt + (size_t)bf * kTcHidTile, 0, b0, a.rec * F + l);
                    tma_store_3d(&maps.dpre, Dt + (size_t)bf * kTcHidTile + kTcHidTile / 2, 64, b0, a.rec * F + l);
                    bulk_commit();
                    if (l - 1 >= 0) {
                        // the weight / activation buffers of level l+1 are free (its dgrad completed before its epilogue ran)
                        load_W(l - 1); load_A(l - 1);
                    }
                }
                mbar_wait(w_full + bf, use_w[bf] & 1u);
                ++use_w[bf];
                tc_fence_after();
                const uint32_t d_s = smem_u32(Dt + (size_t)bf * kTcHidTile), w_s = smem_u32(Wt + (size_t)bf * kTcHidTile);
                {   // dgrad: D[128 b x 128 i] = dpre_l (K-major over o) . W_l (MN-major: N = i contiguous, K = o rows)
                    const uint32_t idesc = make_idesc(kTcM, 128, 0, 1);
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t a_off = (uint32_t)(ks >> 2) * (uint32_t)kTcM * 128u + (uint32_t)(ks & 3) * 32u;
                        umma_bf16(tmem_base, make_sdesc(d_s + a_off, 16, 1024), make_sdesc(w_s + (uint32_t)ks * 2048u, 128u * 128u, 1024), idesc,
                                  ks > 0 ? 1u : 0u);
                    }
                }
                // the epilogue of this level overwrites the tile the store of level l+1's output... no: it writes buffer (l-1)&1,
                // whose previous content dpre_{l+1} was stored two levels ago: that store must be done reading shared memory
                bulk_wait_read<1>();
                umma_commit(dg_bar);
            }
            bulk_wait<0>();   // the records must be complete when the CTA exits
        }
    } else {
        const int wg = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const int64_t b = (int64_t)b0 + row;
        const bool row_ok = b < a.B;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t use_a[2] = {0, 0};
        for (int l = F - 1, n = 0; l >= 0; --l, ++n) {
            const int bf = l & 1;
            mbar_wait(dg_bar, (uint32_t)n & 1u);
            tc_fence_after();
            if (l > 0) {
                mbar_wait(a_full + bf, use_a[bf] & 1u);
                ++use_a[bf];
                const uint32_t a_s = smem_u32(At + (size_t)bf * kTcHidTile);
                const uint32_t dst = smem_u32(Dt + (size_t)((l - 1) & 1) * kTcHidTile);
                const int act = a.act[l - 1];
#pragma unroll
                for (int c0 = wg * 64; c0 < wg * 64 + 64; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32_issue(lane_addr + (uint32_t)c0, r);
                    uint32_t av[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(av[4 * q]), "=r"(av[4 * q + 1]), "=r"(av[4 * q + 2]),
                                     "=r"(av[4 * q + 3]) : "r"(a_s + sw128_off(row, (c0 >> 3) + q, kTcM)));
                    tmem_wait_ld<32>(r);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t pk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 av2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&av[4 * q + j]));
                            const float g0 = row_ok ? __uint_as_float(r[8 * q + 2 * j]) * act_grad_bf(av2.x, act) : 0.f;
                            const float g1 = row_ok ? __uint_as_float(r[8 * q + 2 * j + 1]) * act_grad_bf(av2.y, act) : 0.f;
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(g0, g1);
                            pk[j] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + sw128_off(row, (c0 >> 3) + q, kTcM)), "r"(pk[0]),
                                     "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                    }
                }
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dp_ready);
            } else {
                // dz, feature-major fp32: plain stores, coalesced over the lanes (= rows)
#pragma unroll
                for (int c0 = wg * 64; c0 < wg * 64 + 64; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32_issue(lane_addr + (uint32_t)c0, r);
                    tmem_wait_ld<32>(r);
                    if (row_ok) {
                        if (a.gy_io) {
                            // loads of 8 features are issued together, then the stores (a store between them would order every
                            // later load behind it: the compiler cannot rule out aliasing)
                            const float* d1 = a.n_other > 0 ? a.dz_other[0] : nullptr;
                            const float* d2 = a.n_other > 1 ? a.dz_other[1] : nullptr;
                            const float* d3 = a.n_other > 2 ? a.dz_other[2] : nullptr;
#pragma unroll
                            for (int j0 = 0; j0 < 32; j0 += 8) {
                                float v[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int h = c0 + j0 + j;
                                    const size_t off = (size_t)h * a.Bp + b;
                                    float gsum = 0.f;
                                    if (h < a.H) {
                                        gsum = __ldg(a.gy_io + off) + __uint_as_float(r[j0 + j]);
                                        if (d1) gsum += __ldg(d1 + off);
                                        if (d2) gsum += __ldg(d2 + off);
                                        if (d3) gsum += __ldg(d3 + off);
                                    }
                                    v[j] = gsum;
                                }
                                if (a.gout) {
                                    const float* gp = a.gout + (size_t)b * a.H + c0 + j0;   // 8 consecutive features of this row
                                    if (c0 + j0 + 8 <= a.H && (a.H & 3) == 0) {
                                        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp)), g1 = __ldg(reinterpret_cast<const float4*>(gp) + 1);
                                        v[0] = fmaf(a.gout_scale, g0.x, v[0]); v[1] = fmaf(a.gout_scale, g0.y, v[1]);
                                        v[2] = fmaf(a.gout_scale, g0.z, v[2]); v[3] = fmaf(a.gout_scale, g0.w, v[3]);
                                        v[4] = fmaf(a.gout_scale, g1.x, v[4]); v[5] = fmaf(a.gout_scale, g1.y, v[5]);
                                        v[6] = fmaf(a.gout_scale, g1.z, v[6]); v[7] = fmaf(a.gout_scale, g1.w, v[7]);
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 8; ++j)
                                            if (c0 + j0 + j < a.H) v[j] = fmaf(a.gout_scale, __ldg(gp + j), v[j]);
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (c0 + j0 + j < a.H) a.gy_io[(size_t)(c0 + j0 + j) * a.Bp + b] = v[j];
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c0 + j < a.H) a.dz_out[(size_t)(c0 + j) * a.Bp + b] = __uint_as_float(r[j]);
                        }
                    }
                }
                tc_fence_before();
            }
        }
    }
    tc_fence_before();
    __syncwarp();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// ---------------------------------------------------------------------------------------------------------------
// tc_hidden_wgrad: weight and bias gradients of the hidden layers for a WHOLE backward pass in one launch.
//   dW_l^T[i][o] = sum over (stage, batch tile) units of  a_l^T . dpre_l       (both operands MN-major, K = the 128 rows of a unit)
// CTA (l, split) owns a contiguous range of units, streams their two bf16 tiles through a 2-stage TMA pipeline,
// accumulates in ONE TMEM accumulator over the whole range and adds it to the global accumulator at the end (split-K);
// the bias gradient is the column sum of the dpre tiles, taken by the CUDA cores while the tensor core works.
// ---------------------------------------------------------------------------------------------------------------
struct TcHiddenWgradMaps {
    CUtensorMap act[kTcHidMaxLayers];     // act[l]: saved input of layer l, {128, B, n_rec}
    CUtensorMap dpre;                     // {128, B, n_rec * F}
};
struct TcHiddenWgradArgs {
    int F, n_rec, n_mt, n_split;
    float* dWacc[kTcHidMaxLayers];        // [128 out][128 in] fp32 accumulator of layer l's parameter slot
    float* dbacc[kTcHidMaxLayers];        // [128]
};
static inline size_t tc_hid_wgrad_smem_bytes() { return 1024 + 4 * kTcHidTile + 8 * 128 * 4 + 16 * 8; }

__global__ void __launch_bounds__(kTcThreads, 1) tc_hidden_wgrad_kernel(const __grid_constant__ TcHiddenWgradArgs a,
                                                                        const __grid_constant__ TcHiddenWgradMaps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* At = smem;                          // 2 x a_l tile
    uint8_t* Dt = smem + 2 * kTcHidTile;         // 2 x dpre_l tile
    float* bsum = reinterpret_cast<float*>(smem + 4 * kTcHidTile);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * kTcHidTile + 8 * 128 * 4);
    uint64_t* full = bars;            // [2] both tiles of a unit landed
    uint64_t* mma_done = bars + 2;    // [2] the unit's MMAs completed
    uint64_t* read_done = bars + 4;   // [2] the unit's column sums were taken (8 warp arrivals)
    uint64_t* fin_bar = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int l = blockIdx.x, sp = blockIdx.y;
    const int64_t U = (int64_t)a.n_rec * a.n_mt;
    const int64_t u_begin = sp * U / a.n_split, u_end = (sp + 1) * U / a.n_split;
    const int n = (int)(u_end - u_begin);

    if (warp == 0) tmem_alloc(tmem_slot, 128);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(full + i, 1); mbar_init(mma_done + i, 1); mbar_init(read_done + i, 8); }
        mbar_init(fin_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();   // the dpre records of the last stages come from the kernels just before this one

    if (warp == 8) {
        if (lane == 0 && n > 0) {
            auto load = [&](int j) {
                const int64_t u = u_begin + j;
                const int rec = (int)(u / a.n_mt), b0 = (int)(u % a.n_mt) * kTcM;
                const int bf = j & 1;
                mbar_expect_tx(full + bf, 2 * kTcHidTile);
                tma_load_3d(At + (size_t)bf * kTcHidTile, &maps.act[l], full + bf, 0, b0, rec);
                tma_load_3d(At + (size_t)bf * kTcHidTile + kTcHidTile / 2, &maps.act[l], full + bf, 64, b0, rec);
                tma_load_3d(Dt + (size_t)bf * kTcHidTile, &maps.dpre, full + bf, 0, b0, rec * a.F + l);
                tma_load_3d(Dt + (size_t)bf * kTcHidTile + kTcHidTile / 2, &maps.dpre, full + bf, 64, b0, rec * a.F + l);
            };
            load(0);
            if (n > 1) load(1);
            const uint32_t idesc = make_idesc(128, 128, 1, 1);
            for (int j = 0; j < n; ++j) {
                const int bf = j & 1;
                mbar_wait(full + bf, (uint32_t)(j >> 1) & 1u);
                tc_fence_after();
                const uint32_t a_s = smem_u32(At + (size_t)bf * kTcHidTile), d_s = smem_u32(Dt + (size_t)bf * kTcHidTile);
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t off = (uint32_t)ks * 2048u;
                    umma_bf16(tmem_base, make_sdesc(a_s + off, 128u * 128u, 1024), make_sdesc(d_s + off, 128u * 128u, 1024), idesc,
                              (ks > 0 || j > 0) ? 1u : 0u);
                }
                umma_commit(mma_done + bf);
                if (j + 2 < n) {
                    mbar_wait(mma_done + bf, (uint32_t)(j >> 1) & 1u);
                    mbar_wait(read_done + bf, (uint32_t)(j >> 1) & 1u);
                    load(j + 2);
                }
            }
            umma_commit(fin_bar);
        }
    } else {
        const int wg = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        float bacc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bacc[j] = 0.f;
        for (int j = 0; j < n; ++j) {
            const int bf = j & 1;
            mbar_wait(full + bf, (uint32_t)(j >> 1) & 1u);
            if (lane < 16) {   // 16 chunks of 8 columns x 8 row slices of 16 rows
                const uint32_t d_s = smem_u32(Dt + (size_t)bf * kTcHidTile);
#pragma unroll 4
                for (int r = warp * 16; r < warp * 16 + 16; ++r) {
                    uint32_t w4[4];
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w4[0]), "=r"(w4[1]), "=r"(w4[2]), "=r"(w4[3])
                                 : "r"(d_s + sw128_off(r, lane, kTcM)));
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[q]));
                        bacc[2 * q] += f.x;
                        bacc[2 * q + 1] += f.y;
                    }
                }
            }
            fence_async_smem();   // generic-proxy reads before the next TMA write into this buffer
            __syncwarp();
            if (lane == 0) mbar_arrive(read_done + bf);
        }
        if (n > 0) {
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) bsum[warp * 128 + lane * 8 + j] = bacc[j];
            }
            named_bar_sync(1, kTcEpiThreads);
            if (tid < 128 && a.dbacc[l]) {
                float sacc = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) sacc += bsum[w * 128 + tid];
                atomicAdd(a.dbacc[l] + tid, sacc);
            }
            mbar_wait(fin_bar, 0);
            tc_fence_after();
            float* acc = a.dWacc[l];
            const int i = row;   // TMEM lanes = input feature i, columns = output feature o
#pragma unroll
            for (int o0 = wg * 64; o0 < wg * 64 + 64; o0 += 32) {
                uint32_t r[32];
                tmem_ld32_issue(lane_addr + (uint32_t)o0, r);
                tmem_wait_ld<32>(r);
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(acc + (size_t)(o0 + j) * 128 + i, __uint_as_float(r[j]));
            }
        }
    }
    tc_fence_before();
    __syncwarp();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// gy += dz_0 + ... + dz_{n-1}: every stage input depends on y with unit Jacobian
__global__ void gy_accumulate_kernel(float* __restrict__ gyT, const float* dz0, const float* dz1, const float* dz2, const float* dz3,
                                     int n_dz, int64_t n) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = gyT[i] + dz0[i];
    if (n_dz > 1) g += dz1[i];
    if (n_dz > 2) g += dz2[i];
    if (n_dz > 3) g += dz3[i];
    gyT[i] = g;
}

__global__ void unpack_hidden_grad_kernel(const float* __restrict__ acc, const float* __restrict__ accb, float* __restrict__ gW,
                                          float* __restrict__ gb, int Dout, int Din) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < Dout * Din) gW[idx] += acc[(size_t)(idx / Din) * 128 + idx % Din];
    if (gb && idx < Dout) gb[idx] += accb[idx];
}

}  // namespace ncde
