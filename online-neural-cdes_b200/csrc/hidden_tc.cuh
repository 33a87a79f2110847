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

}  // namespace ncde
