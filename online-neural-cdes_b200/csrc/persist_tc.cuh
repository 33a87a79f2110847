// Persistent fixed-grid solve on the tensor cores: ONE kernel per pass (forward / backward) runs every RK stage of every step.
//
// A CDE solve is sequential in time but the series of a batch are independent, so nothing forces all rows through stage s
// before any row enters stage s+1.  The grid is one CTA per SM with two roles:
//   field CTA (g, p)   owns h-group g of the final layer for its whole life: the W3 slice stays in shared memory and (backward)
//                      the slice's weight gradient stays in TMEM across the pass (added to the global fp32 sum every kPsFlushUnits
//                      units); it streams the 128-row batch tiles p, p + n_part, ... through a warp-specialised tcgen05 pipeline,
//                      stage after stage.
//   hidden CTA j       owns batch tiles j, j + n_hid, ...: the hidden layers of the vector field (forward) / their input-gradient
//                      chain (backward) as 128x128x128 MMAs chained out of shared memory.
// The two roles hand tiles to each other through per-tile counters in global memory (release / acquire at gpu scope): a
// field CTA may start stage s+1 of tile t as soon as the hidden CTA has produced that tile's activations, while other tiles
// are still in stage s.  No kernel boundary, TMEM allocation, weight load or pipeline fill per stage.
//
// NSP = 1: bf16 operand tiles.  NSP = 2 ("bf16x3"): every MMA operand is a (hi, lo) pair of bf16 tiles, hi = bf16(x),
// lo = bf16(x - hi), and every GEMM is three MMAs hi*hi + lo*hi + hi*lo into one fp32 TMEM accumulator: 16 mantissa bits per
// operand instead of 8, with bf16's full exponent range (gradients need it; an fp16 split does not have it).
//
// Warps of a field CTA: 0-7 epilogue (thread = TMEM lane = batch row), 8 producer (TMA of the activation tile; tcgen05.mma of the
// forward GEMM / of the recompute and dgrad), 9 signaller (forward) / weight-gradient MMA issuer (backward), 10 dX/dt loader (TMA
// boxes into a ring of chunk slots, in the order the epilogue consumes them), 11 (backward) forms dL/dk one unit ahead and signals.
// Backward unit = one tile of one stage: recompute -> epilogue 1 (G, handed to the tensor core in 64-column blocks) -> dgrad, wgrad ->
// epilogue 2 (red.global.add.v4 into the tile's dA^T); what waits for what, and what was measured, is in DESIGN.md 4.1.
#pragma once
#include "hidden_tc.cuh"

namespace ncde {

constexpr int kPsThreads = 384;           // 12 warps, roles above (warps 9-11 idle in the hidden CTAs, warp 11 in the forward kernel)
constexpr int kPsEpi = 256;
constexpr int kPsFlushUnits = 512;        // backward field role: units between two flushes of the TMEM-resident weight gradient
constexpr long long kPsSpinLimit = 6000000000ll;   // ~3 s at 2 GHz: a protocol error traps instead of hanging the GPU

struct PsMaps {
    CUtensorMap W3;                              // {128 k, n_hg * Npad, NSP}            box {64, Npad, 1}
    CUtensorMap Wh;                              // {128 in, 128 out, NSP, F}            box {64, 128, 1, 1}
    CUtensorMap act[kTcHidMaxLayers + 1];        // act[l]: bf16 input of layer l, {128, B, NSP, n_rec}, box {64, 128, 1, 1}
    CUtensorMap dpre;                            // backward: {128, B, NSP, n_rec * F}
    CUtensorMap X;                               // dX/dt records fp32 {Cp, B, n_rec}, box {36, 128, 1}, no swizzle
};

struct PsArgs {
    int B, Bp, H, Cp, Hg, n_hg, Npad, F;
    int n_mt, n_part, n_field, n_hid;
    int NS, n_steps, method, need_grad, NA, NX;    // NA: activation buffers of the forward field role; NX: dX/dt chunk slots
    int act[kTcHidMaxLayers];
    const float* dt;            // device [n_steps]
    const int* emit_idx;        // device [n_steps]: output slot that receives the state at the end of step s, or -1
    float* z_out;               // (n_out, B, H) row-major
    const float* grad_out;      // backward: (n_out, B, H)
    float* yT[2];               // forward: state ping-pong [H][Bp];  backward: yT[0] = gy
    float* kT[NCDE_MAX_STAGES]; // forward: stage derivatives;  backward: dz of the stages of the current step
    const float* b3;            // [n_hg][Npad]
    const float* bias_h;        // [F][128]
    __nv_bfloat16* rec0;        // saved records: record r at rec0 + r * rec_stride (bf16 elements); layer l input at + act_off[l]
    size_t rec_stride;          // bf16 elements between stage records (forward without gradient: one scratch record, index 0)
    size_t act_off[kTcHidMaxLayers + 1];   // bf16-element offset of layer l's input inside a record; part p at + p * Bp * 128
    const float* dx0;           // dX/dt records [rec][Bp][Cp] fp32
    size_t dx_stride;           // floats between records
    // backward
    float* dAT;                 // [n_mt][128 k][128 rows] fp32: sum over the h-groups of dL/d(final-layer input) of the tile in flight
    float* dW3acc;              // [n_part][n_hg * Npad][128]
    float* db3acc;              // [n_part][n_hg * Npad]
    __nv_bfloat16* dpre0;       // dpre records [rec * F + l][NSP][Bp][128]
    // synchronisation words (zeroed before the launch)
    int* cnt_f;                 // [n_mt] field -> hidden: arrivals of field CTAs (monotonic)
    int* flag_h;                // [n_mt] hidden -> field: stages completed by the hidden CTA (monotonic)
    unsigned long long* trace;  // debug (NCDE_PS_TRACE): [stage][tile < 8][24] globaltimer stamps of the hand-offs of h-group trace_g, or null
    int trace_g;
};
constexpr int kPsTraceStages = 48, kPsTraceTiles = 8, kPsTraceEv = 40;
// second trace area (after the first): stamps of EVERY h-group for one (stage, tile): [g < 256][8]
constexpr int kPsTraceAllStage = 40, kPsTraceAllTile = 0;
__device__ __forceinline__ void ps_trace_all(const PsArgs& a, int t, int g, int q, int ev) {
    if (a.trace && q == kPsTraceAllStage && t == kPsTraceAllTile && g >= 0 && g < 256) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
        a.trace[kPsTraceStages * kPsTraceTiles * kPsTraceEv + g * 8 + ev] = ns;
        if (ev == 1) {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            a.trace[kPsTraceStages * kPsTraceTiles * kPsTraceEv + g * 8 + 7] = smid + 1;
        }
    }
}
__device__ __forceinline__ void ps_trace(const PsArgs& a, int t, int g, int q, int ev) {
    if (a.trace && t < kPsTraceTiles && (g == a.trace_g || g < 0) && q < kPsTraceStages) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
        a.trace[(q * kPsTraceTiles + t) * kPsTraceEv + ev] = ns;
    }
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
#ifdef NCDE_PS_DBG_PRINT
#include <cassert>
__device__ uint32_t g_ps_dbg_bars;      // shared-memory address of the barrier array of the (last initialised) field role
#define PS_TRAP(what, v0_, v1_) do { printf("PS TIMEOUT %s blk %d tid %d : %d %d\n", what, (int)blockIdx.x, (int)threadIdx.x, (int)(v0_), (int)(v1_)); assert(0); } while (0)
#else
#define PS_TRAP(what, v0_, v1_) __trap()
#endif
__device__ __forceinline__ void ps_spin_ge(const int* p, int target) {
    if (ld_acquire_gpu(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu(p) < target) {
        __nanosleep(20);
        if (clock64() - t0 > kPsSpinLimit) PS_TRAP("spin", target, ld_acquire_gpu(p));
    }
}
// bounded mbarrier wait (same reason: trap, never hang)
__device__ __forceinline__ void ps_wait(uint64_t* bar, uint32_t parity) {
    uint32_t n = 0;
    while (!mbar_try_wait(bar, parity)) {
#ifdef NCDE_PS_DBG_PRINT
        if (++n > 4000000u) PS_TRAP("mbar", ((int)smem_u32(bar) - (int)g_ps_dbg_bars) / 8, parity);
#else
        if (++n > 40000000u) __trap();
#endif
    }
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ float4 ldg128_nc(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// the n floats at p (one batch row of dX/dt, 16-byte aligned) -> L1, one request per 128-byte line
__device__ __forceinline__ void ps_prefetch_row(const float* p, int n) {
    const char* c = reinterpret_cast<const char*>(p);
    const char* e = c + (size_t)n * 4;
    for (const char* q = reinterpret_cast<const char*>((uintptr_t)c & ~(uintptr_t)127); q < e; q += 128)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
}
// (hi, lo) bf16 split of two floats, packed
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h2);
}
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h2);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = pack_bf16x2(a - hf.x, b - hf.y);
}
template <bool EXACT>
__device__ __forceinline__ float ps_tanh(float x) {
    if (EXACT) { float t, q; tanh_sech2(x, t, q); return t; }
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool EXACT>
__device__ __forceinline__ float ps_sech2(float x) {
    if (EXACT) { float t, q; tanh_sech2(x, t, q); return q; }
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return fmaf(-y, y, 1.f);
}

// The producer warp can run CONVERGED (NCDE_PS_CONVERGED=1 at build time): all 32 lanes execute the loops and the waits, and each
// single-thread instruction (tcgen05.mma, tcgen05.commit, TMA, mbarrier arrive) is issued under elect.sync, as cute's atoms do; that
// removes the ELECT / R2UR / branch waterfall ptxas wraps around a tcgen05.mma issued from a divergent `if (lane == 0)` region.
// Measured on cfg 5 it does NOT pay: MMA issue is back-pressured by MMA execution either way (24 MMAs of M128 N112 K16 take
// 0.86 us in both forms: ~70 cycles each, the shared-memory operand bandwidth), and the converged form was 5 % slower overall
// (bf16x3 backward 44.8 vs 42.2 ms).  Default: one lane runs the producer loop.
#ifndef NCDE_PS_CONVERGED
#define NCDE_PS_CONVERGED 0
#endif
constexpr bool kPsConverged = NCDE_PS_CONVERGED != 0;
__device__ __forceinline__ bool ps_elect() {
    if (!kPsConverged) return true;
    uint32_t p;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void ps_syncwarp() { if (kPsConverged) __syncwarp(); }
#define PS_LEAD(...) do { if (ps_elect()) { __VA_ARGS__; } } while (0)
__device__ __forceinline__ void ps_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (ps_elect()) umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
}
// MMA issue is done by ONE thread: every instruction it spends per tcgen05.mma is serial latency of the whole CTA (bf16x3 issues 45
// MMAs per backward unit).  Descriptors are therefore built once per operand and advanced by adding the 16-byte-granular offset to
// their address field (shared-memory addresses stay below 2^18, so the 14-bit field cannot carry), inside fully unrolled loops.
__device__ __forceinline__ uint64_t ps_desc_advance(uint64_t desc, uint32_t byte_off) { return desc + (uint64_t)(byte_off >> 4); }
// D[128 x N] (+)= A[128 x 128] . B[N x 128]^T, both operands K-major swizzled tiles of two 64-column blocks; `first` clears D
__device__ __forceinline__ void ps_issue_kmajor(uint32_t d_tmem, uint32_t a_saddr, uint32_t b_saddr, int b_rows, int N, bool first) {
    const uint32_t idesc = make_idesc(kTcM, N, 0, 0);
    const uint64_t ad = make_sdesc(a_saddr, 16, 1024), bd = make_sdesc(b_saddr, 16, 1024);
    const uint32_t b_blk = (uint32_t)b_rows * 128u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t ao = (uint32_t)(k >> 2) * (uint32_t)kTcM * 128u + (uint32_t)(k & 3) * 32u;
        const uint32_t bo = (uint32_t)(k >> 2) * b_blk + (uint32_t)(k & 3) * 32u;
        ps_umma(d_tmem, ps_desc_advance(ad, ao), ps_desc_advance(bd, bo), idesc, (k > 0 || !first) ? 1u : 0u);
    }
}
// D[128 x N] (+)= A . B^T for NSP-part operands, both K-major swizzled tiles [rows][128] (two 64-column blocks): hi*hi (+ lo*hi + hi*lo)
template <int NSP, bool ROLLED = false>
__device__ __forceinline__ void ps_gemm_kmajor(uint32_t d_tmem, uint32_t a_s, uint32_t a_part_bytes, int a_rows, uint32_t b_s,
                                               uint32_t b_part_bytes, int b_rows, int N) {
    (void)a_rows;
    if (ROLLED) {
        // backward field role: the operand pairs are a rolled loop — the issuing thread is back-pressured by the tensor pipe anyway, and
        // every unrolled tcgen05.mma site costs ~13 instructions of an instruction cache the epilogue warps need (measured: -5 %)
#pragma unroll 1
        for (int pr = 0; pr < (NSP == 2 ? 3 : 1); ++pr)
            ps_issue_kmajor(d_tmem, a_s + (pr == 1 ? a_part_bytes : 0u), b_s + (pr == 2 ? b_part_bytes : 0u), b_rows, N, pr == 0);
        return;
    }
    ps_issue_kmajor(d_tmem, a_s, b_s, b_rows, N, true);
    if (NSP == 2) {
        ps_issue_kmajor(d_tmem, a_s + a_part_bytes, b_s, b_rows, N, false);
        ps_issue_kmajor(d_tmem, a_s, b_s + b_part_bytes, b_rows, N, false);
    }
}

// dX/dt reaches the epilogue threads through a ring of chunk slots in shared memory: slot = [128 rows][36 floats] (32 channels + 4 of
// padding: a 144-byte pitch keeps the per-row LDS.128 of consecutive lanes on distinct banks), filled by the loader warp with one TMA
// box per chunk, in exactly the order the epilogue consumes them (unit after unit, h after h, chunk after chunk).
constexpr int kPsXPitch = 36;
constexpr uint32_t kPsXSlot = kTcM * kPsXPitch * 4;     // 18432 bytes
constexpr int kPsMaxSlots = 8;
// Backward field role with ONE hidden row per h-group (bf16x3 on wide fields: shared memory leaves room for two wide slots only, and a
// ring without depth exposes the TMA latency of every refill): 16-channel chunks, pitch 20 floats (80 bytes: still conflict-free for the
// per-row LDS.128), chunk j worked on by warp group j & 1 — the two groups walk the ring together, so a slot is free as soon as its chunk
// is done, and twice as many slots fit.
constexpr int kPsXPitchN = 20;
constexpr uint32_t kPsXSlotN = kTcM * kPsXPitchN * 4;   // 10240 bytes
__host__ __device__ inline bool ps_bwd_narrow(int Hg) { return Hg == 1; }
__host__ __device__ inline int ps_bwd_xw(int Hg) { return ps_bwd_narrow(Hg) ? 16 : 32; }
__host__ __device__ inline uint32_t ps_bwd_xslot(int Hg) { return ps_bwd_narrow(Hg) ? kPsXSlotN : kPsXSlot; }

// 16 accumulator columns of TMEM lane `row` in flight
struct PsHalf { uint32_t r[16]; };
__device__ __forceinline__ void ps_half_issue(PsHalf& k, uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(k.r[0]), "=r"(k.r[1]), "=r"(k.r[2]), "=r"(k.r[3]), "=r"(k.r[4]), "=r"(k.r[5]), "=r"(k.r[6]), "=r"(k.r[7]),
                   "=r"(k.r[8]), "=r"(k.r[9]), "=r"(k.r[10]), "=r"(k.r[11]), "=r"(k.r[12]), "=r"(k.r[13]), "=r"(k.r[14]), "=r"(k.r[15])
                 : "r"(taddr) : "memory");
}
// sum_j tanh(D[row][j] + b3[j]) * dX[row][j] over the first nv (0, 8 or 16) columns of the half
template <bool EXACT>
__device__ __forceinline__ float ps_fwd_half(const PsHalf& k, int nv, uint32_t b3_s, uint32_t x_s) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    if (nv == 16) {   // the common case, branch-free: every load first, then the arithmetic
        float4 bb[4], dd[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { bb[q] = lds128(b3_s + 16u * q); dd[q] = lds128(x_s + 16u * q); }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            acc0 = fmaf(ps_tanh<EXACT>(__uint_as_float(k.r[4 * q + 0]) + bb[q].x), dd[q].x, acc0);
            acc1 = fmaf(ps_tanh<EXACT>(__uint_as_float(k.r[4 * q + 1]) + bb[q].y), dd[q].y, acc1);
            acc2 = fmaf(ps_tanh<EXACT>(__uint_as_float(k.r[4 * q + 2]) + bb[q].z), dd[q].z, acc2);
            acc3 = fmaf(ps_tanh<EXACT>(__uint_as_float(k.r[4 * q + 3]) + bb[q].w), dd[q].w, acc3);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (4 * q < nv) {
                const float4 bb = lds128(b3_s + 16u * q), dd = lds128(x_s + 16u * q);
                acc0 = fmaf(ps_tanh<EXACT>(__uint_as_float(k.r[4 * q + 0]) + bb.x), dd.x, acc0);
                acc1 = fmaf(ps_tanh<EXACT>(__uint_as_float(k.r[4 * q + 1]) + bb.y), dd.y, acc1);
                acc2 = fmaf(ps_tanh<EXACT>(__uint_as_float(k.r[4 * q + 2]) + bb.z), dd.z, acc2);
                acc3 = fmaf(ps_tanh<EXACT>(__uint_as_float(k.r[4 * q + 3]) + bb.w), dd.w, acc3);
            }
        }
    }
    return (acc0 + acc1) + (acc2 + acc3);
}
// the chunk sequence of one unit, identical for the loader and every epilogue warp: n_pass passes over nch chunks
struct PsXSeq {
    int n_pass, nch;
};
__device__ __forceinline__ PsXSeq ps_xseq(int Hg, int Cp) {
    PsXSeq q;
    q.n_pass = Hg >= 2 ? Hg - Hg / 2 : 1;      // hidden rows per warp-group (the larger share when Hg is odd)
    q.nch = (Cp + 31) / 32;
    return q;
}

// unit i of a field CTA -> (global stage index q, batch tile t)
struct PsUnit { int q, t; };
__device__ __forceinline__ PsUnit ps_unit_fwd(int i, int n_my, int part, int n_part) {
    PsUnit u;
    u.q = i / n_my;
    u.t = part + (i - u.q * n_my) * n_part;
    return u;
}

// ---------------------------------------------------------------------------------------------------------------
// shared-memory layouts (byte offsets from the 1024-aligned base)
// ---------------------------------------------------------------------------------------------------------------
struct PsFieldFwdSmem { uint32_t Ws, As, Xs, b3s, part, bars, total; };
__host__ __device__ inline PsFieldFwdSmem ps_field_fwd_layout(int Npad, int NSP, int NA, int NX) {
    PsFieldFwdSmem L;
    uint32_t o = 0;
    L.Ws = o; o += (uint32_t)NSP * (uint32_t)Npad * 256u;
    o = (o + 1023u) & ~1023u;
    L.As = o; o += (uint32_t)NA * (uint32_t)NSP * kTcHidTile;
    L.Xs = o; o += (uint32_t)NX * kPsXSlot;
    L.b3s = o; o += (uint32_t)Npad * 4;
    L.part = o; o += kTcM * 4;
    o = (o + 15u) & ~15u;
    L.bars = o; o += 40 * 8;
    L.total = o;
    return L;
}
struct PsHidFwdSmem { uint32_t Wt, At, bias, bars, total; };
__host__ __device__ inline PsHidFwdSmem ps_hid_fwd_layout(int NSP, int NW, int n_op = 1) {
    PsHidFwdSmem L;
    uint32_t o = 0;
    L.Wt = o; o += (uint32_t)NW * (uint32_t)NSP * kTcHidTile;
    L.At = o; o += (uint32_t)n_op * (uint32_t)NSP * kTcHidTile;
    L.bias = o; o += kTcHidMaxLayers * 128 * 4;
    L.bars = o; o += 24 * 8;
    L.total = o;
    return L;
}
// weight buffers of the hidden role: all F layers resident when they fit next to the activation tiles, else a ring
__host__ __device__ inline int ps_hid_nw(int NSP, int F, bool* resident, int n_op = 1) {
    const uint32_t budget = 220u * 1024u;
    const uint32_t fixed = (uint32_t)n_op * NSP * kTcHidTile + kTcHidMaxLayers * 512 + 1024 + 24 * 8;
    const int fit = (int)((budget - fixed) / ((uint32_t)NSP * kTcHidTile));
    if (fit >= F) { *resident = true; return F; }
    *resident = false;
    return fit >= 2 ? 2 : 1;
}

// operand tiles of the forward hidden role: two (no wait between a record store and the next epilogue) when the resident weights
// leave room, else one (the epilogue overwrites the tile the MMA has finished reading, after the record store has read it)
__host__ __device__ inline int ps_hid_fwd_nop(int NSP, int F) {
    bool res;
    ps_hid_nw(NSP, F, &res, 2);
    return res ? 2 : 1;
}

// ---------------------------------------------------------------------------------------------------------------
// forward, field role
// ---------------------------------------------------------------------------------------------------------------
// The RK state this thread's (row, h) entry needs when its k is finished: read at the START of the unit (the values were written
// by this same thread in earlier units), so that the L2 latency hides behind the accumulator read-out instead of sitting between
// the last chunk and the hand-off.
struct PsFwdState { float y, k0, k1, k2; };
__device__ __forceinline__ PsFwdState ps_fwd_state_load(const PsArgs& a, int s, int ist, int h, int64_t b) {
    PsFwdState v;
    const size_t off = (size_t)h * a.Bp + b;
    v.y = a.yT[s & 1][off];
    v.k0 = ist >= 1 ? a.kT[0][off] : 0.f;
    v.k1 = ist >= 2 ? a.kT[1][off] : 0.f;
    v.k2 = ist >= 3 ? a.kT[2][off] : 0.f;
    return v;
}
// what the forward epilogue does with one finished k[b, h] of global stage q (cf. tc_fwd_finish_element; reference operation
// order of rk_common.py:106-114, no FMA contraction)
template <int NSP>
__device__ __forceinline__ void ps_fwd_finish_element(const PsArgs& a, const PsFwdState& v, int q, int s, int ist, float dt, int emit_slot,
                                                      int h, int64_t b, float k) {
    const size_t off = (size_t)h * a.Bp + b;
    const int cur = s & 1;
    const float third = 0.3333333432674408f;
    a.kT[ist][off] = k;
    const size_t part_stride = (size_t)a.Bp * 128;
    float znext;
    bool have_next;
    if (ist + 1 < a.NS) {
        // input of the next stage of this step: stage i+1 reads k_1 .. k_{i+1}, of which k_{i+1} is the value just finished
        float inc;
        if (ist == 0) inc = __fmul_rn(__fmul_rn(dt, k), third);                                   // dt * k1 * (1/3)
        else if (ist == 1) inc = __fmul_rn(dt, __fsub_rn(k, __fmul_rn(v.k0, third)));             // dt * (k2 - k1 * (1/3))
        else inc = __fmul_rn(dt, __fadd_rn(__fsub_rn(v.k0, v.k1), k));                            // dt * (k1 - k2 + k3)
        znext = __fadd_rn(v.y, inc);
        have_next = true;
    } else {
        float yn;
        if (a.method == NCDE_RK4_38) {   // (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
            float sm = __fadd_rn(v.k0, __fmul_rn(3.f, __fadd_rn(v.k1, v.k2)));
            sm = __fadd_rn(sm, k);
            yn = __fadd_rn(v.y, __fmul_rn(__fmul_rn(sm, dt), 0.125f));
        } else {
            yn = __fadd_rn(v.y, __fmul_rn(dt, k));
        }
        a.yT[cur ^ 1][off] = yn;
        if (emit_slot >= 0) a.z_out[((size_t)emit_slot * a.B + b) * a.H + h] = yn;
        znext = yn;
        have_next = s + 1 < a.n_steps;
    }
    if (have_next) {
        __nv_bfloat16* zr = a.rec0 + (a.need_grad ? (size_t)(q + 1) * a.rec_stride : 0) + a.act_off[0] + (size_t)b * 128 + h;
        const __nv_bfloat16 hi = __float2bfloat16(znext);
        zr[0] = hi;
        if (NSP == 2) zr[part_stride] = __float2bfloat16(znext - __bfloat162float(hi));
    }
}

template <int NSP>
__device__ void ps_field_fwd(const PsArgs& a, const PsMaps& maps, uint8_t* smem, int g, int part) {
    constexpr bool EXACT = NSP == 2;
    const int Npad = a.Npad, NA = a.NA, NX = a.NX;
    const PsFieldFwdSmem L = ps_field_fwd_layout(Npad, NSP, NA, NX);
    const uint32_t w_part = (uint32_t)Npad * 256u;
    uint8_t* Ws = smem + L.Ws;
    uint8_t* As = smem + L.As;
    uint8_t* Xs = smem + L.Xs;
    float* b3s = reinterpret_cast<float*>(smem + L.b3s);
    float* partial = reinterpret_cast<float*>(smem + L.part);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* full_a = bars;          // [2]
    uint64_t* mma_bar = bars + 2;     // [2]
    uint64_t* done = bars + 4;        // [2]
    uint64_t* w_bar = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
    volatile int* sig_done = reinterpret_cast<volatile int*>(bars + 8);
    uint64_t* x_full = bars + 10;     // [NX <= 8]
    uint64_t* x_free = bars + 18;     // [NX <= 8]

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform
    const uint32_t acc_stride = tc_tmem_cols(Npad);
    const int n_my = part < a.n_mt ? (a.n_mt - part + a.n_part - 1) / a.n_part : 0;
    const int n_q = a.n_steps * a.NS;
    const int n_units = n_my * n_q;
    const PsXSeq xq = ps_xseq(a.Hg, a.Cp);

    if (warp == 0) tmem_alloc(tmem_slot, 2 * acc_stride);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(full_a + i, 1); mbar_init(mma_bar + i, 1); mbar_init(done + i, 8); }
        for (int i = 0; i < kPsMaxSlots; ++i) { mbar_init(x_full + i, 1); mbar_init(x_free + i, 8); }
        mbar_init(w_bar, 1);
        *sig_done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < Npad; i += kPsThreads) b3s[i] = a.b3[(size_t)g * Npad + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if ((kPsConverged || lane == 0) && n_units > 0) {     // see ps_elect
            if (ps_elect()) {
                tma_prefetch_desc(&maps.act[a.F]);
                mbar_expect_tx(w_bar, (uint32_t)NSP * w_part);
                for (int p = 0; p < NSP; ++p) {
                    tma_load_3d(Ws + (size_t)p * w_part, &maps.W3, w_bar, 0, g * Npad, p);
                    tma_load_3d(Ws + (size_t)p * w_part + (size_t)Npad * 128, &maps.W3, w_bar, 64, g * Npad, p);
                }
            }
            ps_syncwarp();
            for (int i = 0; i < n_units; ++i) {
                const PsUnit u = ps_unit_fwd(i, n_my, part, a.n_part);
                const int b = i & 1, ba = NA == 2 ? b : 0;
                if (i >= NA) ps_wait(mma_bar + ((i - NA) & 1), (uint32_t)((i - NA) >> 1) & 1u);   // activation buffer free
                if (lane == 0) {
                    ps_spin_ge(a.flag_h + u.t, u.q + 1);        // the hidden CTA has written the tile's final-layer input of stage q
                    ps_trace(a, u.t, g, u.q, 0);
                }
                ps_syncwarp();
                fence_proxy_async_all();
                uint8_t* dst = As + (size_t)ba * NSP * kTcHidTile;
                const int rec = a.need_grad ? u.q : 0;
                if (ps_elect()) {
                    mbar_expect_tx(full_a + ba, (uint32_t)NSP * kTcHidTile);
                    for (int p = 0; p < NSP; ++p) {
                        tma_load_4d(dst + (size_t)p * kTcHidTile, &maps.act[a.F], full_a + ba, 0, u.t * kTcM, p, rec);
                        tma_load_4d(dst + (size_t)p * kTcHidTile + kTcHidTile / 2, &maps.act[a.F], full_a + ba, 64, u.t * kTcM, p, rec);
                    }
                }
                ps_syncwarp();
                if (i >= 2) {
                    ps_wait(done + b, (uint32_t)((i - 2) >> 1) & 1u);        // accumulator b drained by the epilogue of unit i-2
                    while (*sig_done < i - 1) {}                              // ... and that unit was signalled (keeps the signaller in phase)
                }
                if (i == 0) ps_wait(w_bar, 0);
                ps_wait(full_a + ba, (uint32_t)(i / NA) & 1u);
                if (lane == 0) ps_trace(a, u.t, g, u.q, 1);
                tc_fence_after();
                ps_gemm_kmajor<NSP>(tmem_base + (uint32_t)b * acc_stride, smem_u32(dst), kTcHidTile, kTcM, smem_u32(Ws), w_part, Npad, Npad);
                PS_LEAD(umma_commit(mma_bar + b));
                ps_syncwarp();
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            for (int i = 0; i < n_units; ++i) {
                const PsUnit u = ps_unit_fwd(i, n_my, part, a.n_part);
                ps_wait(done + (i & 1), (uint32_t)(i >> 1) & 1u);
                red_release_gpu_add(a.cnt_f + u.t, 1);      // release: cumulative over the epilogue warps' writes (ordered by the mbarrier)
                ps_trace(a, u.t, g, u.q, 4);
                *sig_done = i + 1;
            }
        }
    } else if (warp == 10) {
        // dX/dt loader: the records were written before the launch, so it simply runs NX chunks ahead of the epilogue
        if (lane == 0) {
            tma_prefetch_desc(&maps.X);
            int slot = 0;
            uint32_t lap = 0;          // completed laps of the ring
            for (int i = 0; i < n_units; ++i) {
                const PsUnit u = ps_unit_fwd(i, n_my, part, a.n_part);
                for (int p = 0; p < xq.n_pass; ++p)
                    for (int j = 0; j < xq.nch; ++j) {
                        if (lap > 0) ps_wait(x_free + slot, (lap - 1) & 1u);
                        mbar_expect_tx(x_full + slot, kPsXSlot);
                        tma_load_3d(Xs + (size_t)slot * kPsXSlot, &maps.X, x_full + slot, 32 * j, u.t * kTcM, u.q);
                        if (++slot == NX) { slot = 0; ++lap; }
                    }
            }
        }
    } else if (warp < 8) {
        const int wg = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const int h_begin = a.Hg >= 2 ? (wg == 0 ? 0 : a.Hg / 2) : 0;
        const int h_end = a.Hg >= 2 ? (wg == 0 ? a.Hg / 2 : a.Hg) : 1;
        const uint32_t b3_s = smem_u32(b3s);
        const uint32_t xs_row = smem_u32(Xs) + (uint32_t)row * (kPsXPitch * 4);
        int slot = 0;
        uint32_t lap = 0;
        for (int i = 0; i < n_units; ++i) {
            const PsUnit u = ps_unit_fwd(i, n_my, part, a.n_part);
            const int b = i & 1;
            const int s = u.q / a.NS, ist = u.q - s * a.NS;
            const float dt = __ldg(a.dt + s);
            const int emit_slot = __ldg(a.emit_idx + s);
            const int64_t b0 = (int64_t)u.t * kTcM;
            const bool row_ok = b0 + row < a.B;
            ps_wait(mma_bar + b, (uint32_t)(i >> 1) & 1u);
            if (tid == 0) ps_trace(a, u.t, g, u.q, 2);
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + (uint32_t)b * acc_stride + ((uint32_t)((warp & 3) * 32) << 16);
            // my chunks of the unit's sequence: Hg >= 2 -> every chunk of my own hidden rows; Hg == 1 -> the chunks of my parity
            PsHalf HA, HB;
            {   // pre-issue the first half of my first chunk
                const int j0 = a.Hg >= 2 ? 0 : wg;
                if (h_begin < h_end && j0 < xq.nch) ps_half_issue(HA, lane_addr + (uint32_t)(h_begin * a.Cp + 32 * j0));
            }
            for (int p = 0; p < xq.n_pass; ++p) {
                const int hl = h_begin + p;
                const bool pass_ok = hl < h_end;
                const int colbase = hl * a.Cp;
                const int h_out = a.Hg >= 2 ? g * a.Hg + hl : g;
                const bool writes = row_ok && h_out < a.H && (a.Hg >= 2 ? pass_ok : wg == 1);
                PsFwdState sv = {0.f, 0.f, 0.f, 0.f};
                if (writes) sv = ps_fwd_state_load(a, s, ist, h_out, b0 + row);
                float acc = 0.f;
                for (int j = 0; j < xq.nch; ++j) {
                    ps_wait(x_full + slot, lap & 1u);
                    const bool mine = pass_ok && (a.Hg >= 2 || (j & 1) == wg);
                    if (mine) {
                        const int c = 32 * j;
                        const int nv0 = min(16, a.Cp - c), nv1 = max(0, min(16, a.Cp - c - 16));
                        const uint32_t xs = xs_row + (uint32_t)slot * kPsXSlot;
                        tmem_wait_ld<16>(HA.r);
                        if (nv1 > 0) ps_half_issue(HB, lane_addr + (uint32_t)(colbase + c + 16));
                        acc += ps_fwd_half<EXACT>(HA, nv0, b3_s + 4u * (colbase + c), xs);
                        tmem_wait_ld<16>(HB.r);
                        {   // first half of my next chunk in this unit
                            int pn = p, jn = j + (a.Hg >= 2 ? 1 : 2);
                            if (jn >= xq.nch) { pn = p + 1; jn = a.Hg >= 2 ? 0 : wg; }
                            if (pn < xq.n_pass && h_begin + pn < h_end && jn < xq.nch)
                                ps_half_issue(HA, lane_addr + (uint32_t)((h_begin + pn) * a.Cp + 32 * jn));
                        }
                        acc += ps_fwd_half<EXACT>(HB, nv1, b3_s + 4u * (colbase + c + 16), xs + 64u);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(x_free + slot);
                    if (++slot == NX) { slot = 0; ++lap; }
                }
                if (a.Hg >= 2) {
                    if (writes) ps_fwd_finish_element<NSP>(a, sv, u.q, s, ist, dt, emit_slot, h_out, b0 + row, acc);
                } else {
                    if (wg == 0) partial[row] = acc;
                    named_bar_sync(1, kPsEpi);
                    if (writes) ps_fwd_finish_element<NSP>(a, sv, u.q, s, ist, dt, emit_slot, h_out, b0 + row, acc + partial[row]);
                    named_bar_sync(1, kPsEpi);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (tid == 0) ps_trace(a, u.t, g, u.q, 3);
            if (lane == 0) mbar_arrive(done + b);
        }
    }
    tc_fence_before();
    __syncwarp();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 2 * acc_stride);
}

// ---------------------------------------------------------------------------------------------------------------
// forward, hidden role: z_q tile -> hidden layers -> input of the final layer; every layer input is kept as a record.
// ONE operand buffer: the epilogue of layer l overwrites the tile MMA l has finished reading with the input of layer l+1.
// ---------------------------------------------------------------------------------------------------------------
template <int NSP>
__device__ void ps_hidden_fwd(const PsArgs& a, const PsMaps& maps, uint8_t* smem, int j) {
    bool resident;
    const int n_op = ps_hid_fwd_nop(NSP, a.F);
    const int NW = ps_hid_nw(NSP, a.F, &resident, n_op);
    const PsHidFwdSmem L = ps_hid_fwd_layout(NSP, NW, n_op);
    constexpr uint32_t kOp = (uint32_t)NSP * kTcHidTile;     // one operand (all parts)
    uint8_t* Wt = smem + L.Wt;
    uint8_t* At = smem + L.At;
    float* bias_s = reinterpret_cast<float*>(smem + L.bias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* w_full = bars;          // [NW <= 8]
    uint64_t* a_full = bars + 8;
    uint64_t* mma_bar = bars + 9;
    uint64_t* a_ready = bars + 10;    // epilogue -> producer: output of the layer written over the operand tile (8 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform
    const int F = a.F;
    const int n_my = j < a.n_mt ? (a.n_mt - j + a.n_hid - 1) / a.n_hid : 0;
    const int n_q = a.n_steps * a.NS;
    const int n_units = n_my * n_q;

    if (warp == 0) tmem_alloc(tmem_slot, 128);
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(w_full + i, 1);
        mbar_init(a_full, 1); mbar_init(mma_bar, 1); mbar_init(a_ready, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < F * 128; i += kPsThreads) bias_s[i] = a.bias_h[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if ((kPsConverged || lane == 0) && n_units > 0) {     // see ps_elect
            auto load_W = [&](int l, int buf) {
                uint8_t* dst = Wt + (size_t)buf * kOp;
                if (ps_elect()) {
                    mbar_expect_tx(w_full + buf, kOp);
                    for (int p = 0; p < NSP; ++p) {
                        tma_load_4d(dst + (size_t)p * kTcHidTile, &maps.Wh, w_full + buf, 0, 0, p, l);
                        tma_load_4d(dst + (size_t)p * kTcHidTile + kTcHidTile / 2, &maps.Wh, w_full + buf, 64, 0, p, l);
                    }
                }
                ps_syncwarp();
            };
            auto store_A = [&](int l, int b0, int rec) {   // the operand tile = input of layer l -> its record
                const uint8_t* src = At + (size_t)((l & 1) % n_op) * kOp;
                if (ps_elect()) {
                    for (int p = 0; p < NSP; ++p) {
                        tma_store_4d(&maps.act[l], src + (size_t)p * kTcHidTile, 0, b0, p, rec);
                        tma_store_4d(&maps.act[l], src + (size_t)p * kTcHidTile + kTcHidTile / 2, 64, b0, p, rec);
                    }
                    bulk_commit();
                }
                ps_syncwarp();
            };
            PS_LEAD(tma_prefetch_desc(&maps.act[0]));
            if (resident) { for (int l = 0; l < F; ++l) load_W(l, l); }
            else { for (int w = 0; w < NW && w < F; ++w) load_W(w, w); }     // uses 0 .. NW-1 of the first unit
            int use = 0;               // layer uses so far (ring position when streaming)
            for (int i = 0; i < n_units; ++i) {
                const int q = i / n_my, t = j + (i - q * n_my) * a.n_hid;
                const int b0 = t * kTcM;
                const int rec = a.need_grad ? q : 0;
                if (lane == 0) {
                    if (q > 0) ps_spin_ge(a.cnt_f + t, a.n_hg * q);      // every h-group wrote its columns of the stage input
                    ps_trace(a, t, -1, q, 5);
                }
                ps_syncwarp();
                fence_proxy_async_all();
                if (ps_elect()) {
                    bulk_wait_read<0>();      // no record store of the previous unit still reads the first tile
                    mbar_expect_tx(a_full, kOp);
                    for (int p = 0; p < NSP; ++p) {
                        tma_load_4d(At + (size_t)p * kTcHidTile, &maps.act[0], a_full, 0, b0, p, rec);
                        tma_load_4d(At + (size_t)p * kTcHidTile + kTcHidTile / 2, &maps.act[0], a_full, 64, b0, p, rec);
                    }
                }
                ps_syncwarp();
                for (int l = 0; l < F; ++l, ++use) {
                    if (l == 0) { ps_wait(a_full, (uint32_t)i & 1u); if (lane == 0) ps_trace(a, t, -1, q, 6); }
                    else {
                        ps_wait(a_ready, (uint32_t)(i * F + l - 1) & 1u);
                        if (a.need_grad) store_A(l, b0, rec);
                    }
                    const int buf = resident ? l : use % NW;
                    if (resident) { if (i == 0) ps_wait(w_full + buf, 0); }
                    else ps_wait(w_full + buf, (uint32_t)(use / NW) & 1u);
                    tc_fence_after();
                    ps_gemm_kmajor<NSP>(tmem_base, smem_u32(At + (size_t)((l & 1) % n_op) * kOp), kTcHidTile, kTcM,
                                        smem_u32(Wt + (size_t)buf * kOp), kTcHidTile, 128, 128);
                    // the epilogue of this layer writes tile (l+1)&1: one buffer -> the record store just issued reads it; two buffers ->
                    // the store of the previous layer does
                    if (ps_elect()) {
                        if (n_op == 1) bulk_wait_read<0>(); else bulk_wait_read<1>();
                        umma_commit(mma_bar);
                    }
                    ps_syncwarp();
                    if (!resident) {
                        const int64_t total_uses = (int64_t)n_units * F;
                        if ((int64_t)use + NW < total_uses) {
                            ps_wait(mma_bar, (uint32_t)use & 1u);     // the MMAs that read the buffer have completed
                            load_W((use + NW) % F, buf);
                        }
                    }
                }
                ps_wait(a_ready, (uint32_t)(i * F + F - 1) & 1u);     // input of the final layer written
                store_A(F, b0, rec);
                if (ps_elect()) {
                    bulk_wait<0>();                                    // ... and complete in global memory
                    ps_trace(a, t, -1, q, 11);
                    fence_proxy_async_all();
                    st_release_gpu(a.flag_h + t, q + 1);
                    ps_trace(a, t, -1, q, 12);
                }
                ps_syncwarp();
            }
        }
    } else if (warp < 8) {
        const int wg = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        int use = 0;
        for (int i = 0; i < n_units; ++i) {
            const int q = i / n_my, t = j + (i - q * n_my) * a.n_hid;
            for (int l = 0; l < F; ++l, ++use) {
                ps_wait(mma_bar, (uint32_t)use & 1u);
                if (tid == 0) ps_trace(a, t, -1, q, 7 + 2 * (l > 0));
                tc_fence_after();
                const uint32_t bias_a = smem_u32(bias_s + l * 128);
                const uint32_t dst = smem_u32(At + (size_t)(((l + 1) & 1) % n_op) * kOp);
                const int act = a.act[l];
#pragma unroll
                for (int c0 = wg * 64; c0 < wg * 64 + 64; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32_issue(lane_addr + (uint32_t)c0, r);
                    float4 bb[8];
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) bb[qq] = lds128(bias_a + 4u * c0 + 16u * qq);
                    tmem_wait_ld<32>(r);
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        float v[8];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const float4 bq = bb[j8 * 2 + (jj >> 2)];
                            const float bj = (jj & 3) == 0 ? bq.x : ((jj & 3) == 1 ? bq.y : ((jj & 3) == 2 ? bq.z : bq.w));
                            v[jj] = apply_act(__uint_as_float(r[j8 * 8 + jj]) + bj, act);
                        }
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            if (NSP == 2) split_bf16x2(v[2 * jj], v[2 * jj + 1], hi[jj], lo[jj]);
                            else hi[jj] = pack_bf16x2(v[2 * jj], v[2 * jj + 1]);
                        }
                        const uint32_t o = sw128_off(row, (c0 >> 3) + j8, kTcM);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + o), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                        if (NSP == 2)
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + kTcHidTile + o), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
                    }
                }
                tc_fence_before();
                if (tid == 0) ps_trace(a, t, -1, q, 8 + 2 * (l > 0));
                fence_async_smem();     // generic-proxy writes before the async-proxy MMA / TMA store
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready);
            }
        }
    }
    tc_fence_before();
    __syncwarp();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

static inline size_t ps_fwd_smem_bytes(int Npad, int NSP, int NA, int F, int NX) {
    bool res;
    const int n_op = ps_hid_fwd_nop(NSP, F);
    const int NW = ps_hid_nw(NSP, F, &res, n_op);
    const size_t f = ps_field_fwd_layout(Npad, NSP, NA, NX).total, h = ps_hid_fwd_layout(NSP, NW, n_op).total;
    return 1024 + (f > h ? f : h);
}

template <int NSP>
__global__ void __launch_bounds__(kPsThreads, 1) persist_fwd_kernel(const __grid_constant__ PsArgs a, const __grid_constant__ PsMaps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int bid = blockIdx.x;
    if (bid < a.n_field) ps_field_fwd<NSP>(a, maps, smem, bid % a.n_hg, bid / a.n_hg);
    else ps_hidden_fwd<NSP>(a, maps, smem, bid - a.n_field);
}


// ===============================================================================================================
// backward pass
// ===============================================================================================================
// The first nv (0, 8 or 16) columns of a half-chunk of epilogue 1 for TMEM lane `row`: G = gk * dX * sech^2(pre + b3) -> bf16 (hi, lo)
// into the swizzled G tile(s) (16-byte stores, n0 is a multiple of 8).
template <int NSP, bool EXACT>
__device__ __forceinline__ void ps_bwd_half(const PsHalf& k, int nv, float gk, uint32_t b3_s, uint32_t x_s, uint32_t gs_s, uint32_t g_part_bytes,
                                            int row, int n0) {
#pragma unroll
    for (int j8 = 0; j8 < 2; ++j8) {
        if (8 * j8 < nv) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 bb = lds128(b3_s + 32u * j8 + 16u * q), dd = lds128(x_s + 32u * j8 + 16u * q);
                // gk is zero for padded rows and dX/dt rows beyond the batch are zero-filled by TMA: no NaN can enter the G tile
                v[4 * q + 0] = gk * dd.x * ps_sech2<EXACT>(__uint_as_float(k.r[8 * j8 + 4 * q + 0]) + bb.x);
                v[4 * q + 1] = gk * dd.y * ps_sech2<EXACT>(__uint_as_float(k.r[8 * j8 + 4 * q + 1]) + bb.y);
                v[4 * q + 2] = gk * dd.z * ps_sech2<EXACT>(__uint_as_float(k.r[8 * j8 + 4 * q + 2]) + bb.z);
                v[4 * q + 3] = gk * dd.w * ps_sech2<EXACT>(__uint_as_float(k.r[8 * j8 + 4 * q + 3]) + bb.w);
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (NSP == 2) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
                else hi[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
            }
            const uint32_t o = sw128_off(row, (n0 >> 3) + j8, kTcM);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(gs_s + o), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            if (NSP == 2)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(gs_s + g_part_bytes + o), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        }
    }
}

struct PsFieldBwdSmem { uint32_t Ws, As, Gs, Xs, b3s, bsum, gks, bars, g_part, total; };
__host__ __device__ inline PsFieldBwdSmem ps_field_bwd_layout(int Npad, int NSP, int NX, int Hg) {
    PsFieldBwdSmem L;
    const uint32_t NP64 = ((uint32_t)Npad + 63u) & ~63u;
    uint32_t o = 0;
    L.Ws = o; o += (uint32_t)NSP * (uint32_t)Npad * 256u;
    o = (o + 1023u) & ~1023u;
    L.As = o; o += (uint32_t)NSP * kTcHidTile;
    L.g_part = kTcM * NP64 * 2;
    L.Gs = o; o += (uint32_t)NSP * L.g_part;
    L.Xs = o; o += (uint32_t)NX * ps_bwd_xslot(Hg);
    L.b3s = o; o += (uint32_t)Npad * 4;
    L.bsum = L.Gs;                              // per-lane-quarter column sums, staged once after the last MMA has read the G tile
    L.gks = o; o += (uint32_t)Hg * kTcM * 4;    // dL/dk of the unit about to enter epilogue 1: [Hg][128 rows]
    o = (o + 15u) & ~15u;
    L.bars = o; o += 40 * 8;
    L.total = o;
    return L;
}

__device__ __forceinline__ PsUnit ps_unit_bwd(int i, int n_my, int part, int n_part, int n_q) {
    PsUnit u;
    const int qi = i / n_my;
    u.q = n_q - 1 - qi;
    u.t = part + (i - qi * n_my) * n_part;
    return u;
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// four consecutive floats in one reduction (16-byte aligned): a quarter of the instructions and LSU slots of the scalar form
__device__ __forceinline__ void red_add_f32x4(float* p, float v0, float v1, float v2, float v3) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
}
// dA^T of a tile, summed over the h-groups in L2: [32 k-quads][128 rows][4] floats — a thread (= row) owns four consecutive k as one
// 16-byte word, and the rows of a warp are contiguous: red.v4 / ld.cg.v4 / st.cg.v4, 512 contiguous bytes per warp instruction
__device__ __forceinline__ size_t ps_dat_index(int kq, int row) { return ((size_t)kq * 128 + row) * 4; }

// backward of one RK stage for one batch tile (cf. tc_field_bwd_kernel): MMA1 recompute pre -> epilogue 1 G -> dgrad P, wgrad
// dW^T (stays in TMEM for the whole pass) -> epilogue 2: P summed over the h-groups with red.global into the tile's fp32 dA^T
template <int NSP>
__device__ void ps_field_bwd(const PsArgs& a, const PsMaps& maps, uint8_t* smem, int g, int part) {
    constexpr bool EXACT = NSP == 2;
    constexpr int EW = 8, kCg = 2, KP = kTcKP;
    const int Npad = a.Npad, NX = a.NX;
    const PsFieldBwdSmem L = ps_field_bwd_layout(Npad, NSP, NX, a.Hg);
    const uint32_t w_part = (uint32_t)Npad * 256u;
    uint8_t* Ws = smem + L.Ws;
    uint8_t* As = smem + L.As;
    uint8_t* Gs = smem + L.Gs;
    uint8_t* Xs = smem + L.Xs;
    float* b3s = reinterpret_cast<float*>(smem + L.b3s);
    float* bsum = reinterpret_cast<float*>(smem + L.bsum);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* full_a = bars;
    uint64_t* wg_bar = bars + 1;
    uint64_t* gk_full = bars + 2;     // dL/dk warp -> epilogue: gks holds the unit's dL/dk (and the tile's hand-off flag has been acquired)
    uint64_t* pre_bar = bars + 3;
    uint64_t* gk_free = bars + 4;     // epilogue -> dL/dk warp: every warp has taken its dL/dk of the unit out of gks
    uint64_t* dg_bar = bars + 5;
    uint64_t* done2 = bars + 6;
    uint64_t* fin_bar = bars + 7;
    uint64_t* w_bar = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    volatile int* sig_done = reinterpret_cast<volatile int*>(bars + 10);
    uint64_t* x_full = bars + 12;     // [NX <= 8]
    uint64_t* x_free = bars + 20;     // [NX <= 8]
    uint64_t* pre_bar2 = bars + 28;   // second recompute accumulator (Npad <= 128)
    uint64_t* g_blk = bars + 29;      // [<= 4] epilogue -> MMA issuers: 64-column block b of the G tile written by all 8 warps
    uint64_t* flush_bar = bars + 38;  // epilogue -> wgrad thread: the weight-gradient accumulator has been added to the global sum (8 arrivals)
    uint64_t* full_lo = bars + 36;    // bf16x3: the lo part of the activation tile has landed (full_a: the hi part)
    uint64_t* lo_bar = bars + 33;     // wgrad issuer -> producer: every MMA that reads the lo part of the activation tile has completed
    float* gks = reinterpret_cast<float*>(smem + L.gks);
    const int n_blk = (Npad + 63) >> 6;
    // TMEM columns.  Npad <= 128: two recompute accumulators (pre of unit i+1 is formed while the epilogues of unit i run) at 0 and
    // 384, P at 128, dW^T at 256.  Wider h-groups: one accumulator at 0 that P re-uses, dW^T at 256.
    const bool pre2 = Npad <= 128;
    const uint32_t p_col = pre2 ? 128u : 0u;
    const bool narrow = ps_bwd_narrow(a.Hg);
    const int xw = ps_bwd_xw(a.Hg);
    const uint32_t xslot = ps_bwd_xslot(a.Hg);
    PsXSeq xq = ps_xseq(a.Hg, a.Cp);
    if (narrow) xq.nch = (a.Cp + 15) / 16;

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform
    const int n_my = part < a.n_mt ? (a.n_mt - part + a.n_part - 1) / a.n_part : 0;
    const int n_q = a.n_steps * a.NS;
    const int n_units = n_my * n_q;

    if (warp == 0) tmem_alloc(tmem_slot, 512);
#ifdef NCDE_PS_DBG_PRINT
    if (tid == 0 && blockIdx.x == 0) g_ps_dbg_bars = smem_u32(bars);
#endif
    if (tid == 0) {
        mbar_init(full_a, 1); mbar_init(full_lo, 1);   // bf16x3: the two parts of the tile are loaded separately
        mbar_init(wg_bar, 1); mbar_init(gk_full, 1); mbar_init(pre_bar, 1); mbar_init(pre_bar2, 1);
        for (int i = 0; i < 4; ++i) mbar_init(g_blk + i, EW);
        mbar_init(gk_free, EW); mbar_init(lo_bar, 1); mbar_init(flush_bar, EW);
        mbar_init(dg_bar, 1); mbar_init(done2, EW); mbar_init(fin_bar, 1); mbar_init(w_bar, 1);
        for (int i = 0; i < kPsMaxSlots; ++i) { mbar_init(x_full + i, 1); mbar_init(x_free + i, ps_bwd_narrow(a.Hg) ? 4 : 8); }
        *sig_done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < Npad; i += kPsThreads) b3s[i] = a.b3[(size_t)g * Npad + i];
    const int S = a.Hg * a.Cp;                          // valid columns (multiple of 8); [S, Npad) is zero padding
    if (S < Npad && tid < kTcM)
        for (int p = 0; p < NSP; ++p) *reinterpret_cast<uint4*>(Gs + (size_t)p * L.g_part + sw128_off(tid, S >> 3, kTcM)) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if ((kPsConverged || lane == 0) && n_units > 0) {     // see ps_elect
            if (ps_elect()) {
                tma_prefetch_desc(&maps.act[a.F]);
                mbar_expect_tx(w_bar, (uint32_t)NSP * w_part);
                for (int p = 0; p < NSP; ++p) {
                    tma_load_3d(Ws + (size_t)p * w_part, &maps.W3, w_bar, 0, g * Npad, p);
                    tma_load_3d(Ws + (size_t)p * w_part + (size_t)Npad * 128, &maps.W3, w_bar, 64, g * Npad, p);
                }
            }
            ps_syncwarp();
            // The saved records are cold (1-2 GB written by the forward pass): the tile of unit i+1 is pulled into L2 while unit i is
            // being worked on, so that its TMA — which can only start when wgrad(i) has released the buffer — is an L2 hit.
            // Parts [p0, p1) of the activation tile of unit i.  bf16x3: the two halves of the buffer swap roles from unit to unit — the half
            // that the weight-gradient MMAs release first (it held the lo part, which one operand pair of three reads) receives the HI
            // part of the next tile, so that two thirds of the next recompute GEMM run before the other half is free at all.
            auto load_A = [&](int i, int p0, int p1) {
                if (ps_elect()) {
                    const PsUnit u = ps_unit_bwd(i, n_my, part, a.n_part, n_q);
                    for (int p = p0; p < p1; ++p) {
                        uint64_t* bar = p == 0 ? full_a : full_lo;
                        uint8_t* dst = As + (size_t)(NSP == 2 ? (p ^ (i & 1)) : p) * kTcHidTile;
                        mbar_expect_tx(bar, kTcHidTile);
                        tma_load_4d(dst, &maps.act[a.F], bar, 0, u.t * kTcM, p, u.q);
                        tma_load_4d(dst + kTcHidTile / 2, &maps.act[a.F], bar, 64, u.t * kTcM, p, u.q);
                    }
                    if (p0 == 0 && i + 1 < n_units) {
                        const PsUnit v = ps_unit_bwd(i + 1, n_my, part, a.n_part, n_q);
                        const int rows = min(kTcM, a.B - v.t * kTcM);
                        for (int p = 0; p < NSP; ++p)
                            l2_prefetch_bulk(a.rec0 + (size_t)v.q * a.rec_stride + a.act_off[a.F] + (size_t)p * a.Bp * 128 + (size_t)v.t * kTcM * 128,
                                             (uint32_t)rows * 256u);
                    }
                }
                ps_syncwarp();
            };
            load_A(0, 0, NSP);
            ps_wait(w_bar, 0);
            const uint32_t As_s = smem_u32(As), Ws_s = smem_u32(Ws), Gs_s = smem_u32(Gs);
            for (int i = 0; i < n_units; ++i) {
                const PsUnit u = ps_unit_bwd(i, n_my, part, a.n_part, n_q);
                const uint32_t ph = (uint32_t)i & 1u;
                ps_wait(full_a, ph);
                // one accumulator: epilogue 2 of the previous unit must have read P out of it; two: accumulator i & 1 was released by
                // epilogue 1 of unit i-2 (the last G block of unit i-1 was awaited below, so that one is long complete)
                if (!pre2 && i > 0) ps_wait(done2, ph ^ 1u);
                tc_fence_after();
                {
                    const uint32_t d0 = tmem_base + ((pre2 && (i & 1)) ? 384u : 0u);
                    if (NSP == 2) {
                        // hi . W_hi + hi . W_lo as soon as the hi part is there, lo . W_hi when the other half has been freed and filled
                        const uint32_t a_hi = As_s + ((i & 1) ? kTcHidTile : 0u), a_lo = As_s + ((i & 1) ? 0u : kTcHidTile);
#pragma unroll 1
                        for (int w = 0; w < 2; ++w) ps_issue_kmajor(d0, a_hi, Ws_s + (w ? w_part : 0u), Npad, Npad, w == 0);
                        ps_wait(full_lo, ph);
                        tc_fence_after();
                        ps_issue_kmajor(d0, a_lo, Ws_s, Npad, Npad, false);
                    } else {
                        ps_gemm_kmajor<NSP, true>(d0, As_s, kTcHidTile, kTcM, Ws_s, w_part, Npad, Npad);
                    }
                }
                if (lane == 0) { ps_trace(a, u.t, g, n_q - 1 - u.q, 1); ps_trace_all(a, u.t, g + part * a.n_hg, n_q - 1 - u.q, 1); }
                PS_LEAD(umma_commit((pre2 && (i & 1)) ? pre_bar2 : pre_bar));
                ps_syncwarp();
                if (lane == 0) ps_trace(a, u.t, g, n_q - 1 - u.q, 0);
                // The G tile is handed over in 64-column blocks (one 128-byte swizzle block each): the weight-gradient MMAs of block b
                // — and, when P has its own TMEM columns, the k-steps of dgrad that read it — run under the rest of epilogue 1.
                const uint32_t idesc_dg = make_idesc(kTcM, KP, 0, 1);
                const int nks = Npad / 16;
                auto issue_dgrad = [&](int blk, bool clear) {
                    // dgrad: D[128 x KP] (+)= G (K-major over n) . W3 (MN-major: N = k contiguous, K = n rows), the (up to) four k-steps
                    // that read block blk of G; fully unrolled, descriptors advanced from one base per operand
                    const int nk = min(4, nks - 4 * blk);
#pragma unroll 1
                    for (int pr = 0; pr < (NSP == 2 ? 3 : 1); ++pr) {
                        const uint64_t gd = make_sdesc(Gs_s + (pr == 1 ? L.g_part : 0u) + (uint32_t)blk * (uint32_t)kTcM * 128u, 16, 1024);
                        const uint64_t wd = make_sdesc(Ws_s + (pr == 2 ? w_part : 0u) + (uint32_t)blk * 8192u, (uint32_t)Npad * 128u, 1024);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            if (kk < nk)
                                ps_umma(tmem_base + p_col, ps_desc_advance(gd, (uint32_t)kk * 32u), ps_desc_advance(wd, (uint32_t)kk * 2048u), idesc_dg,
                                        (!clear || pr > 0 || kk > 0) ? 1u : 0u);
                    }
                };
                for (int blk = 0; blk < n_blk; ++blk) {
                    ps_wait(g_blk + blk, ph);                         // block written by all 8 warps
                    if (lane == 0) ps_trace(a, u.t, g, n_q - 1 - u.q, blk == 0 ? 18 : 19);
                    if (blk == 0 && pre2 && i > 0) ps_wait(done2, ph ^ 1u);      // P of the previous unit has been read out
                    tc_fence_after();
                    const bool last = blk == n_blk - 1;
                    if (pre2) issue_dgrad(blk, blk == 0);
                    else if (last) {                                  // P re-uses the columns of pre: every warp has read it out by now
                        for (int b2 = 0; b2 < n_blk; ++b2) issue_dgrad(b2, b2 == 0);
                    }
                    if (last) { PS_LEAD(umma_commit(dg_bar)); if (lane == 0) ps_trace(a, u.t, g, n_q - 1 - u.q, 20); }
                }
                if (i + 1 < n_units) {
                    // the activation tile is released by the wgrad issuer (warp 9): the lo part first, so that half of the next tile's
                    // load runs under the remaining weight-gradient MMAs
                    while (*sig_done < i) {}                     // keeps the signaller within one unit of the pipeline
                    if (NSP == 2) { ps_wait(lo_bar, ph); load_A(i + 1, 0, 1); }      // hi part of the next tile -> the half released first
                    ps_wait(wg_bar, ph);
                    if (lane == 0) ps_trace(a, u.t, g, n_q - 1 - u.q, 15);
                    load_A(i + 1, NSP == 2 ? 1 : 0, NSP == 2 ? 2 : 1);
                }
            }
        }
    } else if (warp == 9) {
        // Weight-gradient issuer.  A single thread issues a tcgen05.mma every ~70 ns next to the busy epilogue warps of its scheduler, and
        // a backward unit of bf16x3 needs 93 of them: with one issuer the MMA issue itself was the critical path of the unit (traced).
        // The weight-gradient MMAs only meet the producer's at the G blocks and at the release of the activation tile, so they get their
        // own thread (on another scheduler): D[KP x 64 b ..] += A^T (MN-major: M = k contiguous, K = m rows) . G block b (MN-major).
        if ((kPsConverged || lane == 0) && n_units > 0) {
            const uint32_t As_s = smem_u32(As), Gs_s = smem_u32(Gs);
            for (int i = 0; i < n_units; ++i) {
                const PsUnit u = ps_unit_bwd(i, n_my, part, a.n_part, n_q);
                const uint32_t ph = (uint32_t)i & 1u;
                ps_wait(full_a, ph);                                  // (long complete: the producer's recompute MMAs have read the tile)
                if (NSP == 2) ps_wait(full_lo, ph);
                // every kPsFlushUnits units the epilogue warps add dW^T to the global fp32 sum and the accumulation starts afresh: one TMEM
                // accumulator over a whole pass (4544 units x 24 MMAs on cfg 5) was grouping-sensitive at the 3-5e-4 level
                const bool fresh = i % kPsFlushUnits == 0;
                if (fresh && i > 0) ps_wait(flush_bar, (uint32_t)(i / kPsFlushUnits - 1) & 1u);
                for (int blk = 0; blk < n_blk; ++blk) {
                    ps_wait(g_blk + blk, ph);                         // block written by all 8 warps
                    tc_fence_after();
                    const bool last = blk == n_blk - 1;
                    const int nb = min(64, Npad - 64 * blk);
                    const uint32_t idesc = make_idesc(KP, nb, 1, 1);
                    const uint32_t g_blk_off = (uint32_t)blk * (uint32_t)kTcM * 128u;
#pragma unroll 1
                    for (int pi = 0; pi < (NSP == 2 ? 3 : 1); ++pi) {
                        const int pr = NSP == 2 ? (pi == 0 ? 1 : (pi == 1 ? 0 : 2)) : 0;     // the pair that reads A's lo part goes first
                        const uint64_t ad = make_sdesc(As_s + ((NSP == 2 && ((pr == 1) != ((i & 1) != 0))) ? kTcHidTile : 0u), (uint32_t)kTcM * 128u, 1024);   // bf16x3: halves swap per unit
                        const uint64_t gd = make_sdesc(Gs_s + (pr == 2 ? L.g_part : 0u) + g_blk_off, (uint32_t)kTcM * 128u, 1024);
#pragma unroll
                        for (int ks = 0; ks < kTcM / 16; ++ks) {
                            const uint32_t off = (uint32_t)ks * 2048u;
                            ps_umma(tmem_base + kTcDwCol + 64u * (uint32_t)blk, ps_desc_advance(ad, off), ps_desc_advance(gd, off), idesc,
                                    (ks > 0 || pi > 0 || !fresh) ? 1u : 0u);
                        }
                        if (NSP == 2 && pi == 0 && last) { PS_LEAD(umma_commit(lo_bar)); }
                    }
                }
                PS_LEAD(umma_commit(wg_bar));
                if (lane == 0) ps_trace(a, u.t, g, n_q - 1 - u.q, 21);
            }
            PS_LEAD(umma_commit(fin_bar));
        }
    } else if (warp == 10) {
        if (lane == 0) {
            tma_prefetch_desc(&maps.X);
            int slot = 0;
            uint32_t lap = 0;          // completed laps of the ring
            for (int i = 0; i < n_units; ++i) {
                const PsUnit u = ps_unit_bwd(i, n_my, part, a.n_part, n_q);
                if (i + 1 < n_units) {   // next unit's dX/dt tile (cold, contiguous rows) -> L2
                    const PsUnit un = ps_unit_bwd(i + 1, n_my, part, a.n_part, n_q);
                    const int rows = min(kTcM, a.B - un.t * kTcM);
                    l2_prefetch_bulk(a.dx0 + (size_t)un.q * a.dx_stride + (size_t)un.t * kTcM * a.Cp, (uint32_t)rows * a.Cp * 4u);
                }
                for (int p = 0; p < xq.n_pass; ++p)
                    for (int j = 0; j < xq.nch; ++j) {
                        if (lap > 0) ps_wait(x_free + slot, (lap - 1) & 1u);
                        mbar_expect_tx(x_full + slot, xslot);
                        tma_load_3d(Xs + (size_t)slot * xslot, &maps.X, x_full + slot, xw * j, u.t * kTcM, u.q);
                        if (++slot == NX) { slot = 0; ++lap; }
                    }
            }
        }
    } else if (warp == 11) {
        // dL/dk former (one unit ahead of the epilogue warps) and signaller.  lane = row (mod 32); entries (row, h) of the tile are owned
        // by this CTA alone, and gy is read and written by the same lane in program order.
        const float third = 0.3333333432674408f;
        auto signal_unit = [&](int i) {          // every epilogue warp has issued its share of unit i's partial sums
            if (lane == 0) {
                const PsUnit u = ps_unit_bwd(i, n_my, part, a.n_part, n_q);
                ps_wait(done2, (uint32_t)i & 1u);
                red_release_gpu_add(a.cnt_f + u.t, 1);
                ps_trace(a, u.t, g, n_q - 1 - u.q, 6); ps_trace_all(a, u.t, g + part * a.n_hg, n_q - 1 - u.q, 6);
                *sig_done = i + 1;
            }
        };
        for (int i = 0; i < n_units; ++i) {
            const PsUnit un = ps_unit_bwd(i, n_my, part, a.n_part, n_q);
            const int s = un.q / a.NS, ist = un.q - s * a.NS;
            const float dt = __ldg(a.dt + s);
            const int64_t b0 = (int64_t)un.t * kTcM;
            float* gy_new = a.yT[s & 1];
            const float* gy_old = a.yT[(s + 1) & 1];
            const int emit = __ldg(a.emit_idx + s);
            const bool first = ist == a.NS - 1;          // first backward stage of step s
            const bool carry = first && s + 1 < a.n_steps;
            // the gradients this unit's dL/dk is formed from: the hidden CTA has finished every later stage of the tile (the same
            // acquire makes the tile's zeroed dA^T visible to epilogue 2, through gk_full)
            // With several tiles per CTA the flag of unit i was set long ago and dL/dk is formed a whole unit ahead; with one tile it
            // depends on this CTA's own signal for unit i-1, which then has to go out first (the order below can never deadlock).
            bool signalled = i == 0;
            {
                int ready = 0;
                if (lane == 0) ready = ld_acquire_gpu(a.flag_h + un.t) >= n_q - 1 - un.q;
                ready = __shfl_sync(0xffffffffu, ready, 0);
                if (!ready && !signalled) { signal_unit(i - 1); signalled = true; }
            }
            if (lane == 0) { ps_spin_ge(a.flag_h + un.t, n_q - 1 - un.q); ps_trace(a, un.t, g, n_q - 1 - un.q, 16); }
            __syncwarp();
            if (i > 0) {      // the epilogue warps have taken the previous unit's values out of gks
                if (lane == 0) ps_wait(gk_free, ((uint32_t)i & 1u) ^ 1u);
                __syncwarp();
            }
            for (int hl = 0; hl < a.Hg; ++hl) {
                const int h = g * a.Hg + hl;
                const bool h_ok = h < a.H;
                const size_t hoff = (size_t)(h_ok ? h : 0) * a.Bp;
                // every load of the four row groups first (L2 latency once, not four times), then the arithmetic
                float v0[kTcM / 32], v1[kTcM / 32], v2[kTcM / 32], v3[kTcM / 32], v4[kTcM / 32];
#pragma unroll
                for (int rr = 0; rr < kTcM / 32; ++rr) {
                    const int64_t b = b0 + rr * 32 + lane;
                    const size_t off = hoff + (size_t)(b < a.B ? b : a.B - 1);
                    v0[rr] = first ? gy_old[off] : gy_new[off];      // written by this very lane (plain accesses: same-thread order)
                    v1[rr] = v2[rr] = v3[rr] = v4[rr] = 0.f;
                    if (first) {
                        if (carry) {
                            v1[rr] = __ldcg(a.kT[0] + off);
                            if (a.NS > 1) v2[rr] = __ldcg(a.kT[1] + off);
                            if (a.NS > 2) v3[rr] = __ldcg(a.kT[2] + off);
                            if (a.NS > 3) v4[rr] = __ldcg(a.kT[3] + off);
                        }
                        if (emit >= 0) v1[rr] += __ldg(a.grad_out + ((size_t)emit * a.B + (b < a.B ? b : a.B - 1)) * a.H + (h_ok ? h : 0));
                    } else {
                        if (ist == 0) v1[rr] = __ldcg(a.kT[1] + off);
                        if (ist <= 1) v2[rr] = __ldcg(a.kT[2] + off);
                        v3[rr] = __ldcg(a.kT[3] + off);
                    }
                }
#pragma unroll
                for (int rr = 0; rr < kTcM / 32; ++rr) {
                    const int row = rr * 32 + lane;
                    const int64_t b = b0 + row;
                    float gk = 0.f;
                    if (b < a.B && h_ok) {
                        if (first) {
                            // the gradient of the step's end state = the one carried over from step s+1 + the stage-input gradients of
                            // step s+1 (unit Jacobian w.r.t. the state) + the output gradient at this point
                            const float gy = v0[rr] + v1[rr] + v2[rr] + v3[rr] + v4[rr];
                            gy_new[hoff + b] = gy;
                            gk = (a.method == NCDE_RK4_38 ? dt * 0.125f : dt) * gy;
                        } else {
                            // rk_common.py:106-114 transposed: dL/dk_i = c_i dt gy + sum over later stages q of d(stage input q)/dk_i * dz_q
                            const float c8 = dt * 0.125f;
                            gk = ((ist == 0) ? c8 : 3.f * c8) * v0[rr];
                            if (ist == 0) {
                                gk = fmaf(dt * third, v1[rr], gk);
                                gk = fmaf(-(dt * third), v2[rr], gk);
                                gk = fmaf(dt, v3[rr], gk);
                            } else if (ist == 1) {
                                gk = fmaf(dt, v2[rr], gk);
                                gk = fmaf(-dt, v3[rr], gk);
                            } else {
                                gk = fmaf(dt, v3[rr], gk);
                            }
                        }
                    }
                    gks[hl * kTcM + row] = gk;
                }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(gk_full); ps_trace(a, un.t, g, n_q - 1 - un.q, 17); }
            if (!signalled) signal_unit(i - 1);
        }
        if (n_units > 0) signal_unit(n_units - 1);
    } else if (warp < 8) {
        const int cg = warp >> 2;                    // column group
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const int h_begin = a.Hg >= 2 ? (cg == 0 ? 0 : a.Hg / 2) : 0;
        const int h_end = a.Hg >= 2 ? (cg == 0 ? a.Hg / 2 : a.Hg) : 1;
        const uint32_t xs_row = smem_u32(Xs) + (uint32_t)row * ((narrow ? kPsXPitchN : kPsXPitch) * 4);
        int slot = 0;
        uint32_t lap = 0;
        uint32_t xbase = 0;        // one-row groups: chunks of all earlier units (the ring position of a chunk follows from its number)
        const int col_begin = (cg * Npad / kCg) & ~15;
        const int col_end = cg == kCg - 1 ? Npad : (((cg + 1) * Npad / kCg) & ~15);
        const uint32_t b3_s = smem_u32(b3s), gs_s = smem_u32(Gs);
        // Bias gradient = column sums of G.  Each warp sums exactly what it wrote — its 32 rows x its own 8-column chunks — right after
        // epilogue 1, while the tensor core works on dgrad (no CTA-wide barrier, no second reader of other warps' rows).  Lane = (chunk
        // slot, row subset); the slot -> chunk map is the same for every unit, so the running sums stay in registers for the whole pass.
        // dW^T (TMEM lanes = k, columns = n) is ADDED to the global accumulator [part][g*Npad + n][k] (zeroed by the host): at the end of
        // the pass and every kPsFlushUnits units on the way (this CTA alone owns its slice; a thread its lane k and half of the columns)
        auto flush_dw = [&]() {
            for (int n0 = col_begin; n0 < col_end; n0 += 16) {
                float v[16];
                tmem_ld16(lane_addr + kTcDwCol + (uint32_t)n0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float* p = a.dW3acc + (((size_t)part * a.n_hg + g) * Npad + n0 + j) * 128 + row;
                    *p += v[j];
                }
            }
        };
        int bs_nc;                                   // my 8-column chunks
        if (a.Hg >= 2) bs_nc = (h_end - h_begin) * (a.Cp >> 3);
        else { const int c8 = a.Cp >> 3; bs_nc = (c8 / 4) * 2 + max(0, min(2, c8 % 4 - 2 * cg)); }     // 16-column chunks j with j & 1 == cg
        const int bs_S = bs_nc <= 8 ? 8 : 16;        // chunk slots per pass over the lanes (bs_nc <= 16: TMEM limits Npad / 2 to 120 columns)
        const int bs_slot = lane % bs_S, bs_rsub = lane / bs_S, bs_R = 32 / bs_S;
        const int bs_chunk = a.Hg >= 2 ? h_begin * (a.Cp >> 3) + bs_slot : (bs_slot >> 1) * 4 + cg * 2 + (bs_slot & 1);
        const bool bs_on = bs_slot < bs_nc;
        // one-row groups: the eight lanes of a quarter-warp read chunks of BOTH 64-column blocks, whose 16-byte positions coincide in the
        // swizzled rows; the lanes of the second block take the row two below / above instead (position ^ 2: disjoint again)
        const int bs_rx = (narrow && ((bs_chunk >> 3) & 1)) ? 2 : 0;
        float bacc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bacc[j] = 0.f;

        for (int i = 0; i < n_units; ++i) {
            const PsUnit un = ps_unit_bwd(i, n_my, part, a.n_part, n_q);
            const uint32_t ph = (uint32_t)i & 1u;
            const int64_t b0 = (int64_t)un.t * kTcM;
            const int64_t b = b0 + row;
            const bool row_ok = b < a.B;
            // dL/dk of this thread's (row, h) entries, formed by warp 11 while the previous unit was in its second epilogue
            ps_wait(gk_full, ph);
            if (lane == 0) ps_trace(a, un.t, g, n_q - 1 - un.q, 32 + warp);
            float gkv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (p < xq.n_pass && h_begin + p < h_end) gkv[p] = gks[(h_begin + p) * kTcM + row];
            __syncwarp();
            if (lane == 0) mbar_arrive(gk_free);
            if (pre2) ps_wait((i & 1) ? pre_bar2 : pre_bar, (uint32_t)(i >> 1) & 1u);
            else ps_wait(pre_bar, ph);
            if (tid == 0) { ps_trace(a, un.t, g, n_q - 1 - un.q, 2); ps_trace_all(a, un.t, g + part * a.n_hg, n_q - 1 - un.q, 2); }
            tc_fence_after();
            const uint32_t pre_addr = lane_addr + ((pre2 && (i & 1)) ? 384u : 0u);
            // ---- epilogue 1 ----
            PsHalf H, Hn;             // 16 accumulator columns being worked on / in flight from TMEM
            int nb_done = 0;          // 64-column blocks of the G tile this warp has handed over
            // every column of mine below `col` is written: hand over the blocks that lie entirely below it
            auto hand_over = [&](int col) {
                if (nb_done < n_blk && 64 * (nb_done + 1) <= col) {
                    fence_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    do {
                        if (lane == 0) mbar_arrive(g_blk + nb_done);
                        ++nb_done;
                    } while (nb_done < n_blk && 64 * (nb_done + 1) <= col);
                }
            };
            if (narrow) {
                // One hidden row per group: 16-channel chunk j of the unit is warp group (j & 1)'s and ONLY that group's — it alone waits
                // for the slot and releases it (x_free counts 4 arrivals in this mode), so a warp runs 3-4 hand-shakes per unit, not 7.
                // ONE inlined copy of the 16-column body (with a copy per accumulator buffer and per unrolled chunk the kernel had outgrown
                // the instruction cache: 17 % of its stall samples were instruction fetches): the accumulator columns of my next chunk
                // are loaded from TMEM while the current one is worked on and change registers by 16 moves after the wait.
                const float gk = gkv[0];
                if (cg < xq.nch) ps_half_issue(Hn, pre_addr + (uint32_t)(16 * cg));
                else hand_over(1 << 30);
#pragma unroll 1
                for (int j = cg; j < xq.nch; j += 2) {
                    const uint32_t cj = xbase + (uint32_t)j;                // running chunk number -> ring position
                    const uint32_t xlap = cj / (uint32_t)NX, xslot_i = cj - xlap * (uint32_t)NX;
                    ps_wait(x_full + xslot_i, xlap & 1u);
                    const int col = 16 * j, jn = j + 2;
                    const int nv = min(16, a.Cp - col);
                    const uint32_t xs = xs_row + xslot_i * xslot;
                    tmem_wait_ld<16>(Hn.r);
#pragma unroll
                    for (int e = 0; e < 16; ++e) H.r[e] = Hn.r[e];
                    const int next_col = jn < xq.nch ? 16 * jn : (1 << 30);
                    if (jn < xq.nch) ps_half_issue(Hn, pre_addr + (uint32_t)next_col);
                    ps_bwd_half<NSP, EXACT>(H, nv, gk, b3_s + 4u * col, xs, gs_s, L.g_part, row, col);
                    hand_over(next_col);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(x_free + xslot_i);
                }
                xbase += (uint32_t)xq.nch;
            } else {
                // wider groups: both halves of a 32-channel chunk are mine; two named buffers, software-pipelined (the single-copy loop
                // above was measured 17 % slower here: cfg 5 in bf16, 31.0 vs 26.6 ms per step)
            {
                const bool any = h_begin < h_end;
                if (any) ps_half_issue(H, pre_addr + (uint32_t)(h_begin * a.Cp));
                hand_over(any ? h_begin * a.Cp : (1 << 30));
            }
            for (int p = 0; p < xq.n_pass; ++p) {
                const int hl = h_begin + p;
                const bool pass_ok = hl < h_end;
                const float gk = p == 0 ? gkv[0] : (p == 1 ? gkv[1] : (p == 2 ? gkv[2] : gkv[3]));
                const int colbase = hl * a.Cp;
                for (int j = 0; j < xq.nch; ++j) {
                    ps_wait(x_full + slot, lap & 1u);
                    const bool mine = pass_ok;
                    if (mine) {
                        const int c = 32 * j;
                        const int nv0 = min(16, a.Cp - c), nv1 = max(0, min(16, a.Cp - c - 16));
                        const uint32_t xs = xs_row + (uint32_t)slot * kPsXSlot;
                        tmem_wait_ld<16>(H.r);
                        if (nv1 > 0) ps_half_issue(Hn, pre_addr + (uint32_t)(colbase + c + 16));
                        ps_bwd_half<NSP, EXACT>(H, nv0, gk, b3_s + 4u * (colbase + c), xs, gs_s, L.g_part, row, colbase + c);
                        tmem_wait_ld<16>(Hn.r);
                        int next_col = 1 << 30;     // first column of my next chunk in this unit
                        {
                            int pn = p, jn = j + 1;
                            if (jn >= xq.nch) { pn = p + 1; jn = 0; }
                            if (pn < xq.n_pass && h_begin + pn < h_end && jn < xq.nch) {
                                next_col = (h_begin + pn) * a.Cp + 32 * jn;
                                ps_half_issue(H, pre_addr + (uint32_t)next_col);
                            }
                        }
                        ps_bwd_half<NSP, EXACT>(Hn, nv1, gk, b3_s + 4u * (colbase + c + 16), xs + 64u, gs_s, L.g_part, row, colbase + c + 16);
                        hand_over(next_col);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(x_free + slot);
                    if (++slot == NX) { slot = 0; ++lap; }
                }
            }
            }
            hand_over(1 << 30);       // (a warp without columns in this unit hands every block over here)
            if (lane == 0) ps_trace(a, un.t, g, n_q - 1 - un.q, 24 + warp);
            if (tid == 0) { ps_trace(a, un.t, g, n_q - 1 - un.q, 3); ps_trace_all(a, un.t, g + part * a.n_hg, n_q - 1 - un.q, 3); }
            __syncwarp();
            if (bs_on) {
                const int r0 = (warp & 3) * 32 + bs_rsub;
#pragma unroll 4
                for (int r = r0; r < (warp & 3) * 32 + 32; r += bs_R) {
#pragma unroll
                    for (int p = 0; p < NSP; ++p) {
                        uint32_t w4[4];
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w4[0]), "=r"(w4[1]), "=r"(w4[2]), "=r"(w4[3])
                                     : "r"(gs_s + (uint32_t)p * L.g_part + sw128_off(r ^ bs_rx, bs_chunk, kTcM)));
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            bacc[2 * j] += __uint_as_float(w4[j] << 16);
                            bacc[2 * j + 1] += __uint_as_float(w4[j] & 0xffff0000u);
                        }
                    }
                }
            }
            if (tid == 0) ps_trace(a, un.t, g, n_q - 1 - un.q, 14);
            ps_wait(dg_bar, ph);
            if (tid == 0) { ps_trace(a, un.t, g, n_q - 1 - un.q, 4); ps_trace_all(a, un.t, g + part * a.n_hg, n_q - 1 - un.q, 4); }
            tc_fence_after();
            // ---- epilogue 2: this h-group's share of dL/d(final-layer input), summed over the groups in L2 ----
            {
                const int kb = cg * (KP / kCg);
                float* dtile = a.dAT + (size_t)un.t * 128 * 128;
                uint32_t r0[32], r1[32];
                tmem_ld32_issue(lane_addr + p_col + (uint32_t)kb, r0);
                tmem_wait_ld<32>(r0);
                tmem_ld32_issue(lane_addr + p_col + (uint32_t)kb + 32u, r1);
                if (row_ok) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        red_add_f32x4(dtile + ps_dat_index(kb / 4 + j4, row), __uint_as_float(r0[4 * j4]), __uint_as_float(r0[4 * j4 + 1]),
                                      __uint_as_float(r0[4 * j4 + 2]), __uint_as_float(r0[4 * j4 + 3]));
                }
                tmem_wait_ld<32>(r1);
                if (row_ok) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        red_add_f32x4(dtile + ps_dat_index(kb / 4 + 8 + j4, row), __uint_as_float(r1[4 * j4]), __uint_as_float(r1[4 * j4 + 1]),
                                      __uint_as_float(r1[4 * j4 + 2]), __uint_as_float(r1[4 * j4 + 3]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (tid == 0) { ps_trace(a, un.t, g, n_q - 1 - un.q, 5); ps_trace_all(a, un.t, g + part * a.n_hg, n_q - 1 - un.q, 5); }
            if (lane == 0) mbar_arrive(done2);
            if ((i + 1) % kPsFlushUnits == 0 && i + 1 < n_units) {
                ps_wait(wg_bar, ph);           // every weight-gradient MMA of the unit has completed
                tc_fence_after();
                flush_dw();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(flush_bar);
            }
        }
        // ---- dW^T (TMEM lanes = k, columns = n) -> global accumulator [part][g*Npad + n][k];  bias gradient ----
        if (n_units > 0) {
            ps_wait(fin_bar, 0);
            tc_fence_after();
            // column sums: over the row subsets of a warp by shuffles, then over the four lane quarters through shared memory
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                for (int d = bs_S; d < 32; d <<= 1) bacc[j] += __shfl_xor_sync(0xffffffffu, bacc[j], d);
            }
            if (bs_on && bs_rsub == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) bsum[(warp & 3) * Npad + bs_chunk * 8 + j] = bacc[j];
            }
            named_bar_sync(1, kPsEpi);
            flush_dw();
            for (int n = tid; n < Npad; n += kPsEpi) {
                // columns no warp owns ([Hg * Cp, Npad) padding) were never written: they are zero by construction
                const bool owned = n < a.Hg * a.Cp;
                a.db3acc[((size_t)part * a.n_hg + g) * Npad + n] = owned ? (bsum[n] + bsum[Npad + n]) + (bsum[2 * Npad + n] + bsum[3 * Npad + n]) : 0.f;
            }
        }
    }
    tc_fence_before();
    __syncwarp();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// backward, hidden role: sum of the h-group partials (dA^T, fp32, L2) x act' -> dpre_{F-1}; input-gradient chain l = F-1 .. 0 on the
// tensor core; dz -> global.  dpre_l of every stage is kept as a record for the weight-gradient kernel.
template <int NSP>
__device__ void ps_hidden_bwd(const PsArgs& a, const PsMaps& maps, uint8_t* smem, int j) {
    bool resident;
    const int NW = ps_hid_nw(NSP, a.F, &resident, 2);
    const PsHidFwdSmem L = ps_hid_fwd_layout(NSP, NW, 2);
    constexpr uint32_t kOp = (uint32_t)NSP * kTcHidTile;
    uint8_t* Wt = smem + L.Wt;
    uint8_t* Dt = smem + L.At;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* w_full = bars;          // [NW <= 8]
    uint64_t* top_bar = bars + 8;     // producer -> epilogue: all h-groups have added their partials for this unit
    uint64_t* dg_bar = bars + 9;
    uint64_t* dp_ready = bars + 10;   // epilogue -> producer: dpre tile written (8 warp arrivals)
    uint64_t* out_done = bars + 11;   // epilogue -> producer: dz written (8 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform
    const int F = a.F;
    const int n_my = j < a.n_mt ? (a.n_mt - j + a.n_hid - 1) / a.n_hid : 0;
    const int n_q = a.n_steps * a.NS;
    const int n_units = n_my * n_q;

    if (warp == 0) tmem_alloc(tmem_slot, 128);
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(w_full + i, 1);
        mbar_init(top_bar, 1); mbar_init(dg_bar, 1); mbar_init(dp_ready, 8); mbar_init(out_done, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if ((kPsConverged || lane == 0) && n_units > 0) {     // see ps_elect
            auto load_W = [&](int l, int buf) {
                uint8_t* dst = Wt + (size_t)buf * kOp;
                if (ps_elect()) {
                    mbar_expect_tx(w_full + buf, kOp);
                    for (int p = 0; p < NSP; ++p) {
                        tma_load_4d(dst + (size_t)p * kTcHidTile, &maps.Wh, w_full + buf, 0, 0, p, l);
                        tma_load_4d(dst + (size_t)p * kTcHidTile + kTcHidTile / 2, &maps.Wh, w_full + buf, 64, 0, p, l);
                    }
                }
                ps_syncwarp();
            };
            PS_LEAD(tma_prefetch_desc(&maps.dpre));
            // layer order of the chain: F-1, F-2, ..., 0, F-1, ...   use u -> layer F-1 - (u % F)
            if (resident) { for (int l = 0; l < F; ++l) load_W(l, l); }
            else { for (int w = 0; w < NW && w < F; ++w) load_W(F - 1 - w, w); }
            int use = 0;
            for (int i = 0; i < n_units; ++i) {
                const int qi = i / n_my, q = n_q - 1 - qi, t = j + (i - qi * n_my) * a.n_hid;
                const int b0 = t * kTcM;
                if (lane == 0) {
                    ps_spin_ge(a.cnt_f + t, a.n_hg * (qi + 1));          // every h-group has added its partial of stage q
                    ps_trace(a, t, -1, qi, 7);
                    mbar_arrive(top_bar);
                }
                ps_syncwarp();
                for (int l = F - 1, n = 0; l >= 0; --l, ++n, ++use) {
                    const int bf = l & 1;
                    ps_wait(dp_ready, (uint32_t)(i * F + n) & 1u);        // dpre_l written (and fenced)
                    uint8_t* d_tile = Dt + (size_t)bf * kOp;
                    if (ps_elect()) {
                        for (int p = 0; p < NSP; ++p) {                   // keep dpre_l for the weight-gradient kernel
                            tma_store_4d(&maps.dpre, d_tile + (size_t)p * kTcHidTile, 0, b0, p, q * F + l);
                            tma_store_4d(&maps.dpre, d_tile + (size_t)p * kTcHidTile + kTcHidTile / 2, 64, b0, p, q * F + l);
                        }
                        bulk_commit();
                    }
                    ps_syncwarp();
                    const int buf = resident ? l : use % NW;
                    if (resident) { if (i == 0) ps_wait(w_full + buf, 0); }
                    else ps_wait(w_full + buf, (uint32_t)(use / NW) & 1u);
                    tc_fence_after();
                    {   // dgrad: D[128 b x 128 i] = dpre_l (K-major over o) . W_l (MN-major: N = i contiguous, K = o rows)
                        const uint32_t idesc = make_idesc(kTcM, 128, 0, 1);
                        const uint32_t d_s = smem_u32(d_tile), w_s = smem_u32(Wt + (size_t)buf * kOp);
#pragma unroll
                        for (int pr = 0; pr < (NSP == 2 ? 3 : 1); ++pr) {
                            const uint64_t dd = make_sdesc(d_s + (pr == 1 ? kTcHidTile : 0u), 16, 1024);
                            const uint64_t wd = make_sdesc(w_s + (pr == 2 ? kTcHidTile : 0u), 128u * 128u, 1024);
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks) {
                                const uint32_t a_off = (uint32_t)(ks >> 2) * (uint32_t)kTcM * 128u + (uint32_t)(ks & 3) * 32u;
                                ps_umma(tmem_base, ps_desc_advance(dd, a_off), ps_desc_advance(wd, (uint32_t)ks * 2048u), idesc,
                                        (pr > 0 || ks > 0) ? 1u : 0u);
                            }
                        }
                    }
                    if (ps_elect()) {
                        // the epilogue of this level writes buffer (l-1)&1, whose previous content was stored one level ago
                        bulk_wait_read<1>();
                        umma_commit(dg_bar);
                    }
                    ps_syncwarp();
                    if (!resident) {
                        const int64_t total_uses = (int64_t)n_units * F;
                        if ((int64_t)use + NW < total_uses) {
                            ps_wait(dg_bar, (uint32_t)use & 1u);
                            load_W(F - 1 - ((use + NW) % F), buf);
                        }
                    }
                }
                ps_wait(out_done, (uint32_t)i & 1u);
                if (ps_elect()) {
                    bulk_wait_read<0>();     // the next unit's top tile overwrites a buffer the last store may still read
                    st_release_gpu(a.flag_h + t, qi + 1);
                    ps_trace(a, t, -1, qi, 13);
                }
                ps_syncwarp();
            }
            PS_LEAD(bulk_wait<0>());
        }
    } else if (warp < 8) {
        const int wg = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        int use = 0;
        for (int i = 0; i < n_units; ++i) {
            const int qi = i / n_my, q = n_q - 1 - qi, t = j + (i - qi * n_my) * a.n_hid;
            const int s = q / a.NS, ist = q - s * a.NS;
            const int64_t b = (int64_t)t * kTcM + row;
            const bool row_ok = b < a.B;
            const __nv_bfloat16* recq = a.rec0 + (size_t)q * a.rec_stride;
            // ---- top: dpre_{F-1} = (sum over the h-groups of P) * act'(a_F) ----
            ps_wait(top_bar, (uint32_t)i & 1u);
            {
                float* dtile = a.dAT + (size_t)t * 128 * 128;
                const __nv_bfloat16* arow = recq + a.act_off[F] + (size_t)b * 128 + wg * 64;
                const uint32_t dst = smem_u32(Dt + (size_t)((F - 1) & 1) * kOp);
                const int act = a.act[F - 1];
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    float v[8];
#pragma unroll
                    for (int j4 = 0; j4 < 2; ++j4) {
                        float4* p4 = reinterpret_cast<float4*>(dtile + ps_dat_index(wg * 16 + c8 * 2 + j4, row));
                        const float4 x4 = __ldcg(p4);
                        v[4 * j4] = x4.x; v[4 * j4 + 1] = x4.y; v[4 * j4 + 2] = x4.z; v[4 * j4 + 3] = x4.w;
                        __stcg(p4, make_float4(0.f, 0.f, 0.f, 0.f));   // ready for the tile's next stage
                    }
                    uint4 av = make_uint4(0, 0, 0, 0), al = make_uint4(0, 0, 0, 0);
                    if (row_ok) {
                        av = __ldg(reinterpret_cast<const uint4*>(arow + c8 * 8));
                        if (NSP == 2 && act == NCDE_ACT_TANH) al = __ldg(reinterpret_cast<const uint4*>(arow + (size_t)a.Bp * 128 + c8 * 8));
                    }
                    const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, lw[4] = {al.x, al.y, al.z, al.w};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        float2 a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[jj]));
                        const float2 l2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[jj]));
                        a2.x += l2.x; a2.y += l2.y;
                        const float g0 = row_ok ? v[2 * jj] * act_grad_bf(a2.x, act) : 0.f;
                        const float g1 = row_ok ? v[2 * jj + 1] * act_grad_bf(a2.y, act) : 0.f;
                        if (NSP == 2) split_bf16x2(g0, g1, hi[jj], lo[jj]);
                        else hi[jj] = pack_bf16x2(g0, g1);
                    }
                    const uint32_t o = sw128_off(row, wg * 8 + c8, kTcM);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + o), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                    if (NSP == 2)
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + kTcHidTile + o), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
                }
                fence_async_smem();
                __syncwarp();
                if (tid == 0) ps_trace(a, t, -1, qi, 8);
                if (lane == 0) mbar_arrive(dp_ready);
            }
            for (int l = F - 1; l >= 0; --l, ++use) {
                ps_wait(dg_bar, (uint32_t)use & 1u);
                if (tid == 0) ps_trace(a, t, -1, qi, l == 0 ? 11 : 9);
                tc_fence_after();
                if (l > 0) {
                    const __nv_bfloat16* arow = recq + a.act_off[l] + (size_t)b * 128;
                    const uint32_t dst = smem_u32(Dt + (size_t)((l - 1) & 1) * kOp);
                    const int act = a.act[l - 1];
#pragma unroll
                    for (int c0 = wg * 64; c0 < wg * 64 + 64; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32_issue(lane_addr + (uint32_t)c0, r);
                        uint4 av[4], al[4];
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            av[qq] = row_ok ? __ldg(reinterpret_cast<const uint4*>(arow + c0 + qq * 8)) : make_uint4(0, 0, 0, 0);
                            al[qq] = (NSP == 2 && act == NCDE_ACT_TANH && row_ok)
                                         ? __ldg(reinterpret_cast<const uint4*>(arow + (size_t)a.Bp * 128 + c0 + qq * 8)) : make_uint4(0, 0, 0, 0);
                        }
                        tmem_wait_ld<32>(r);
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const uint32_t aw[4] = {av[qq].x, av[qq].y, av[qq].z, av[qq].w}, lw[4] = {al[qq].x, al[qq].y, al[qq].z, al[qq].w};
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                float2 a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[jj]));
                                const float2 l2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[jj]));
                                a2.x += l2.x; a2.y += l2.y;
                                const float g0 = row_ok ? __uint_as_float(r[8 * qq + 2 * jj]) * act_grad_bf(a2.x, act) : 0.f;
                                const float g1 = row_ok ? __uint_as_float(r[8 * qq + 2 * jj + 1]) * act_grad_bf(a2.y, act) : 0.f;
                                if (NSP == 2) split_bf16x2(g0, g1, hi[jj], lo[jj]);
                                else hi[jj] = pack_bf16x2(g0, g1);
                            }
                            const uint32_t o = sw128_off(row, (c0 >> 3) + qq, kTcM);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + o), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                            if (NSP == 2)
                                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + kTcHidTile + o), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
                        }
                    }
                    fence_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (tid == 0) ps_trace(a, t, -1, qi, 10);
                    if (lane == 0) mbar_arrive(dp_ready);
                } else {
                    // dz of stage `ist`, feature-major fp32: plain stores, coalesced over the lanes (= rows)
                    float* dz = a.kT[ist];
#pragma unroll
                    for (int c0 = wg * 64; c0 < wg * 64 + 64; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32_issue(lane_addr + (uint32_t)c0, r);
                        tmem_wait_ld<32>(r);
                        if (row_ok) {
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj)
                                if (c0 + jj < a.H) dz[(size_t)(c0 + jj) * a.Bp + b] = __uint_as_float(r[jj]);
                        }
                    }
                    tc_fence_before();
                    if (tid == 0) ps_trace(a, t, -1, qi, 12);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(out_done);
                }
            }
        }
    }
    tc_fence_before();
    __syncwarp();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

static inline size_t ps_bwd_smem_bytes(int Npad, int NSP, int F, int NX, int Hg) {
    bool res;
    const int NW = ps_hid_nw(NSP, F, &res, 2);
    const size_t f = ps_field_bwd_layout(Npad, NSP, NX, Hg).total, h = ps_hid_fwd_layout(NSP, NW, 2).total;
    return 1024 + (f > h ? f : h);
}

template <int NSP>
__global__ void __launch_bounds__(kPsThreads, 1) persist_bwd_kernel(const __grid_constant__ PsArgs a, const __grid_constant__ PsMaps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int bid = blockIdx.x;
    if (bid < a.n_field) ps_field_bwd<NSP>(a, maps, smem, bid % a.n_hg, bid / a.n_hg);
    else ps_hidden_bwd<NSP>(a, maps, smem, bid - a.n_field);
}


// ---------------------------------------------------------------------------------------------------------------
// packing for the persistent path: bf16 (hi [, lo]) operand tiles
// ---------------------------------------------------------------------------------------------------------------
// final layer: Wp[part][g][nl][k], k contiguous, zero padded (cf. pack_final_bf16_kernel)
__global__ void ps_pack_final_kernel(const float* __restrict__ W, const float* __restrict__ bias, __nv_bfloat16* __restrict__ Wp,
                                     float* __restrict__ b3, int H, int C, int Cp, int Hg, int n_hg, int Npad, int DF, int NSP) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)n_hg * Npad * 128;
    if (idx < total) {
        const int k = (int)(idx & 127);
        const int nl = (int)((idx >> 7) % Npad);
        const int g = (int)(idx / ((int64_t)128 * Npad));
        const int hl = nl / Cp, c = nl % Cp, h = g * Hg + hl;
        float v = 0.f;
        if (hl < Hg && h < H && c < C && k < DF) v = W[((int64_t)h * C + c) * DF + k];
        const __nv_bfloat16 hi = __float2bfloat16(v);
        Wp[idx] = hi;
        if (NSP == 2) Wp[total + idx] = __float2bfloat16(v - __bfloat162float(hi));
    }
    if (idx < (int64_t)n_hg * Npad) {
        const int nl = (int)(idx % Npad), g = (int)(idx / Npad);
        const int hl = nl / Cp, c = nl % Cp, h = g * Hg + hl;
        b3[idx] = (bias && hl < Hg && h < H && c < C) ? bias[(int64_t)h * C + c] : 0.f;
    }
}
// hidden layer l: Wh[l][part][o][i] zero padded to 128 x 128; bh[l][o]
__global__ void ps_pack_hidden_kernel(const float* __restrict__ W, const float* __restrict__ bias, __nv_bfloat16* __restrict__ Wh,
                                      float* __restrict__ bh, int Dout, int Din, int NSP) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 128 * 128) {
        const int o = idx >> 7, i = idx & 127;
        const float v = (o < Dout && i < Din) ? W[(size_t)o * Din + i] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16(v);
        Wh[idx] = hi;
        if (NSP == 2) Wh[128 * 128 + idx] = __float2bfloat16(v - __bfloat162float(hi));
    }
    if (idx < 128) bh[idx] = (bias && idx < Dout) ? bias[idx] : 0.f;
}
// z0 (B, H) fp32 row-major -> stage-input record 0: bf16 parts [NSP][Bp][128], feature padding zeroed
__global__ void ps_z0_record_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int Bp, int H, int NSP) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (int64_t)B * 128) {
        const int h = (int)(idx & 127);
        const int64_t b = idx >> 7;
        const float v = h < H ? src[b * H + h] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16(v);
        dst[idx] = hi;
        if (NSP == 2) dst[(size_t)Bp * 128 + idx] = __float2bfloat16(v - __bfloat162float(hi));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight and bias gradients of the hidden layers for a whole backward pass (cf. tc_hidden_wgrad_kernel), (hi, lo) operands:
//   dW_l^T[i][o] = sum over (stage, batch tile) units of  a_l^T . dpre_l     (hi*hi + lo*hi + hi*lo)
// ---------------------------------------------------------------------------------------------------------------
struct PsWgradArgs {
    int F, n_rec, n_mt, n_split;
    float* dWacc[kTcHidMaxLayers];
    float* dbacc[kTcHidMaxLayers];
};
static inline size_t ps_wgrad_smem_bytes(int NSP) {
    const int NST = NSP == 1 ? 2 : 1;
    return 1024 + (size_t)NST * 2 * NSP * kTcHidTile + 8 * 128 * 4 + 16 * 8;
}
template <int NSP>
__global__ void __launch_bounds__(kTcThreads, 1) ps_hidden_wgrad_kernel(const __grid_constant__ PsWgradArgs a, const __grid_constant__ PsMaps maps) {
    constexpr int NST = NSP == 1 ? 2 : 1;                 // pipeline stages
    constexpr uint32_t kOp = (uint32_t)NSP * kTcHidTile;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* At = smem;                          // NST x a_l operand
    uint8_t* Dt = smem + NST * kOp;              // NST x dpre_l operand
    float* bsum = reinterpret_cast<float*>(smem + 2 * NST * kOp);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * NST * kOp + 8 * 128 * 4);
    uint64_t* full = bars;            // [2]
    uint64_t* mma_done = bars + 2;    // [2]
    uint64_t* read_done = bars + 4;   // [2]
    uint64_t* fin_bar = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // provably warp-uniform
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

    if (warp == 8) {
        if (lane == 0 && n > 0) {
            auto load = [&](int j) {
                const int64_t u = u_begin + j;
                const int rec = (int)(u / a.n_mt), b0 = (int)(u % a.n_mt) * kTcM;
                const int bf = j % NST;
                mbar_expect_tx(full + bf, 2 * kOp);
                for (int p = 0; p < NSP; ++p) {
                    tma_load_4d(At + (size_t)bf * kOp + (size_t)p * kTcHidTile, &maps.act[l], full + bf, 0, b0, p, rec);
                    tma_load_4d(At + (size_t)bf * kOp + (size_t)p * kTcHidTile + kTcHidTile / 2, &maps.act[l], full + bf, 64, b0, p, rec);
                    tma_load_4d(Dt + (size_t)bf * kOp + (size_t)p * kTcHidTile, &maps.dpre, full + bf, 0, b0, p, rec * a.F + l);
                    tma_load_4d(Dt + (size_t)bf * kOp + (size_t)p * kTcHidTile + kTcHidTile / 2, &maps.dpre, full + bf, 64, b0, p, rec * a.F + l);
                }
            };
            for (int j = 0; j < NST && j < n; ++j) load(j);
            const uint32_t idesc = make_idesc(128, 128, 1, 1);
            for (int j = 0; j < n; ++j) {
                const int bf = j % NST;
                const uint32_t ph = (uint32_t)(j / NST) & 1u;
                mbar_wait(full + bf, ph);
                tc_fence_after();
                const uint32_t a_s = smem_u32(At + (size_t)bf * kOp), d_s = smem_u32(Dt + (size_t)bf * kOp);
                for (int pr = 0; pr < (NSP == 2 ? 3 : 1); ++pr) {
                    const uint32_t as = a_s + (pr == 1 ? kTcHidTile : 0u), ds = d_s + (pr == 2 ? kTcHidTile : 0u);
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t off = (uint32_t)ks * 2048u;
                        umma_bf16(tmem_base, make_sdesc(as + off, 128u * 128u, 1024), make_sdesc(ds + off, 128u * 128u, 1024), idesc,
                                  (ks > 0 || pr > 0 || j > 0) ? 1u : 0u);
                    }
                }
                umma_commit(mma_done + bf);
                if (j + NST < n) {
                    mbar_wait(mma_done + bf, ph);
                    mbar_wait(read_done + bf, ph);
                    load(j + NST);
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
            const int bf = j % NST;
            mbar_wait(full + bf, (uint32_t)(j / NST) & 1u);
            if (lane < 16) {   // 16 chunks of 8 columns x 8 row slices of 16 rows
                for (int p = 0; p < NSP; ++p) {
                    const uint32_t d_s = smem_u32(Dt + (size_t)bf * kOp + (size_t)p * kTcHidTile);
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
            }
            fence_async_smem();
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

// gy_final = gy + sum of the stage-input gradients of the first step (the persistent backward kernel leaves this last sum open)
__global__ void ps_gy_final_kernel(float* __restrict__ gyT, const float* dz0, const float* dz1, const float* dz2, const float* dz3,
                                   int n_dz, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = gyT[i] + dz0[i];
    if (n_dz > 1) g += dz1[i];
    if (n_dz > 2) g += dz2[i];
    if (n_dz > 3) g += dz3[i];
    gyT[i] = g;
}

}  // namespace ncde
