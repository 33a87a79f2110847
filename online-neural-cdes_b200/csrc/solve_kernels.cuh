// Device kernels of the fixed-grid CDE solve (fp32 arithmetic).  See DESIGN.md §3 for the decomposition.
//
// Internal state layout is FEATURE-MAJOR: y, k_i, stage inputs, activations and dX/dt are all stored [feature][Bp]
// with the batch index contiguous (Bp = B rounded up to 64), so that
//   * batch-split kernels ("hidden") read/write R consecutive rows of every feature as one segment, and
//   * weight-stationary kernels ("field") stream 64-row chunks of every feature with 16-byte loads.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace ncde {

constexpr int kThreads = 256;
constexpr int kChunk = 64;   // batch rows per field-kernel chunk
constexpr int kGnStride = 68;  // padded row of the n-major G tile (bank-conflict-free float4 stores)

// ---------------------------------------------------------------------------------------------------------------
// argument blocks (passed by value as __grid_constant__)
// ---------------------------------------------------------------------------------------------------------------
struct PathArgs {
    int kind;
    int K;
    const float* knots;
    const float* coeffs;
    const float* derivs;
    float t;  // stage time, already cast to fp32 (torchdiffeq/_impl/misc.py:181)
    const float* match;   // SmoothLinearInterpolation: (B, K-2, C, match_terms) or null
    int match_terms;
    float match_eps;
};

// ---- adaptive (dopri5) control state, resident in device memory; every time-like scalar is fp64 like the
// reference (torchdiffeq/_impl/rk_common.py:139-160) ----
struct StageTab {
    float t;                        // stage time after the cast to fp32 (and the one-ulp nudge for alpha == 1)
    float coef[NCDE_MAX_STAGES];    // beta_ij * dt in fp32
};
struct AdaptCtrl {
    double t0, dt;                  // end of the last accepted step (= start of the next attempt), next step size
    double t_lo, t_hi;              // interval of the last accepted step (dense output domain)
    double step_dt;                 // dt of the attempt in flight
    double acc_dt;                  // dt of the last accepted step (dense-output coefficients)
    double rtol, atol, min_step, max_step, safety, ifactor, dfactor;
    int accept, done, j_begin, j_end, j_out, n_out;
    int time_sign;                  // +1 forward solve; -1 adjoint (controller runs in reversed time, paths are evaluated at -tau)
    long long attempted, accepted, nfe, max_attempts;
    int flags;
    float h0, d0, d1, d2;           // initial-step selection scratch (misc.py:32-71)
    double dt_init;
    double trace[64][3];            // first 64 attempts: dt, error ratio, accepted (diagnostics)
    StageTab tab[NCDE_MAX_STAGES + 1];
};

struct HiddenFwdArgs {
    int B, Bp, H, C, Cp, R, F, Dmax;
    int D[NCDE_MAX_LAYERS + 1];   // D[l] = input width of layer l (D[0] = H, D[F] = input of the final layer)
    int ldw[NCDE_MAX_LAYERS];     // leading dimension (padded out width) of WT[l]
    int act[NCDE_MAX_LAYERS];
    const float* WT[NCDE_MAX_LAYERS];  // [D[l]][ldw[l]]  k-major
    const float* bp[NCDE_MAX_LAYERS];  // [ldw[l]]
    // stage input: see combine_stage_input()
    int combine;
    float dt;
    float coef[NCDE_MAX_STAGES];
    const float* yT;
    const float* kT[NCDE_MAX_STAGES];
    float* actT[NCDE_MAX_LAYERS + 1];  // actT[l] = [D[l] (padded to 4)][Bp], l = 0..F, for THIS stage
    // vector_field_type evaluate / derivative: rows H..H+n_u-1 of the first layer's input are X(t) / dX/dt(t), read from
    // uT [n_u][Bp] (precomputed for every stage by dx_all_kernel; may alias those rows of actT[0])
    const float* uT;
    int n_u;
    int vf;                            // ncde_vf_type; with vf != 0 and uT == null the control rows are evaluated here, at the stage time
    float* dXT;                        // [Cp][Bp]  (tensor-core path, dx_row_major: [Bp][Cp])
    int dx_row_major;
    float* ddXT;                       // [Cp][Bp] or null: d2X/dt2 (cubic paths; time-gradient component of the adjoint)
    __nv_bfloat16* abf;                // [Bp][KP] bf16 row-major copy of the final-layer input (tensor-core path) or null
    int KP;
    int w_in_smem;                     // hidden weights are staged in shared memory for the whole launch
    int wsm_off[NCDE_MAX_LAYERS];      // float offset of layer l's weights inside the shared staging area
    int wsm_floats;                    // total floats staged (layers sharing a slot share the copy)
    PathArgs path;
    // adaptive solver: stage time and combine coefficients come from device memory; nothing runs once ctrl->done
    const AdaptCtrl* ctrl;
    int tab_index;
    // continuous adjoint: the state is integrated in reversed time, its stage derivative is -f, so the increment of the
    // stage-input combination is subtracted (comb_sign = -1)
    float comb_sign;
};

struct FieldArgs {
    int B, Bp, H, Cp, DF, DFP, S, Hg, n_hg, Np, Bt;
    const float* W3T;   // [DF][Np]    k-major, n = h*Cp + c
    const float* W3R;   // [Np][DFP]   n-major (backward only)
    const float* b3p;   // [Np]
    const float* actT;  // [DF(pad4)][Bp] input of the final layer
    const float* dXT;   // [Cp][Bp]
    float* koutT;       // [H][Bp]            (forward)
    const float* gkT;   // [H][Bp]            (backward) dL/dk for this stage
    float* P;           // [n_hg][B][DFP]     (backward) per-h-group partial of dL/d(act)
    float* dW3acc;      // [n_bt][Np][DFP]    (backward) accumulated across stages
    float* db3acc;      // [n_bt][Np]
    float* gdXT;        // [Cp][Bp] or null   (backward) dL/d(dX/dt) partials are not supported yet
    const AdaptCtrl* ctrl;  // adaptive solver: skip all work once ctrl->done
};

struct AdvanceArgs {
    int B, Bp, H, method;
    float dt;
    const float* yT;
    float* ynewT;
    const float* kT[NCDE_MAX_STAGES];
    int n_emit;
    float* emit_ptr[4];  // (B,H) row-major slices of z_out
    int emit_mode[4];
    float emit_slope[4];
    __nv_bfloat16* ybf;  // all-tensor-core path: bf16 row-major [Bp][128] copy of y_{n+1} (input record of the next stage) or null
};

struct HiddenBwdArgs {
    int B, Bp, H, R, F, Dmax, DFP, n_hg;
    int D[NCDE_MAX_LAYERS + 1];
    int act[NCDE_MAX_LAYERS];
    const float* W[NCDE_MAX_LAYERS];       // row-major [D[l+1]][ldi[l]] (packed copy, ldi = D[l] padded to 4)
    int ldi[NCDE_MAX_LAYERS];
    int w_in_smem;
    int wsm_off[NCDE_MAX_LAYERS];
    int wsm_floats;
    const float* P;                         // [n_hg][B][DFP]
    float* dz_out;                          // adjoint: write dL/d(stage input) here [H][Bp] instead of the RK update
    const AdaptCtrl* ctrl;                  // adaptive adjoint: no-op once ctrl->done
    const float* actT[NCDE_MAX_LAYERS + 1]; // saved activations of this stage
    float* dpreT[NCDE_MAX_LAYERS];          // [D[l+1] pad4][Bp] scratch, consumed by hidden_wgrad
    float* gyT;                             // [H][Bp]  += dzs
    int n_k;                                // number of earlier-stage k gradients to update
    float* gkT[NCDE_MAX_STAGES];
    float kcoef[NCDE_MAX_STAGES];           // gk_j += kcoef[j] * dzs
};

constexpr int kWgTile = 64;    // weight-gradient CTA tile (out x in)
constexpr int kWgRows = 64;    // batch rows per staging chunk

struct WgradArgs {
    int B, Bp, n_slots, n_split, rows_per_split;
    int tile_begin[NCDE_MAX_LAYERS + 1];
    int Dout[NCDE_MAX_LAYERS], Din[NCDE_MAX_LAYERS];
    int n_lay[NCDE_MAX_LAYERS];
    int lay[NCDE_MAX_LAYERS][NCDE_MAX_LAYERS];
    int n_stage;                                            // RK stages folded into this launch
    const float* dpreT[NCDE_MAX_STAGES][NCDE_MAX_LAYERS];   // per stage, per layer
    const float* actT[NCDE_MAX_STAGES][NCDE_MAX_LAYERS];    // per stage, per layer: its input
    float* gWp[NCDE_MAX_LAYERS];          // per slot: [n_split][Dout][Din] partial accumulators (this CTA's split only)
    float* gbp[NCDE_MAX_LAYERS];          // per slot: [n_split][Dout]
    float* gW[NCDE_MAX_LAYERS];           // per slot: caller's gradient (torch layout), used by the final reduction
    float* gb[NCDE_MAX_LAYERS];           // per slot, nullable
    const AdaptCtrl* ctrl;                // adaptive adjoint: no-op once ctrl->done
};

// ---------------------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == NCDE_ACT_RELU) return v > 0.f ? v : 0.f;   // clamp_min(0): NaN stays NaN either way is irrelevant here
    if (act == NCDE_ACT_TANH) return tanhf(v);
    return v;
}
__device__ __forceinline__ float act_grad_from_output(float out, int act) {
    if (act == NCDE_ACT_RELU) return out > 0.f ? 1.f : 0.f;
    if (act == NCDE_ACT_TANH) return 1.f - out * out;
    return 1.f;
}

enum { COMBINE_Y = 0, COMBINE_RK4_S2 = 1, COMBINE_RK4_S3 = 2, COMBINE_RK4_S4 = 3, COMBINE_LINEAR = 4 };

// Stage input of the RK scheme with the reference's operation order and no FMA contraction
// (torchdiffeq/_impl/rk_common.py:106-114; _one_third / _two_thirds are rounded to fp32 by the tensor multiply).
// increment of an RK stage input from the earlier stage derivatives (reference operation order, no FMA contraction)
__device__ __forceinline__ float stage_increment(int combine, float dt, const float* const* kT, const float* coef, int64_t off) {
    const float third = 0.3333333432674408f;  // float32(1/3)
    switch (combine) {
        case COMBINE_Y: return 0.f;
        case COMBINE_RK4_S2:  // dt * k1 * (1/3)
            return __fmul_rn(__fmul_rn(dt, kT[0][off]), third);
        case COMBINE_RK4_S3:  // dt * (k2 - k1 * (1/3))
            return __fmul_rn(dt, __fsub_rn(kT[1][off], __fmul_rn(kT[0][off], third)));
        case COMBINE_RK4_S4:  // dt * (k1 - k2 + k3)
            return __fmul_rn(dt, __fadd_rn(__fsub_rn(kT[0][off], kT[1][off]), kT[2][off]));
        default: {            // sum_j k_j * coef_j
            float acc = 0.f;
            for (int j = 0; j < NCDE_MAX_STAGES; ++j)
                if (coef[j] != 0.f) acc = fmaf(kT[j][off], coef[j], acc);
            return acc;
        }
    }
}
__device__ __forceinline__ float combine_stage_input(const HiddenFwdArgs& a, int64_t off, const float* coef) {
    const float y = a.yT[off];
    if (a.combine == COMBINE_Y) return y;
    const float inc = stage_increment(a.combine, a.dt, a.kT, coef, off);
    return a.comb_sign < 0.f ? __fsub_rn(y, inc) : __fadd_rn(y, inc);
}

// generic elementwise pieces of the augmented (adjoint) RK scheme -------------------------------------------------
struct AugCombineArgs {
    int64_t n;
    int combine;
    float dt, sign;
    const float* base;
    const float* k[NCDE_MAX_STAGES];
    float coef[NCDE_MAX_STAGES];
    float* out;
    const AdaptCtrl* ctrl;   // adaptive: coefficients from ctrl->tab[tab_index]; no-op once ctrl->done
    int tab_index;
};
// out = base + sign * increment(combine, k...)
__global__ void aug_combine_kernel(const __grid_constant__ AugCombineArgs a0) {
    pdl_trigger();
    pdl_wait();
    AugCombineArgs a = a0;
    if (a.ctrl) {
        if (a.ctrl->done) return;
#pragma unroll
        for (int j = 0; j < NCDE_MAX_STAGES; ++j) a.coef[j] = a.ctrl->tab[a.tab_index].coef[j];
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const float b = a.base[i];
        if (a.combine == COMBINE_Y) { a.out[i] = b; continue; }
        const float inc = stage_increment(a.combine, a.dt, a.k, a.coef, i);
        a.out[i] = a.sign < 0.f ? __fsub_rn(b, inc) : __fadd_rn(b, inc);
    }
}
// s += sign * (k1 + 3 (k2 + k3) + k4) * dt * 0.125   (rk4)   |   s += sign * dt * k1   (euler)
__global__ void aug_advance_kernel(float* __restrict__ s, const float* k0, const float* k1, const float* k2, const float* k3,
                                   int method, float dt, float sign, int64_t n) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float inc;
        if (method == NCDE_RK4_38) {
            float q = __fadd_rn(k0[i], __fmul_rn(3.f, __fadd_rn(k1[i], k2[i])));
            q = __fadd_rn(q, k3[i]);
            inc = __fmul_rn(__fmul_rn(q, dt), 0.125f);
        } else {
            inc = __fmul_rn(dt, k0[i]);
        }
        s[i] = sign < 0.f ? __fsub_rn(s[i], inc) : __fadd_rn(s[i], inc);
    }
}

// dX/dt for one (row, channel) at the stage time — LinearInterpolation.derivative / NaturalCubicSpline.derivative
// (torchcde/interpolation_linear.py:231-234, interpolation_cubic.py:331-336)
__device__ __forceinline__ float path_derivative(const PathArgs& p, int idx, float frac, int64_t b, int c, int C) {
    if (p.kind == NCDE_PATH_LINEAR) {
        if (p.match && idx > 0 && frac < p.match_eps) {
            // gradient-matching region after an interior knot (src/ncde/interpolation.py:72-143): derivative of the polynomial
            const float* m = p.match + (((int64_t)b * (p.K - 2) + idx - 1) * C + c) * p.match_terms;
            float acc = 0.f;
            for (int i = 0; i < p.match_terms - 1; ++i) {
                const int pw = p.match_terms - 1 - i;
                float tp = 1.f;
                for (int q = 0; q < pw - 1; ++q) tp = __fmul_rn(tp, frac);
                acc = __fadd_rn(acc, __fmul_rn(m[i], __fmul_rn((float)pw, tp)));
            }
            return acc;
        }
        if (p.derivs) return p.derivs[((int64_t)b * (p.K - 1) + idx) * C + c];
        const float* cs = p.coeffs + (int64_t)b * p.K * C;
        return __fdiv_rn(__fsub_rn(cs[(int64_t)(idx + 1) * C + c], cs[(int64_t)idx * C + c]),
                         __fsub_rn(p.knots[idx + 1], p.knots[idx]));
    }
    const float* row = p.coeffs + ((int64_t)b * (p.K - 1) + idx) * 4 * C;
    const float bb = row[C + c], two_c = row[2 * C + c], three_d = row[3 * C + c];
    const float inner = __fadd_rn(two_c, __fmul_rn(three_d, frac));
    return __fadd_rn(bb, __fmul_rn(inner, frac));
}

// X(t) for one (row, channel) — LinearInterpolation.evaluate / NaturalCubicSpline.evaluate in the reference's operation
// order (torchcde/interpolation_linear.py:221-229, interpolation_cubic.py:324-329); un-smoothed paths only
__device__ __forceinline__ float path_value(const PathArgs& p, int idx, float frac, int64_t b, int c, int C) {
    if (p.kind == NCDE_PATH_LINEAR) {
        const float* cs = p.coeffs + (int64_t)b * p.K * C;
        const float lo = cs[(int64_t)idx * C + c], hi = cs[(int64_t)(idx + 1) * C + c];
        const float width = __fsub_rn(p.knots[idx + 1], p.knots[idx]);
        return __fadd_rn(lo, __fdiv_rn(__fmul_rn(frac, __fsub_rn(hi, lo)), width));
    }
    const float* row = p.coeffs + ((int64_t)b * (p.K - 1) + idx) * 4 * C;
    const float aa = row[c], bb = row[C + c], two_c = row[2 * C + c], three_d = row[3 * C + c];
    float inner = __fadd_rn(__fmul_rn(0.5f, two_c), __fdiv_rn(__fmul_rn(three_d, frac), 3.f));
    inner = __fadd_rn(bb, __fmul_rn(inner, frac));
    return __fadd_rn(aa, __fmul_rn(inner, frac));
}

// the constant "path derivative" of the one-channel contraction that vector_field_type evaluate / derivative run through the
// field kernels with: channel 0 = 1, padding channels = 0
__global__ void fill_e0_kernel(float* __restrict__ e0, int Bp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * (int64_t)Bp) e0[i] = i < Bp ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------------------------
// dx_all: dX/dt at EVERY stage time of the fixed grid in one launch (it depends on the path only, not on the state,
// so it is taken off the sequential stage chain).  CTA (stage, 32-row tile): knot lookup once (bucketize - 1 clamp,
// interpolation_linear.py:216), gather / Horner, transpose through shared memory so that both the read (channels
// contiguous) and the write (batch contiguous, feature-major) are coalesced.
// ---------------------------------------------------------------------------------------------------------------
struct DxAllArgs {
    int B, Bp, C, Cp;
    PathArgs path;             // path.t unused
    const float* stage_t;      // device [n_stages_total]
    float* dx_base;            // first stage's dXT
    size_t stage_stride;       // floats between consecutive stages' dXT
    int row_major;             // tensor-core path: [Bp][Cp] instead of [Cp][Bp]
    int value;                 // X(t) instead of dX/dt(t) (vector_field_type evaluate)
};

__global__ void __launch_bounds__(256) dx_all_kernel(const __grid_constant__ DxAllArgs a) {
    __shared__ float tile[32][129];
    __shared__ int s_idx;
    __shared__ float s_frac;
    const int st = blockIdx.x;
    const int64_t b0 = (int64_t)blockIdx.y * 32;
    const int tid = threadIdx.x;
    if (tid == 0) {
        const float t = a.stage_t[st];
        const int idx = knot_index<float>(a.path.knots, a.path.K, t);
        s_idx = idx;
        s_frac = __fsub_rn(t, a.path.knots[idx]);
    }
    __syncthreads();
    const int idxk = s_idx;
    const float frac = s_frac;
    float* out = a.dx_base + (size_t)st * a.stage_stride;
    for (int c0 = 0; c0 < a.Cp; c0 += 128) {
        const int cw = min(128, a.Cp - c0);
        for (int i = tid; i < 32 * cw; i += 256) {
            const int r = i / cw, c = c0 + i % cw;
            const int64_t b = b0 + r;
            float v = 0.f;
            if (b < a.B && c < a.C) v = a.value ? path_value(a.path, idxk, frac, b, c, a.C) : path_derivative(a.path, idxk, frac, b, c, a.C);
            if (a.row_major) { if (b < a.B) out[(size_t)b * a.Cp + c] = v; }
            else tile[r][c - c0] = v;
        }
        if (a.row_major) continue;
        __syncthreads();
        for (int i = tid; i < 32 * cw; i += 256) {
            const int c = i / 32, r = i % 32;
            const int64_t b = b0 + r;
            if (b < a.B) out[(size_t)(c0 + c) * a.Bp + b] = tile[r][c];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------------------------
// hidden layer: WT[k*ld + o] = W[o*Din + k]; bp[o] = bias[o] (0 when absent / padded)
__global__ void pack_hidden_kernel(const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ WT,
                                   float* __restrict__ bp, float* __restrict__ WR, int Dout, int Din, int ld, int ldi) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < Din * ld) {
        int k = idx / ld, o = idx % ld;
        WT[idx] = (o < Dout) ? W[(int64_t)o * Din + k] : 0.f;
    }
    if (WR && idx < Dout * ldi) {
        int o = idx / ldi, i = idx % ldi;
        WR[idx] = (i < Din) ? W[(int64_t)o * Din + i] : 0.f;
    }
    if (idx < ld) bp[idx] = (bias && idx < Dout) ? bias[idx] : 0.f;
}

// asynchronous 16-byte global->shared copy
__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// final layer: n = h*Cp + c (c padded to Cp, h padded to n_hg*Hg); W3T[k*Np + n], W3R[n*DFP + k], b3p[n].
// Gated final layer (Wg != null; MinimalGatedVectorField, src/ncde/vector_fields/gating.py:7-32): the sigmoid head and the tanh
// head are interleaved column by column, n = h*Cp + 2c + {0: sigmoid head Wg, 1: tanh head W}, so that one thread of the
// field kernels owns both pre-activations of a (h, c) entry and the GEMMs themselves are unchanged.
__global__ void pack_final_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ Wg,
                                  const float* __restrict__ biasg, float* __restrict__ W3T,
                                  float* __restrict__ W3R, float* __restrict__ b3p, int H, int C, int Cp, int DF,
                                  int DFP, int Np) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)Np * DFP;
    if (idx < total) {
        int n = (int)(idx / DFP), k = (int)(idx % DFP);
        int h = n / Cp, c = n % Cp;
        const float* src = W;
        if (Wg) { src = (c & 1) ? W : Wg; c >>= 1; }
        float v = (h < H && c < C && k < DF) ? src[((int64_t)h * C + c) * DF + k] : 0.f;
        if (W3R) W3R[idx] = v;
        if (k < DF) W3T[(int64_t)k * Np + n] = v;
    }
    if (idx < Np) {
        int h = (int)idx / Cp, c = (int)idx % Cp;
        const float* src = bias;
        if (Wg) { src = (c & 1) ? bias : biasg; c >>= 1; }
        b3p[idx] = (src && h < H && c < C) ? src[(int64_t)h * C + c] : 0.f;
    }
}

// Final layer packed with the roles of h and c exchanged: n' = c*Hp + h (h padded to Hp, c padded to n_cg*Cg rows-of-Hp).
// With this packing field_fwd_kernel, called with (H', C', dX') = (C, H, gk), computes
//     gdX[b, c] = sum_h tanh( act[b,:] . W3[(h,c),:] + b3[(h,c)] ) * gk[b, h],
// the gradient of the f(z).dX/dt contraction (torchcde/solver.py:132) w.r.t. dX/dt — the same code that computes k[b,h]
// in the forward pass, so the path gradient needs no kernel of its own and no atomics.
__global__ void pack_final_swapped_kernel(const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ W3T,
                                          float* __restrict__ b3p, int H, int C, int Hp, int DF, int Np) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)Np * DF;
    if (idx < total) {
        const int n = (int)(idx / DF), k = (int)(idx % DF);
        const int c = n / Hp, h = n % Hp;
        W3T[(int64_t)k * Np + n] = (h < H && c < C) ? W[((int64_t)h * C + c) * DF + k] : 0.f;
    }
    if (idx < Np) {
        const int c = (int)idx / Hp, h = (int)idx % Hp;
        b3p[idx] = (bias && h < H && c < C) ? bias[(int64_t)h * C + c] : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// path_grad: chain rule from dL/d(dX/dt) at the stage times of ONE solver step to the path coefficients, i.e. the
// backward of LinearInterpolation.derivative (derivs = (c[1:] - c[:-1]) / (t[1:] - t[:-1]) gathered at the knot index,
// torchcde/interpolation_linear.py:198,231-234) or of NaturalCubicSpline.derivative (b + (2c + 3d f) f,
// interpolation_cubic.py:331-336).  CTA = 32 batch rows; the feature-major gradient tile is transposed through shared
// memory so that reads (batch contiguous) and the read-modify-write of grad_coeffs (channels contiguous) are both
// coalesced.  One thread owns one (row, channel) for every stage of the step, launches are stream ordered: no atomics,
// deterministic.
// ---------------------------------------------------------------------------------------------------------------
struct PathGradArgs {
    int B, Bp, C, n_stage, kind, K;
    const float* knots;
    float t[NCDE_MAX_STAGES];
    const float* gdXT[NCDE_MAX_STAGES];  // [C][Bp]
    float* grad_coeffs;                  // LINEAR (B,K,C) | CUBIC (B,K-1,4C)
};

__global__ void __launch_bounds__(256) path_grad_kernel(const __grid_constant__ PathGradArgs a) {
    __shared__ float tile[32][129];
    __shared__ int s_idx[NCDE_MAX_STAGES];
    __shared__ float s_frac[NCDE_MAX_STAGES], s_width[NCDE_MAX_STAGES];
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * 32;
    if (tid < a.n_stage) {
        const float t = a.t[tid];
        const int idx = knot_index<float>(a.knots, a.K, t);
        s_idx[tid] = idx;
        s_frac[tid] = __fsub_rn(t, a.knots[idx]);
        s_width[tid] = __fsub_rn(a.knots[idx + 1], a.knots[idx]);
    }
    __syncthreads();
    for (int c0 = 0; c0 < a.C; c0 += 128) {
        const int cw = min(128, a.C - c0);
        for (int st = 0; st < a.n_stage; ++st) {
            const float* __restrict__ src = a.gdXT[st];
            for (int i = tid; i < cw * 32; i += 256) {
                const int c = i / 32, r = i % 32;
                const int64_t b = b0 + r;
                tile[r][c] = b < a.B ? src[(size_t)(c0 + c) * a.Bp + b] : 0.f;
            }
            __syncthreads();
            const int idx = s_idx[st];
            const float frac = s_frac[st], width = s_width[st];
            for (int i = tid; i < 32 * cw; i += 256) {
                const int r = i / cw, c = c0 + i % cw;
                const int64_t b = b0 + r;
                if (b >= a.B) continue;
                const float g = tile[r][c - c0];
                if (a.kind == NCDE_PATH_LINEAR) {
                    const float q = __fdiv_rn(g, width);
                    float* p = a.grad_coeffs + ((size_t)b * a.K + idx) * a.C + c;
                    p[0] = __fsub_rn(p[0], q);
                    p[a.C] = __fadd_rn(p[a.C], q);
                } else {
                    float* row = a.grad_coeffs + ((size_t)b * (a.K - 1) + idx) * 4 * a.C;
                    const float gi = __fmul_rn(g, frac);
                    row[a.C + c] = __fadd_rn(row[a.C + c], g);
                    row[2 * a.C + c] = __fadd_rn(row[2 * a.C + c], gi);
                    row[3 * a.C + c] = __fadd_rn(row[3 * a.C + c], __fmul_rn(gi, frac));
                }
            }
            __syncthreads();
        }
    }
}

// (B,H) row-major <-> [H][Bp] feature-major
__global__ void to_feature_major_kernel(const float* __restrict__ src, float* __restrict__ dstT, int B, int Bp, int H) {
    __shared__ float tile[32][33];
    int b0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int b = b0 + i, h = h0 + threadIdx.x;
        tile[i][threadIdx.x] = (b < B && h < H) ? src[(int64_t)b * H + h] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int h = h0 + i, b = b0 + threadIdx.x;
        if (h < H && b < Bp) dstT[(int64_t)h * Bp + b] = tile[threadIdx.x][i];
    }
}

// dst (B,H) row-major  (=|+=)  scale * srcT [H][Bp]   (+ add (B,H) if given)
__global__ void from_feature_major_kernel(const float* __restrict__ srcT, const float* __restrict__ add,
                                          float* __restrict__ dst, int B, int Bp, int H) {
    __shared__ float tile[32][33];
    int b0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int h = h0 + i, b = b0 + threadIdx.x;
        tile[i][threadIdx.x] = (h < H && b < B) ? srcT[(int64_t)h * Bp + b] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int b = b0 + i, h = h0 + threadIdx.x;
        if (b < B && h < H) {
            float v = tile[threadIdx.x][i];
            if (add) v += add[(int64_t)b * H + h];
            dst[(int64_t)b * H + h] = v;
        }
    }
}

// gyT[h][b] += scale * g[b][h]     (output-gradient injection in the backward sweep)
__global__ void add_out_grad_kernel(float* __restrict__ gyT, const float* __restrict__ g, float scale, int B, int Bp,
                                    int H) {
    __shared__ float tile[32][33];
    pdl_trigger();
    pdl_wait();
    int b0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int b = b0 + i, h = h0 + threadIdx.x;
        tile[i][threadIdx.x] = (b < B && h < H) ? g[(int64_t)b * H + h] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int h = h0 + i, b = b0 + threadIdx.x;
        if (h < H && b < B) gyT[(int64_t)h * Bp + b] += scale * tile[threadIdx.x][i];
    }
}

// acc[0..3] += sum_k w[k * ldk] * x[k * R .. +3]: one output feature for four batch rows.  Loads are issued in batches
// of 8 before the FMAs, two accumulator sets halve the dependent chain.  SMEM selects shared vs read-only global
// weights at compile time (a runtime pointer select would degrade every load to a generic LD).
template <bool SMEM>
__device__ __forceinline__ float4 matvec4(const float* __restrict__ w, int ldk, const float* __restrict__ x, int R, int n,
                                          float4 acc) {
    float4 acc2 = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0;
    for (; k + 8 <= n; k += 8) {
        float wv[8];
        float4 xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            wv[u] = SMEM ? w[(size_t)(k + u) * ldk] : __ldg(w + (size_t)(k + u) * ldk);
            xv[u] = *reinterpret_cast<const float4*>(x + (k + u) * R);
        }
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
            acc.x = fmaf(wv[u], xv[u].x, acc.x); acc.y = fmaf(wv[u], xv[u].y, acc.y);
            acc.z = fmaf(wv[u], xv[u].z, acc.z); acc.w = fmaf(wv[u], xv[u].w, acc.w);
            acc2.x = fmaf(wv[u + 1], xv[u + 1].x, acc2.x); acc2.y = fmaf(wv[u + 1], xv[u + 1].y, acc2.y);
            acc2.z = fmaf(wv[u + 1], xv[u + 1].z, acc2.z); acc2.w = fmaf(wv[u + 1], xv[u + 1].w, acc2.w);
        }
    }
    for (; k < n; ++k) {
        const float wk = SMEM ? w[(size_t)k * ldk] : __ldg(w + (size_t)k * ldk);
        const float4 xk = *reinterpret_cast<const float4*>(x + k * R);
        acc.x = fmaf(wk, xk.x, acc.x); acc.y = fmaf(wk, xk.y, acc.y);
        acc.z = fmaf(wk, xk.z, acc.z); acc.w = fmaf(wk, xk.w, acc.w);
    }
    return make_float4(acc.x + acc2.x, acc.y + acc2.y, acc.z + acc2.z, acc.w + acc2.w);
}

// ---------------------------------------------------------------------------------------------------------------
// hidden_fwd: batch-split.  One CTA owns R rows: forms the RK stage input, evaluates dX/dt at the stage time
// (knot lookup + gather / Horner) and runs the hidden Linear+activation layers.  Everything is written
// feature-major for the field kernel and for the backward pass.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) hidden_fwd_kernel(const __grid_constant__ HiddenFwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int R = a.R;
    const int64_t b0 = (int64_t)blockIdx.x * R;
    float* buf0 = sm;
    float* buf1 = sm + (size_t)a.Dmax * R;
    float* wsm = buf1 + (size_t)a.Dmax * R;
    if (a.w_in_smem) {
        // every distinct hidden weight matrix is copied once (asynchronously, overlapping steps 1-2 below); the packed
        // matrices of the hidden layers are contiguous in the workspace, so this is one flat copy
        const float* src = a.WT[0];
        for (int i = tid * 4; i < a.wsm_floats; i += kThreads * 4) cp_async_16(wsm + i, src + i);
    }
    pdl_trigger();
    pdl_wait();  // y and k_i come from the previous kernels
    float coef[NCDE_MAX_STAGES];
    float t_stage = a.path.t;
    if (a.ctrl) {
        if (a.ctrl->done) return;
        const StageTab& tb = a.ctrl->tab[a.tab_index];
        t_stage = tb.t;
#pragma unroll
        for (int j = 0; j < NCDE_MAX_STAGES; ++j) coef[j] = tb.coef[j];
    } else {
#pragma unroll
        for (int j = 0; j < NCDE_MAX_STAGES; ++j) coef[j] = a.coef[j];
    }

    // 1. stage input  zs[h][r]
    for (int idx = tid; idx < a.H * R; idx += kThreads) {
        const int h = idx / R, r = idx % R;
        const int64_t b = b0 + r;
        float v = 0.f;
        if (b < a.B) {
            const int64_t off = (int64_t)h * a.Bp + b;
            v = combine_stage_input(a, off, coef);
            a.actT[0][off] = v;
        }
        buf0[h * R + r] = v;
    }
    if (a.n_u && !a.uT) {
        // continuous adjoint: the stage times are not known ahead of the launch loop's state, evaluate X(t) / dX/dt(t) in place
        const int idxk = knot_index<float>(a.path.knots, a.path.K, t_stage);
        const float frac = __fsub_rn(t_stage, a.path.knots[idxk]);
        float* __restrict__ dst = a.actT[0] + (size_t)a.H * a.Bp;
        for (int idx = tid; idx < R * a.n_u; idx += kThreads) {
            const int r = idx / a.n_u, c = idx % a.n_u;
            const int64_t b = b0 + r;
            float v = 0.f;
            if (b < a.B) {
                v = a.vf == NCDE_VF_EVALUATE ? path_value(a.path, idxk, frac, b, c, a.n_u) : path_derivative(a.path, idxk, frac, b, c, a.n_u);
                dst[(size_t)c * a.Bp + b] = v;
            }
            buf0[(a.H + c) * R + r] = v;
        }
    } else if (a.n_u) {
        float* __restrict__ dst = a.actT[0] + (size_t)a.H * a.Bp;
        for (int idx = tid; idx < a.n_u * R; idx += kThreads) {
            const int c = idx / R, r = idx % R;
            const int64_t b = b0 + r;
            float v = 0.f;
            if (b < a.B) {
                v = a.uT[(size_t)c * a.Bp + b];
                if (dst != a.uT) dst[(size_t)c * a.Bp + b] = v;
            }
            buf0[(a.H + c) * R + r] = v;
        }
    }
    // 2. dX/dt at the stage time -> dXT[c][b] (zero in the padded channels); skipped when dx_all_kernel already did it
    if (a.dXT) {
        const int idxk = knot_index<float>(a.path.knots, a.path.K, t_stage);
        const float frac = __fsub_rn(t_stage, a.path.knots[idxk]);
        float* tmp = buf1;  // [Cp][R]
        for (int idx = tid; idx < R * a.Cp; idx += kThreads) {
            const int r = idx / a.Cp, c = idx % a.Cp;
            const int64_t b = b0 + r;
            float v = 0.f;
            if (b < a.B && c < a.C) v = path_derivative(a.path, idxk, frac, b, c, a.C);
            if (a.dx_row_major) { if (b < a.B) a.dXT[(int64_t)b * a.Cp + c] = v; }
            else tmp[c * R + r] = v;
        }
        if (!a.dx_row_major) {
            __syncthreads();
            for (int idx = tid; idx < R * a.Cp; idx += kThreads) {
                const int c = idx / R, r = idx % R;
                const int64_t b = b0 + r;
                if (b < a.B) a.dXT[(int64_t)c * a.Bp + b] = tmp[c * R + r];
            }
        }
        if (a.ddXT) {
            // d/dt of NaturalCubicSpline.derivative (interpolation_cubic.py:331-336): 2c + 2 (3d) frac; zero for linear paths
            for (int idx = tid; idx < R * a.Cp; idx += kThreads) {
                const int r = idx / a.Cp, c = idx % a.Cp;
                const int64_t b = b0 + r;
                float v = 0.f;
                if (b < a.B && c < a.C && a.path.kind != NCDE_PATH_LINEAR) {
                    const float* row = a.path.coeffs + ((int64_t)b * (a.path.K - 1) + idxk) * 4 * a.C;
                    v = __fadd_rn(row[2 * a.C + c], __fmul_rn(2.f, __fmul_rn(row[3 * a.C + c], frac)));
                }
                if (b < a.B) a.ddXT[a.dx_row_major ? (int64_t)b * a.Cp + c : (int64_t)c * a.Bp + b] = v;
            }
        }
    }
    if (a.w_in_smem) cp_async_wait_all_();
    __syncthreads();
    // 3. hidden layers
    const int RQ = R / 4;
    for (int l = 0; l < a.F; ++l) {
        const float* in = (l & 1) ? buf1 : buf0;
        float* out = (l & 1) ? buf0 : buf1;
        const int Din = a.D[l], Dout = a.D[l + 1], ld = a.ldw[l];
        const float* __restrict__ bp = a.bp[l];
        for (int item = tid; item < Dout * RQ; item += kThreads) {
            const int o = item % Dout, q = item / Dout;
            const float bias = bp[o];
            float4 acc = make_float4(bias, bias, bias, bias);
            if (a.w_in_smem) acc = matvec4<true>(wsm + a.wsm_off[l] + o, ld, in + q * 4, R, Din, acc);
            else acc = matvec4<false>(a.WT[l] + o, ld, in + q * 4, R, Din, acc);
            const int act = a.act[l];
            if (act == NCDE_ACT_GATE_IN) {
                // rows [0, Din) pass the input through (identity block of W), rows [Din, 2 Din) are sigmoid(W_f x + b_f) * x
                if (o >= Din) {
                    const float4 xi = *reinterpret_cast<const float4*>(in + (o - Din) * R + q * 4);
                    acc.x = sigmoidf_(acc.x) * xi.x; acc.y = sigmoidf_(acc.y) * xi.y;
                    acc.z = sigmoidf_(acc.z) * xi.z; acc.w = sigmoidf_(acc.w) * xi.w;
                }
            } else {
                acc.x = apply_act(acc.x, act); acc.y = apply_act(acc.y, act);
                acc.z = apply_act(acc.z, act); acc.w = apply_act(acc.w, act);
            }
            *reinterpret_cast<float4*>(out + o * R + q * 4) = acc;
        }
        __syncthreads();
        float* __restrict__ g = a.actT[l + 1];
        for (int idx = tid; idx < Dout * R; idx += kThreads) {
            const int o = idx / R, r = idx % R;
            const int64_t b = b0 + r;
            if (b < a.B) g[(int64_t)o * a.Bp + b] = out[o * R + r];
        }
        // `out` is read-only until the next layer finishes writing the other buffer: no extra barrier needed
    }
    if (a.abf) {
        // bf16 row-major copy [b][k] (k padded with zeros to KP) for the tensor-core final layer
        const float* last = (a.F & 1) ? buf1 : buf0;
        const int DF = a.D[a.F];
        for (int idx = tid; idx < R * a.KP; idx += kThreads) {
            const int r = idx / a.KP, k = idx % a.KP;
            const int64_t b = b0 + r;
            if (b < a.B) a.abf[(size_t)b * a.KP + k] = __float2bfloat16(k < DF ? last[k * R + r] : 0.f);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// field_fwd: weight-stationary.  CTA (g, bt) keeps the W3 slice of h-group g in shared memory and streams the
// rows of batch tile bt through it in 64-row chunks:
//     k[b, h] = sum_c tanh( act[b,:] . W3[(h,c),:] + b3[(h,c)] ) * dX[b, c]
// The (B, H*C) matrix of the reference (src/ncde/vector_fields/base.py:99-104, torchcde/solver.py:132) is never
// materialised.  Thread (nt, mt) owns 4 consecutive n (= 4 channels of one h) x TM rows.
// ---------------------------------------------------------------------------------------------------------------
template <int TM, bool GATED = false>
__global__ void __launch_bounds__(kThreads) field_fwd_kernel(const __grid_constant__ FieldArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int S = a.S, DF = a.DF;
    const int NT = S / 4;
    constexpr int MT = kChunk / TM;
    const int nt = tid % NT, mt = tid / NT;
    const bool active = mt < MT;
    const int g = blockIdx.x, bt = blockIdx.y;
    float* Ws = sm;                       // [DF][S]
    float* As = Ws + (size_t)DF * S;      // [DF][64]
    float* red = As + (size_t)DF * kChunk;  // [NT][64]

    for (int idx = tid; idx < DF * NT; idx += kThreads) {
        const int k = idx / NT, j = idx % NT;
        reinterpret_cast<float4*>(Ws)[idx] =
            __ldg(reinterpret_cast<const float4*>(a.W3T + (size_t)k * a.Np + (size_t)g * S) + j);
    }
    float bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = active ? a.b3p[(size_t)g * S + nt * 4 + j] : 0.f;
    const int c0 = (nt * 4) % a.Cp;
    pdl_trigger();
    pdl_wait();  // activations come from hidden_fwd
    if (a.ctrl && a.ctrl->done) return;

    const int64_t row_begin = (int64_t)bt * a.Bt;
    const int64_t row_end = min((int64_t)a.B, row_begin + a.Bt);
    for (int64_t b0 = row_begin; b0 < row_end; b0 += kChunk) {
        for (int idx = tid; idx < DF * (kChunk / 4); idx += kThreads) {
            const int k = idx / (kChunk / 4), j = idx % (kChunk / 4);
            reinterpret_cast<float4*>(As)[idx] =
                __ldg(reinterpret_cast<const float4*>(a.actT + (size_t)k * a.Bp + b0) + j);
        }
        __syncthreads();
        if (active) {
            float acc[TM][4];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
            for (int k = 0; k < DF; ++k) {
                const float4 w = *reinterpret_cast<const float4*>(Ws + k * S + nt * 4);
                float av[TM];
#pragma unroll
                for (int i4 = 0; i4 < TM / 4; ++i4) {
                    const float4 x = *reinterpret_cast<const float4*>(As + k * kChunk + mt * TM + i4 * 4);
                    av[i4 * 4 + 0] = x.x; av[i4 * 4 + 1] = x.y; av[i4 * 4 + 2] = x.z; av[i4 * 4 + 3] = x.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    acc[i][0] = fmaf(av[i], w.x, acc[i][0]); acc[i][1] = fmaf(av[i], w.y, acc[i][1]);
                    acc[i][2] = fmaf(av[i], w.z, acc[i][2]); acc[i][3] = fmaf(av[i], w.w, acc[i][3]);
                }
            }
            // epilogue: tanh, multiply by dX/dt, reduce this thread's 4 channels
            float part[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) part[i] = 0.f;
            if (GATED) {
                // columns (2c, 2c+1) = (sigmoid head, tanh head) of entry (h, c): M = sigmoid(z) * tanh(r)  (gating.py:30-32)
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) {
                    const float* dxrow = a.dXT + (size_t)(c0 / 2 + pr) * a.Bp + b0 + mt * TM;
#pragma unroll
                    for (int i = 0; i < TM; ++i) {
                        const float mv = sigmoidf_(acc[i][2 * pr] + bias[2 * pr]) * tanhf(acc[i][2 * pr + 1] + bias[2 * pr + 1]);
                        part[i] = fmaf(mv, __ldg(dxrow + i), part[i]);
                    }
                }
            } else
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* dxrow = a.dXT + (size_t)(c0 + j) * a.Bp + b0 + mt * TM;
#pragma unroll
                for (int i4 = 0; i4 < TM / 4; ++i4) {
                    const float4 dx = __ldg(reinterpret_cast<const float4*>(dxrow) + i4);
                    part[i4 * 4 + 0] = fmaf(tanhf(acc[i4 * 4 + 0][j] + bias[j]), dx.x, part[i4 * 4 + 0]);
                    part[i4 * 4 + 1] = fmaf(tanhf(acc[i4 * 4 + 1][j] + bias[j]), dx.y, part[i4 * 4 + 1]);
                    part[i4 * 4 + 2] = fmaf(tanhf(acc[i4 * 4 + 2][j] + bias[j]), dx.z, part[i4 * 4 + 2]);
                    part[i4 * 4 + 3] = fmaf(tanhf(acc[i4 * 4 + 3][j] + bias[j]), dx.w, part[i4 * 4 + 3]);
                }
            }
#pragma unroll
            for (int i4 = 0; i4 < TM / 4; ++i4)
                *reinterpret_cast<float4*>(red + nt * kChunk + mt * TM + i4 * 4) =
                    make_float4(part[i4 * 4], part[i4 * 4 + 1], part[i4 * 4 + 2], part[i4 * 4 + 3]);
        }
        __syncthreads();
        const int Cq = a.Cp / 4;
        for (int idx = tid; idx < a.Hg * kChunk; idx += kThreads) {
            const int hl = idx / kChunk, m = idx % kChunk;
            float s = 0.f;
            for (int q = 0; q < Cq; ++q) s += red[(hl * Cq + q) * kChunk + m];
            const int64_t b = b0 + m;
            const int h = g * a.Hg + hl;
            if (b < a.B && h < a.H) a.koutT[(size_t)h * a.Bp + b] = s;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// advance: y_{n+1} from the stage derivatives (fixed_grid.py:6-29, rk_common.py:114) + output emission
// (solvers.py:106-117,166-172).  Elementwise on feature-major state; emits (B,H) row-major slices of z_out.
// ---------------------------------------------------------------------------------------------------------------
__global__ void advance_kernel(const __grid_constant__ AdvanceArgs a) {
    __shared__ float t_old[32][33];
    __shared__ float t_new[32][33];
    pdl_trigger();
    pdl_wait();
    const int b0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int h = h0 + i, b = b0 + threadIdx.x;
        float y = 0.f, yn = 0.f;
        if (h < a.H && b < a.B) {
            const size_t off = (size_t)h * a.Bp + b;
            y = a.yT[off];
            if (a.method == NCDE_RK4_38) {
                // (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
                float s = __fadd_rn(a.kT[0][off], __fmul_rn(3.f, __fadd_rn(a.kT[1][off], a.kT[2][off])));
                s = __fadd_rn(s, a.kT[3][off]);
                yn = __fadd_rn(y, __fmul_rn(__fmul_rn(s, a.dt), 0.125f));
            } else {  // Euler: y0 + dt * f0
                yn = __fadd_rn(y, __fmul_rn(a.dt, a.kT[0][off]));
            }
            a.ynewT[off] = yn;
            if (a.ybf) a.ybf[(size_t)b * 128 + h] = __float2bfloat16(yn);
        }
        t_old[i][threadIdx.x] = y;
        t_new[i][threadIdx.x] = yn;
    }
    if (a.n_emit == 0) return;
    __syncthreads();
    for (int e = 0; e < a.n_emit; ++e) {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int b = b0 + i, h = h0 + threadIdx.x;
            if (b < a.B && h < a.H) {
                const float y = t_old[threadIdx.x][i], yn = t_new[threadIdx.x][i];
                float v;
                if (a.emit_mode[e] == 0) v = y;
                else if (a.emit_mode[e] == 1) v = yn;
                else v = __fadd_rn(y, __fmul_rn(a.emit_slope[e], __fsub_rn(yn, y)));  // _linear_interp
                a.emit_ptr[e][(size_t)b * a.H + h] = v;
            }
        }
    }
}

// backward: gk_i = c_i * dt * gy1   (derivative of the RK increment w.r.t. each stage derivative)
__global__ void rk_bwd_begin_kernel(const float* __restrict__ gyT, float* gk0, float* gk1, float* gk2, float* gk3,
                                    int method, float dt, int64_t n) {
    pdl_trigger();
    pdl_wait();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = gyT[i];
    if (method == NCDE_RK4_38) {
        const float c1 = dt * 0.125f, c3 = 3.f * (dt * 0.125f);
        gk0[i] = g * c1; gk1[i] = g * c3; gk2[i] = g * c3; gk3[i] = g * c1;
    } else {
        gk0[i] = g * dt;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// field_bwd: weight-stationary backward of the final layer + contraction, for one RK stage.
//   pre = act . W3^T + b3 ; T = tanh(pre) ; G[b,(h,c)] = gk[b,h] * dX[b,c] * (1 - T^2)
//   dW3[(h,c),:] += sum_b G * act[b,:]      (register accumulators across all chunks, then += into dW3acc)
//   db3[(h,c)]   += sum_b G
//   P_g[b,:]      = sum_{(h,c) in group} G * W3[(h,c),:]   (partial of dL/d act; summed over groups by hidden_bwd)
// ---------------------------------------------------------------------------------------------------------------
template <int TM, bool GATED = false>
__global__ void __launch_bounds__(kThreads, 1) field_bwd_kernel(const __grid_constant__ FieldArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int S = a.S, DF = a.DF, DFP = a.DFP;
    const int NT = S / 4;
    constexpr int MT = kChunk / TM;
    const int nt = tid % NT, mt = tid / NT;
    const bool gemm_active = mt < MT;
    const int g = blockIdx.x, bt = blockIdx.y;
    const int ARS = DFP + 4;  // row stride of the row-major activation tile

    float* Ws = sm;                              // [DF][S]        k-major slice
    float* Wr = Ws + (size_t)DF * S;             // [S][DFP]       n-major slice
    float* As = Wr + (size_t)S * DFP;            // [DF][64]       k-major chunk
    float* Ar = As + (size_t)DF * kChunk;        // [64][DFP+4]    row-major chunk
    float* Gn = Ar + (size_t)kChunk * ARS;       // [S][68]        n-major G
    float* Gm = Gn + (size_t)S * kGnStride;      // [64][S]        m-major G
    float* gks = Gm + (size_t)kChunk * S;        // [Hg][64]

    for (int idx = tid; idx < DF * NT; idx += kThreads) {
        const int k = idx / NT, j = idx % NT;
        reinterpret_cast<float4*>(Ws)[idx] =
            __ldg(reinterpret_cast<const float4*>(a.W3T + (size_t)k * a.Np + (size_t)g * S) + j);
    }
    for (int idx = tid; idx < S * (DFP / 4); idx += kThreads)
        reinterpret_cast<float4*>(Wr)[idx] =
            __ldg(reinterpret_cast<const float4*>(a.W3R + (size_t)g * S * DFP) + idx);
    // zero the K padding of Ar once (columns DF..DFP-1 never change afterwards)
    for (int idx = tid; idx < kChunk * ARS; idx += kThreads) Ar[idx] = 0.f;

    float bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = gemm_active ? a.b3p[(size_t)g * S + nt * 4 + j] : 0.f;
    const int c0 = (nt * 4) % a.Cp;
    const int hl_of_nt = (nt * 4) / a.Cp;

    // wgrad mapping: thread (nt, kw) owns 4 n x 16 k, k = q*(DFP/4) + kw*4 + j
    const int KW = DFP / 16;
    const int kw = tid / NT;
    const bool wg_active = kw < KW;
    float accw[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 16; ++q) accw[j][q] = 0.f;
    float accb[4] = {0.f, 0.f, 0.f, 0.f};

    // dgrad mapping: thread (kd, md) owns 4 m x 8 k per k-chunk, k = {kc*4+j} U {DFP/2 + kc*4 + j}, kc = kd, kd + 16, ...
    // (one chunk per thread up to DFP = 128; wider final-layer inputs loop)
    const int KD = DFP / 8;
    const int KDT = KD < 16 ? KD : 16;
    const int kd = tid % KDT, md = tid / KDT;
    const bool dg_active = md < kChunk / 4;
    pdl_trigger();
    pdl_wait();  // gk comes from the previous kernels
    if (a.ctrl && a.ctrl->done) return;

    const int64_t row_begin = (int64_t)bt * a.Bt;
    const int64_t row_end = min((int64_t)a.B, row_begin + a.Bt);
    for (int64_t b0 = row_begin; b0 < row_end; b0 += kChunk) {
        for (int idx = tid; idx < DF * (kChunk / 4); idx += kThreads) {
            const int k = idx / (kChunk / 4), j = idx % (kChunk / 4);
            float4 x = __ldg(reinterpret_cast<const float4*>(a.actT + (size_t)k * a.Bp + b0) + j);
            // batch padding holds uninitialised memory: zero it so that 0 * garbage cannot poison the weight gradient
            const int64_t bb = b0 + j * 4;
            if (bb + 0 >= a.B) x.x = 0.f;
            if (bb + 1 >= a.B) x.y = 0.f;
            if (bb + 2 >= a.B) x.z = 0.f;
            if (bb + 3 >= a.B) x.w = 0.f;
            reinterpret_cast<float4*>(As)[idx] = x;
        }
        for (int idx = tid; idx < a.Hg * kChunk; idx += kThreads) {
            const int hl = idx / kChunk, m = idx % kChunk;
            const int h = g * a.Hg + hl;
            const int64_t b = b0 + m;
            gks[idx] = (h < a.H && b < a.B) ? a.gkT[(size_t)h * a.Bp + b] : 0.f;
        }
        __syncthreads();
        // row-major copy of the chunk for the weight-gradient product
        for (int idx = tid; idx < DF * kChunk; idx += kThreads) {
            const int k = idx / kChunk, m = idx % kChunk;
            Ar[m * ARS + k] = As[idx];
        }
        if (gemm_active) {
            float acc[TM][4];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
            for (int k = 0; k < DF; ++k) {
                const float4 w = *reinterpret_cast<const float4*>(Ws + k * S + nt * 4);
                float av[TM];
#pragma unroll
                for (int i4 = 0; i4 < TM / 4; ++i4) {
                    const float4 x = *reinterpret_cast<const float4*>(As + k * kChunk + mt * TM + i4 * 4);
                    av[i4 * 4 + 0] = x.x; av[i4 * 4 + 1] = x.y; av[i4 * 4 + 2] = x.z; av[i4 * 4 + 3] = x.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    acc[i][0] = fmaf(av[i], w.x, acc[i][0]); acc[i][1] = fmaf(av[i], w.y, acc[i][1]);
                    acc[i][2] = fmaf(av[i], w.z, acc[i][2]); acc[i][3] = fmaf(av[i], w.w, acc[i][3]);
                }
            }
            // G = gk * dX * (1 - tanh^2), zero outside the batch
            if (GATED) {
                // M = s * t with s = sigmoid(z), t = tanh(r):  dz = gk dX t s (1 - s),  dr = gk dX s (1 - t^2)
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) {
                    const float* dxrow = a.dXT + (size_t)(c0 / 2 + pr) * a.Bp + b0 + mt * TM;
#pragma unroll
                    for (int i = 0; i < TM; ++i) {
                        const int m = mt * TM + i;
                        const float sg = sigmoidf_(acc[i][2 * pr] + bias[2 * pr]);
                        const float t = tanhf(acc[i][2 * pr + 1] + bias[2 * pr + 1]);
                        float dm = gks[hl_of_nt * kChunk + m] * __ldg(dxrow + i);
                        if (b0 + m >= a.B) dm = 0.f;
                        acc[i][2 * pr] = dm * t * sg * (1.f - sg);
                        acc[i][2 * pr + 1] = dm * sg * (1.f - t * t);
                    }
                }
            } else
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* dxrow = a.dXT + (size_t)(c0 + j) * a.Bp + b0 + mt * TM;
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const int m = mt * TM + i;
                    const float t = tanhf(acc[i][j] + bias[j]);
                    const float gk = gks[hl_of_nt * kChunk + m];
                    float gv = gk * __ldg(dxrow + i) * (1.f - t * t);
                    if (b0 + m >= a.B) gv = 0.f;
                    acc[i][j] = gv;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i4 = 0; i4 < TM / 4; ++i4)
                    *reinterpret_cast<float4*>(Gn + (nt * 4 + j) * kGnStride + mt * TM + i4 * 4) =
                        make_float4(acc[i4 * 4][j], acc[i4 * 4 + 1][j], acc[i4 * 4 + 2][j], acc[i4 * 4 + 3][j]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
                *reinterpret_cast<float4*>(Gm + (mt * TM + i) * S + nt * 4) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        // weight gradient: accw[j][q*4+jj] += G[m][n0+j] * act[m][k]
        if (wg_active) {
            const int QS = DFP / 4;
#pragma unroll 2
            for (int m = 0; m < kChunk; ++m) {
                const float4 gv = *reinterpret_cast<const float4*>(Gm + m * S + nt * 4);
                const float gj[4] = {gv.x, gv.y, gv.z, gv.w};
                if (kw == 0) { accb[0] += gv.x; accb[1] += gv.y; accb[2] += gv.z; accb[3] += gv.w; }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 x = *reinterpret_cast<const float4*>(Ar + m * ARS + q * QS + kw * 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        accw[j][q * 4 + 0] = fmaf(gj[j], x.x, accw[j][q * 4 + 0]);
                        accw[j][q * 4 + 1] = fmaf(gj[j], x.y, accw[j][q * 4 + 1]);
                        accw[j][q * 4 + 2] = fmaf(gj[j], x.z, accw[j][q * 4 + 2]);
                        accw[j][q * 4 + 3] = fmaf(gj[j], x.w, accw[j][q * 4 + 3]);
                    }
                }
            }
        }
        // input gradient partial: P[b][k] = sum_n G[n][m] * W3[n][k]
        if (dg_active)
        for (int kc = kd; kc < KD; kc += KDT) {
            float accp[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int q = 0; q < 8; ++q) accp[i][q] = 0.f;
#pragma unroll 2
            for (int n = 0; n < S; ++n) {
                const float4 gv = *reinterpret_cast<const float4*>(Gn + n * kGnStride + md * 4);
                const float4 w0 = *reinterpret_cast<const float4*>(Wr + n * DFP + kc * 4);
                const float4 w1 = *reinterpret_cast<const float4*>(Wr + n * DFP + DFP / 2 + kc * 4);
                const float gi[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    accp[i][0] = fmaf(gi[i], w0.x, accp[i][0]); accp[i][1] = fmaf(gi[i], w0.y, accp[i][1]);
                    accp[i][2] = fmaf(gi[i], w0.z, accp[i][2]); accp[i][3] = fmaf(gi[i], w0.w, accp[i][3]);
                    accp[i][4] = fmaf(gi[i], w1.x, accp[i][4]); accp[i][5] = fmaf(gi[i], w1.y, accp[i][5]);
                    accp[i][6] = fmaf(gi[i], w1.z, accp[i][6]); accp[i][7] = fmaf(gi[i], w1.w, accp[i][7]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t b = b0 + md * 4 + i;
                if (b < a.B) {
                    float* prow = a.P + ((size_t)g * a.B + b) * DFP;
                    *reinterpret_cast<float4*>(prow + kc * 4) = make_float4(accp[i][0], accp[i][1], accp[i][2], accp[i][3]);
                    *reinterpret_cast<float4*>(prow + DFP / 2 + kc * 4) =
                        make_float4(accp[i][4], accp[i][5], accp[i][6], accp[i][7]);
                }
            }
        }
        __syncthreads();
    }
    // accumulate this stage's weight gradient into the per-(bt) accumulator (exclusively owned by this CTA)
    if (wg_active) {
        const int QS = DFP / 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float* wrow = a.dW3acc + ((size_t)bt * a.Np + (size_t)g * S + nt * 4 + j) * DFP;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // single writer per element: reductions without return value (no stall on the read, deterministic)
                float* p = wrow + q * QS + kw * 4;
                atomicAdd(p + 0, accw[j][q * 4 + 0]); atomicAdd(p + 1, accw[j][q * 4 + 1]);
                atomicAdd(p + 2, accw[j][q * 4 + 2]); atomicAdd(p + 3, accw[j][q * 4 + 3]);
            }
        }
        if (kw == 0) {
            float* p = a.db3acc + (size_t)bt * a.Np + (size_t)g * S + nt * 4;
            atomicAdd(p + 0, accb[0]); atomicAdd(p + 1, accb[1]); atomicAdd(p + 2, accb[2]); atomicAdd(p + 3, accb[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// hidden_bwd: batch-split.  Sums the per-group partials, walks the hidden layers backwards, stores each layer's
// pre-activation gradient feature-major for hidden_wgrad, and applies the RK adjoint update
//     gy0 += dzs ;  gk_j += kcoef[j] * dzs  (j = earlier stages feeding this stage's input).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) hidden_bwd_kernel(const __grid_constant__ HiddenBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int R = a.R;
    const int RQ = R / 4;
    const int64_t b0 = (int64_t)blockIdx.x * R;
    float* buf0 = sm;
    float* buf1 = sm + (size_t)a.Dmax * R;
    float* acts = buf1 + (size_t)a.Dmax * R;            // [F][Dmax][R] saved layer outputs of these rows
    float* wsm = acts + (size_t)a.F * a.Dmax * R;
    // everything this kernel will need from global memory is requested up front (asynchronously):
    //   the saved outputs of every hidden layer (for act'), the weights, and (below) gy / gk for the final update
    for (int l = 0; l < a.F; ++l) {
        const float* __restrict__ src = a.actT[l + 1];
        const int Dout = a.D[l + 1];
        for (int i = tid; i < Dout * (R / 4); i += kThreads) {
            const int o = i / (R / 4), q = i % (R / 4);
            if (b0 + q * 4 < a.Bp) cp_async_16(acts + ((size_t)l * a.Dmax + o) * R + q * 4, src + (size_t)o * a.Bp + b0 + q * 4);
        }
    }
    if (a.w_in_smem) {
        const float* src = a.W[0];
        for (int i = tid * 4; i < a.wsm_floats; i += kThreads * 4) cp_async_16(wsm + i, src + i);
    }
    pdl_trigger();
    pdl_wait();  // P, gy, gk come from the previous kernels
    if (a.ctrl && a.ctrl->done) { cp_async_wait_all_(); return; }
    // gy / gk elements this thread updates at the end (fast path: at most 4 per thread)
    const bool few = a.dz_out == nullptr && a.H * R <= 4 * kThreads;
    float pre_gy[4], pre_gk[4][3];
    if (few) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * kThreads;
            const int h = idx / R, r = idx % R;
            const bool ok = idx < a.H * R && b0 + r < a.B;
            const size_t off = (size_t)h * a.Bp + b0 + r;
            pre_gy[e] = ok ? a.gyT[off] : 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) pre_gk[e][j] = (ok && j < a.n_k && a.kcoef[j] != 0.f) ? a.gkT[j][off] : 0.f;
        }
    }

    // 1. dL/d(final-layer input)[k][r] = sum_g P[g][b][k]: one float4 of k per thread, groups streamed 16 deep
    {
        const int DF = a.D[a.F];
        const int K4 = a.DFP / 4;
        for (int idx = tid; idx < R * K4; idx += kThreads) {
            const int r = idx / K4, k4 = idx % K4;
            const int64_t b = b0 + r;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b < a.B) {
                const float4* p = reinterpret_cast<const float4*>(a.P + (size_t)b * a.DFP) + k4;
                const size_t gs = (size_t)a.B * a.DFP / 4;
#pragma unroll 16
                for (int g = 0; g < a.n_hg; ++g) {
                    const float4 v = __ldg(p + (size_t)g * gs);
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
            }
            const int k = k4 * 4;
            if (k + 0 < DF) buf0[(k + 0) * R + r] = s.x;
            if (k + 1 < DF) buf0[(k + 1) * R + r] = s.y;
            if (k + 2 < DF) buf0[(k + 2) * R + r] = s.z;
            if (k + 3 < DF) buf0[(k + 3) * R + r] = s.w;
        }
    }
    cp_async_wait_all_();
    __syncthreads();
    // 2. hidden layers, last to first.  cur = gradient w.r.t. the OUTPUT of layer l (post-activation)
    float* cur = buf0;
    float* nxt = buf1;
    for (int l = a.F - 1; l >= 0; --l) {
        const int Din = a.D[l], Dout = a.D[l + 1];
        // dpre = cur * act'(out), in place in `cur`, and feature-major to global for the weight gradient
        const float* __restrict__ outs = acts + (size_t)l * a.Dmax * R;
        float* __restrict__ dpreT = a.dpreT[l];
        const int act = a.act[l];
        for (int idx = tid; idx < Dout * R; idx += kThreads) {
            const int o = idx / R, r = idx % R;
            const int64_t b = b0 + r;
            float v = 0.f;
            if (b < a.B) {
                if (act == NCDE_ACT_GATE_IN && o >= Din) {
                    // out = s * x with s = sigmoid(pre):  d/dpre = g x s (1 - s);  the direct d/dx = g s is stashed in the free
                    // rows [Din, 2 Din) of the other buffer and added to the input gradient after the matrix product below
                    const float x = l == 0 ? a.actT[0][(size_t)(o - Din) * a.Bp + b] : acts[((size_t)(l - 1) * a.Dmax + (o - Din)) * R + r];
                    const float sg = x != 0.f ? outs[idx] / x : 0.f;
                    const float g = cur[idx];
                    nxt[idx] = g * sg;
                    v = (g * x) * ((1.f - sg) * sg);
                } else {
                    v = cur[idx] * act_grad_from_output(outs[idx], act);
                }
                dpreT[(size_t)o * a.Bp + b] = v;
            } else if (act == NCDE_ACT_GATE_IN && o >= Din) {
                nxt[idx] = 0.f;
            }
            cur[idx] = v;
        }
        __syncthreads();
        // d(in)[i][r] = sum_o dpre[o][r] * W[o][i]
        const int ldi = a.ldi[l];
        for (int item = tid; item < Din * RQ; item += kThreads) {
            const int i = item % Din, q = item / Din;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.w_in_smem) acc = matvec4<true>(wsm + a.wsm_off[l] + i, ldi, cur + q * 4, R, Dout, acc);
            else acc = matvec4<false>(a.W[l] + i, ldi, cur + q * 4, R, Dout, acc);
            *reinterpret_cast<float4*>(nxt + i * R + q * 4) = acc;
        }
        __syncthreads();
        if (act == NCDE_ACT_GATE_IN) {
            for (int idx = tid; idx < Din * R; idx += kThreads) nxt[idx] += nxt[Din * R + idx];
            __syncthreads();
        }
        float* t = cur; cur = nxt; nxt = t;
    }
    // 3. RK adjoint update with dzs = cur[h][r]  (or, for the continuous adjoint, just hand dzs out)
    if (a.dz_out) {
        for (int idx = tid; idx < a.H * R; idx += kThreads) {
            const int h = idx / R, r = idx % R;
            if (b0 + r < a.B) a.dz_out[(size_t)h * a.Bp + b0 + r] = cur[idx];
        }
        return;
    }
    if (few) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * kThreads;
            const int h = idx / R, r = idx % R;
            if (idx < a.H * R && b0 + r < a.B) {
                const size_t off = (size_t)h * a.Bp + b0 + r;
                const float d = cur[idx];
                a.gyT[off] = pre_gy[e] + d;
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (j < a.n_k && a.kcoef[j] != 0.f) a.gkT[j][off] = fmaf(a.kcoef[j], d, pre_gk[e][j]);
            }
        }
        return;
    }
    for (int idx = tid; idx < a.H * R; idx += kThreads) {
        const int h = idx / R, r = idx % R;
        const int64_t b = b0 + r;
        if (b < a.B) {
            const size_t off = (size_t)h * a.Bp + b;
            const float d = cur[idx];
            a.gyT[off] += d;
            for (int j = 0; j < a.n_k; ++j)
                if (a.kcoef[j] != 0.f) a.gkT[j][off] = fmaf(a.kcoef[j], d, a.gkT[j][off]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// hidden_wgrad: weight-stationary and batch-split.  CTA (tile, split) owns a 64x64 tile of one hidden weight matrix
// and the batch rows of one split:
//     gWp[split][o][i] += sum_{b in split} dpre[o][b] * in[i][b] ;  gbp[split][o] += sum_b dpre[o][b]
// Thread = 4 out x 4 in.  Layers that share a slot (the reference's repeated nn.Linear) are folded into the
// same tile, so the accumulation into the shared gradient is race-free and ordered.  The per-split accumulators live in
// the workspace across all stages (each is owned by exactly one CTA) and are summed once by hidden_wgrad_reduce.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) hidden_wgrad_kernel(const __grid_constant__ WgradArgs a) {
    __shared__ __align__(16) float dS[kWgRows][kWgTile + 4];  // [b][o]
    __shared__ __align__(16) float aS[kWgRows][kWgTile + 4];  // [b][i]
    pdl_trigger();
    pdl_wait();  // dpre comes from hidden_bwd
    if (a.ctrl && a.ctrl->done) return;
    const int tid = threadIdx.x;
    int slot = 0;
    while (slot + 1 < a.n_slots && (int)blockIdx.x >= a.tile_begin[slot + 1]) ++slot;
    const int tile = blockIdx.x - a.tile_begin[slot];
    const int split = blockIdx.y;
    const int Dout = a.Dout[slot], Din = a.Din[slot];
    const int tiles_i = (Din + kWgTile - 1) / kWgTile;
    const int o0 = (tile / tiles_i) * kWgTile, i0 = (tile % tiles_i) * kWgTile;
    const int ty = tid / 16, tx = tid % 16;  // 16 x 16 threads, each 4 x 4
    float acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
    float accb[4] = {0.f, 0.f, 0.f, 0.f};
    const int row_begin = split * a.rows_per_split;
    const int row_end = min(a.B, row_begin + a.rows_per_split);
    for (int sl = 0; sl < a.n_stage * a.n_lay[slot]; ++sl) {
        const int sg = sl / a.n_lay[slot];
        const int l = a.lay[slot][sl % a.n_lay[slot]];
        const float* __restrict__ dT = a.dpreT[sg][l];
        const float* __restrict__ xT = a.actT[sg][l];
        for (int bc = row_begin; bc < row_end; bc += kWgRows) {
            // global reads are coalesced along the batch; the tiles are stored transposed ([b][feature])
            for (int idx = tid; idx < kWgTile * kWgRows; idx += kThreads) {
                const int f = idx / kWgRows, bb = idx % kWgRows;
                const int b = bc + bb;
                const bool in_rows = b < row_end;
                dS[bb][f] = (in_rows && o0 + f < Dout) ? dT[(size_t)(o0 + f) * a.Bp + b] : 0.f;
                aS[bb][f] = (in_rows && i0 + f < Din) ? xT[(size_t)(i0 + f) * a.Bp + b] : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int bb = 0; bb < kWgRows; ++bb) {
                const float4 d = *reinterpret_cast<const float4*>(&dS[bb][ty * 4]);
                const float4 x = *reinterpret_cast<const float4*>(&aS[bb][tx * 4]);
                const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc[u][0] = fmaf(dv[u], x.x, acc[u][0]); acc[u][1] = fmaf(dv[u], x.y, acc[u][1]);
                    acc[u][2] = fmaf(dv[u], x.z, acc[u][2]); acc[u][3] = fmaf(dv[u], x.w, acc[u][3]);
                    accb[u] += dv[u];
                }
            }
            __syncthreads();
        }
    }
    float* gWp = a.gWp[slot] + (size_t)split * Dout * Din;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int o = o0 + ty * 4 + u;
        if (o >= Dout) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int i = i0 + tx * 4 + v;
            if (i < Din) gWp[(size_t)o * Din + i] += acc[u][v];
        }
        if (i0 == 0 && tx == 0) a.gbp[slot][(size_t)split * Dout + o] += accb[u];
    }
}

// gW[slot] += sum_split gWp[slot][split] (once per backward pass)
__global__ void hidden_wgrad_reduce_kernel(const __grid_constant__ WgradArgs a) {
    const int slot = blockIdx.y;
    const int n = a.Dout[slot] * a.Din[slot];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) {
        float s = 0.f;
        for (int sp = 0; sp < a.n_split; ++sp) s += a.gWp[slot][(size_t)sp * n + idx];
        a.gW[slot][idx] += s;
    }
    if (a.gb[slot] && idx < a.Dout[slot]) {
        float s = 0.f;
        for (int sp = 0; sp < a.n_split; ++sp) s += a.gbp[slot][(size_t)sp * a.Dout[slot] + idx];
        a.gb[slot][idx] += s;
    }
}

// dst[i] += src[i]
__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

// packed row index of (h, c): h-groups of Hg rows-of-Cp, each group padded to Npad rows
__host__ __device__ inline int64_t packed_n(int h, int c, int Cp, int Hg, int Npad) {
    return (int64_t)(h / Hg) * Npad + (int64_t)(h % Hg) * Cp + c;
}

// cmul / coff: packed column of channel c is c * cmul + coff (gated final layer: cmul 2, coff 0 sigmoid head / 1 tanh head)
__global__ void unpack_final_grad_kernel(const float* __restrict__ dW3acc, const float* __restrict__ db3acc,
                                         float* __restrict__ gW, float* __restrict__ gb, int H, int C, int Cp, int Hg,
                                         int Npad, int DF, int DFP, int Np, int n_bt, int cmul = 1, int coff = 0) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)H * C * DF;
    if (idx < total) {
        const int k = (int)(idx % DF);
        const int hc = (int)(idx / DF);
        const int h = hc / C, c = hc % C;
        const int64_t n = packed_n(h, c * cmul + coff, Cp, Hg, Npad);
        float s = 0.f;
        for (int bt = 0; bt < n_bt; ++bt) s += dW3acc[((size_t)bt * Np + n) * DFP + k];
        gW[idx] += s;
    }
    if (gb && idx < (int64_t)H * C) {
        const int h = (int)idx / C, c = (int)idx % C;
        const int64_t n = packed_n(h, c * cmul + coff, Cp, Hg, Npad);
        float s = 0.f;
        for (int bt = 0; bt < n_bt; ++bt) s += db3acc[(size_t)bt * Np + n];
        gb[idx] += s;
    }
}

}  // namespace ncde
