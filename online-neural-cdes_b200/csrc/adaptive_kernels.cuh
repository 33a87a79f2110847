// Device-side step control of the adaptive Dormand-Prince 5(4) solve.
//
// Everything the reference decides on the host per attempted step — error norm, accept/reject with min_step /
// max_step overrides, the next step size, the dense-output evaluation at the requested times, the initial step
// selection — runs in these kernels on a control block in device memory (torchdiffeq/_impl/rk_common.py:163-313,
// misc.py:32-89, interp.py:1-48).  The host only enqueues attempts; kernels of attempts after the last output are
// no-ops (ctrl->done).  State is fp32, every time-like scalar fp64, exactly as in the reference.
#pragma once
#include <math.h>

#include "solve_kernels.cuh"

namespace ncde {

// Dormand-Prince tableau (torchdiffeq/_impl/dopri5.py:5-30), rounded to fp32 where the reference does
// (`tableau.*.to(dtype=y0.dtype)`, rk_common.py:155-159)
__constant__ double kDpAlpha[6] = {1 / 5., 3 / 10., 4 / 5., 8 / 9., 1., 1.};
__constant__ double kDpBeta[6][6] = {
    {1 / 5., 0, 0, 0, 0, 0},
    {3 / 40., 9 / 40., 0, 0, 0, 0},
    {44 / 45., -56 / 15., 32 / 9., 0, 0, 0},
    {19372 / 6561., -25360 / 2187., 64448 / 6561., -212 / 729., 0, 0},
    {9017 / 3168., -355 / 33., 46732 / 5247., 49 / 176., -5103 / 18656., 0},
    {35 / 384., 0, 500 / 1113., 125 / 192., -2187 / 6784., 11 / 84.},
};
__constant__ double kDpErr[7] = {35 / 384. - 1951 / 21600., 0, 500 / 1113. - 22642 / 50085., 125 / 192. - 451 / 720.,
                                 -2187 / 6784. - -12231 / 42400., 11 / 84. - 649 / 6300., -1. / 60.};
__constant__ double kDpMid[7] = {6025192743 / 30085553152. / 2, 0, 51252292925 / 65400821598. / 2,
                                 -2691868925 / 45128329728. / 2, 187940372067 / 1594534317056. / 2,
                                 -1776094331 / 19743644256. / 2, 11237099 / 235043384. / 2};

struct AdaptParams {
    double t0, rtol, atol, min_step, max_step, first_step, safety, ifactor, dfactor;
    long long max_attempts;
    int n_out;
    int time_sign;
    int keep_counters;   // adjoint: one controller per output interval, statistics accumulate over intervals
};

// stage times and combine coefficients of the attempt that starts at (t0, dt) — rk_common.py:58-75
__device__ inline void prep_stage_tabs(AdaptCtrl& c) {
    const float t0f = (float)c.t0, dtf = (float)c.dt, t1f = (float)(c.t0 + c.dt);
    c.step_dt = c.dt;
    for (int i = 0; i < 6; ++i) {
        StageTab& tb = c.tab[i + 1];
        const float alpha = (float)kDpAlpha[i];
        // stages with alpha == 1 are evaluated one ulp before t1 (Perturb.PREV, misc.py:182-187)
        const float ts = (kDpAlpha[i] == 1.) ? nextafterf(t1f, -INFINITY) : __fadd_rn(t0f, __fmul_rn(alpha, dtf));
        tb.t = c.time_sign < 0 ? -ts : ts;
        for (int j = 0; j < NCDE_MAX_STAGES; ++j) tb.coef[j] = j <= i ? __fmul_rn((float)kDpBeta[i][j], dtf) : 0.f;
    }
}

__global__ void adapt_init_kernel(AdaptCtrl* ctrl, AdaptParams p) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    AdaptCtrl& c = *ctrl;
    c.t0 = p.t0; c.dt = p.first_step; c.t_lo = p.t0; c.t_hi = p.t0; c.step_dt = 0; c.acc_dt = 0;
    c.rtol = p.rtol; c.atol = p.atol; c.min_step = p.min_step; c.max_step = p.max_step;
    c.safety = p.safety; c.ifactor = p.ifactor; c.dfactor = p.dfactor;
    c.accept = 0; c.done = p.n_out <= 1; c.j_begin = 1; c.j_end = 1; c.j_out = 1; c.n_out = p.n_out;
    c.time_sign = p.time_sign;
    if (p.keep_counters) { c.nfe += 1; c.max_attempts = c.attempted + p.max_attempts; }
    else { c.attempted = 0; c.accepted = 0; c.nfe = 1; c.max_attempts = p.max_attempts; c.flags = 0; }
    for (int i = 0; i <= NCDE_MAX_STAGES; ++i) {
        c.tab[i].t = p.time_sign < 0 ? -(float)p.t0 : (float)p.t0;
        for (int j = 0; j < NCDE_MAX_STAGES; ++j) c.tab[i].coef[j] = 0.f;
    }
    c.h0 = c.d0 = c.d1 = c.d2 = 0.f; c.dt_init = p.first_step;
    if (p.first_step > 0) prep_stage_tabs(c);
}

// block-wide sum of two doubles; result valid in thread 0
__device__ inline void block_sum2(double& a, double& b) {
    __shared__ double sa[32], sb[32];
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        a = l < nw ? sa[l] : 0.0; b = l < nw ? sb[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    }
}

// mode 0: partial sums of (y0/scale)^2 and (f0/scale)^2;  mode 1: ((f1-f0)/scale)^2      scale = atol + |y0| rtol
__global__ void adapt_norm_kernel(const AdaptCtrl* ctrl, int mode, const float* __restrict__ yT, const float* __restrict__ f0T,
                                  const float* __restrict__ f1T, int B, int Bp, int H, double* __restrict__ partials) {
    pdl_trigger();
    pdl_wait();
    const float atol = (float)ctrl->atol, rtol = (float)ctrl->rtol;
    double s0 = 0, s1 = 0;
    const int64_t n = (int64_t)H * Bp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if ((int)(i % Bp) >= B) continue;
        const float y = yT[i];
        const float scale = __fadd_rn(atol, __fmul_rn(fabsf(y), rtol));
        if (mode == 0) {
            const float a = y / scale, b = f0T[i] / scale;
            s0 += (double)(a * a); s1 += (double)(b * b);
        } else {
            const float a = __fsub_rn(f1T[i], f0T[i]) / scale;
            s0 += (double)(a * a);
        }
    }
    block_sum2(s0, s1);
    if (threadIdx.x == 0) { partials[blockIdx.x] = s0; partials[gridDim.x + blockIdx.x] = s1; }
}

// first half of _select_initial_step (misc.py:47-60): h0 and the probe evaluation point
__global__ void adapt_init_step1_kernel(AdaptCtrl* ctrl, const double* __restrict__ partials, int nblocks, double n_elems) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x != 0) return;
    AdaptCtrl& c = *ctrl;
    double s0 = 0, s1 = 0;
    for (int i = 0; i < nblocks; ++i) { s0 += partials[i]; s1 += partials[nblocks + i]; }
    const float d0 = sqrtf((float)(s0 / n_elems)), d1 = sqrtf((float)(s1 / n_elems));
    float h0;
    if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f; else h0 = __fdiv_rn(__fmul_rn(0.01f, d0), d1);
    c.h0 = h0; c.d0 = d0; c.d1 = d1;
    StageTab& tb = c.tab[NCDE_MAX_STAGES];
    tb.t = __fadd_rn((float)c.t0, h0);
    if (c.time_sign < 0) tb.t = -tb.t;
    for (int j = 0; j < NCDE_MAX_STAGES; ++j) tb.coef[j] = j == 0 ? h0 : 0.f;
    c.nfe += 1;
}

// second half (misc.py:62-71): d2, h1, dt = min(100 h0, h1); order + 1 = 5 (rk_common.py:167 passes order - 1 = 4)
__global__ void adapt_init_step2_kernel(AdaptCtrl* ctrl, const double* __restrict__ partials, int nblocks, double n_elems) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x != 0) return;
    AdaptCtrl& c = *ctrl;
    double s0 = 0;
    for (int i = 0; i < nblocks; ++i) s0 += partials[i];
    const float d2 = __fdiv_rn(sqrtf((float)(s0 / n_elems)), c.h0);
    float h1;
    if (c.d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, __fmul_rn(c.h0, 1e-3f));
    else h1 = powf(__fdiv_rn(0.01f, fmaxf(c.d1, d2)), 0.2f);
    c.dt = (double)fminf(__fmul_rn(100.f, c.h0), h1);
    c.d2 = d2; c.dt_init = c.dt;
    prep_stage_tabs(c);
}

struct DopriArgs {
    AdaptCtrl* ctrl;
    int B, Bp, H;
    float* yT;             // current state (updated in place on accept)
    const float* y1T;      // candidate = input of the 7th stage
    float* kT[7];          // k0 (FSAL) .. k6
    const double* out_t;   // device (T)
    float* z_out;          // (T,B,H)
    double* partials;
    int nblocks;
};

// err = k . (dt c_err);  ratio^2 partial sums of (err / (atol + rtol max(|y0|,|y1|)))^2   (rk_common.py:84, misc.py:74-76)
__global__ void dopri_err_kernel(const __grid_constant__ DopriArgs a) {
    pdl_trigger();
    pdl_wait();
    const AdaptCtrl& c = *a.ctrl;
    if (c.done) return;
    const float atol = (float)c.atol, rtol = (float)c.rtol, dtf = (float)c.step_dt;
    float ce[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) ce[j] = __fmul_rn(dtf, (float)kDpErr[j]);
    double s = 0, bad = 0;
    const int64_t n = (int64_t)a.H * a.Bp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if ((int)(i % a.Bp) >= a.B) continue;
        float err = 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) err = fmaf(a.kT[j][i], ce[j], err);
        const float y0 = a.yT[i], y1 = a.y1T[i];
        const float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(y0), fabsf(y1))));
        const float r = err / tol;
        s += (double)(r * r);
        if (!isfinite(y0)) bad += 1.0;
    }
    block_sum2(s, bad);
    if (threadIdx.x == 0) { a.partials[blockIdx.x] = s; a.partials[gridDim.x + blockIdx.x] = bad; }
}

// accept / reject, next step size, output bookkeeping (rk_common.py:269-305, misc.py:79-89)
__device__ inline void controller_update(AdaptCtrl& c, float ratio_f, const double* out_t) {
    const double dt = c.step_dt;
    if (!(c.t0 + dt > c.t0)) { c.flags |= NCDE_FLAG_DT_UNDERFLOW; c.done = 1; c.accept = 0; return; }
    bool accept = ratio_f <= 1.f;
    if (dt > c.max_step) accept = false;
    if (dt <= c.min_step) accept = true;
    if (c.attempted < 64) { c.trace[c.attempted][0] = dt; c.trace[c.attempted][1] = (double)ratio_f; c.trace[c.attempted][2] = accept ? 1.0 : 0.0; }
    c.attempted += 1;
    c.nfe += 6;
    if (accept) {
        c.accepted += 1;
        c.acc_dt = dt;
        c.t_lo = c.t0;
        c.t_hi = c.t0 + dt;
        c.t0 = c.t_hi;
        c.j_begin = c.j_out;
        while (c.j_out < c.n_out && out_t[c.j_out] <= c.t_hi) ++c.j_out;
        c.j_end = c.j_out;
        if (c.j_out >= c.n_out) c.done = 1;
    }
    c.accept = accept ? 1 : 0;
    // _optimal_step_size, order = 5
    const double ratio = (double)ratio_f;
    double dt_next;
    if (ratio_f == 0.f) dt_next = dt * c.ifactor;
    else {
        const double dfac = ratio_f < 1.f ? 1.0 : c.dfactor;
        const double fac = fmin(c.ifactor, fmax(c.safety / pow(ratio, 0.2), dfac));
        dt_next = dt * fac;
    }
    dt_next = fmin(fmax(dt_next, c.min_step), c.max_step);
    c.dt = dt_next;
    if (!c.done) {
        if (c.attempted >= c.max_attempts) { c.flags |= NCDE_FLAG_MAX_STEPS; c.done = 1; }
        else prep_stage_tabs(c);
    }
}

__global__ void dopri_ctrl_kernel(const __grid_constant__ DopriArgs a) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x != 0) return;
    AdaptCtrl& c = *a.ctrl;
    if (c.done) { c.accept = 0; return; }
    double s = 0, bad = 0;
    for (int i = 0; i < a.nblocks; ++i) { s += a.partials[i]; bad += a.partials[a.nblocks + i]; }
    if (bad > 0) c.flags |= NCDE_FLAG_NONFINITE;
    controller_update(c, sqrtf((float)(s / ((double)a.B * a.H))), a.out_t);
}

// on accept: dense output at every requested time inside the step (interp.py:1-48, rk_common.py:307-313), then
// y <- y1 and k0 <- k6 (first-same-as-last)
__global__ void dopri_accept_kernel(const __grid_constant__ DopriArgs a) {
    __shared__ float tile[32][33];
    pdl_trigger();
    pdl_wait();
    const AdaptCtrl& c = *a.ctrl;
    if (!c.accept) return;
    const float dtf = (float)c.acc_dt;  // NOT step_dt: the controller has already prepared the next attempt
    float cm[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) cm[j] = __fmul_rn(dtf, (float)kDpMid[j]);
    const int b0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    float ce[4][5];  // 4 elements per thread (32x32 tile, 32x8 threads): coefficients e, d, c, b, a
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = threadIdx.y + e * 8;
        const int h = h0 + i, b = b0 + threadIdx.x;
        float e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
        if (h < a.H && b < a.B) {
            const size_t off = (size_t)h * a.Bp + b;
            const float y0 = a.yT[off], y1 = a.y1T[off];
            float k[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) k[j] = a.kT[j][off];
            float ym = 0.f;
#pragma unroll
            for (int j = 0; j < 7; ++j) ym = fmaf(k[j], cm[j], ym);
            ym = y0 + ym;
            const float f0 = k[0], f1 = k[6];
            e4 = 2.f * dtf * (f1 - f0) - 8.f * (y1 + y0) + 16.f * ym;
            e3 = dtf * (5.f * f0 - 3.f * f1) + 18.f * y0 + 14.f * y1 - 32.f * ym;
            e2 = dtf * (f1 - 4.f * f0) - 11.f * y0 - 5.f * y1 + 16.f * ym;
            e1 = dtf * f0;
            e0 = y0;
            a.yT[off] = y1;
            a.kT[0][off] = f1;
        }
        ce[e][0] = e0; ce[e][1] = e1; ce[e][2] = e2; ce[e][3] = e3; ce[e][4] = e4;
    }
    for (int j = c.j_begin; j < c.j_end; ++j) {
        const float x = (float)((a.out_t[j] - c.t_lo) / (c.t_hi - c.t_lo));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float total = ce[e][0] + x * ce[e][1];
            float xp = x;
#pragma unroll
            for (int q = 2; q < 5; ++q) { xp = xp * x; total = total + xp * ce[e][q]; }
            tile[threadIdx.y + e * 8][threadIdx.x] = total;
        }
        __syncthreads();
        float* out = a.z_out + (size_t)j * a.B * a.H;
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int b = b0 + i, h = h0 + threadIdx.x;
            if (b < a.B && h < a.H) out[(size_t)b * a.H + h] = tile[threadIdx.x][i];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Adaptive continuous adjoint: the same controller over the augmented state (y, a, g_theta...) with the reference's
// mixed norm  max(|vjp_t|, rms(y), rms(a), max_p rms(g_theta_p))  (adjoint.py:235-246).
// Each component ("segment") is handled by its own launch; partial sums live in partials[seg][2][nblocks].
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxSeg = 2 + 2 * (NCDE_MAX_LAYERS + 1);

struct AugCtrlArgs {
    AdaptCtrl* ctrl;
    const double* partials;
    int nblocks, n_seg;
    double count[kMaxSeg];
    const double* out_t;   // device [2]: interval start and target in reversed time
};

__device__ inline int seg_norms(const AugCtrlArgs& a, int which, float* out_max) {
    float mx = 0.f;
    int arg = 0;
    for (int sg = 0; sg < a.n_seg; ++sg) {
        double s = 0;
        const double* p = a.partials + ((size_t)sg * 2 + which) * a.nblocks;
        for (int i = 0; i < a.nblocks; ++i) s += p[i];
        const float v = sqrtf((float)(s / a.count[sg]));
#ifdef NCDE_DEBUG_SEG
        if (a.ctrl->attempted < 3) printf("attempt %lld which %d seg %d sum %g count %g norm %g\n", a.ctrl->attempted, which, sg, s, a.count[sg], v);
#endif
        if (v > mx) { mx = v; arg = sg; }
    }
    *out_max = mx;
    return arg;
}

__global__ void aug_init_step1_kernel(const __grid_constant__ AugCtrlArgs a) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x != 0) return;
    AdaptCtrl& c = *a.ctrl;
    float d0, d1;
    seg_norms(a, 0, &d0);
    seg_norms(a, 1, &d1);
    float h0;
    if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f; else h0 = __fdiv_rn(__fmul_rn(0.01f, d0), d1);
    c.h0 = h0; c.d0 = d0; c.d1 = d1;
    StageTab& tb = c.tab[NCDE_MAX_STAGES];
    tb.t = __fadd_rn((float)c.t0, h0);
    if (c.time_sign < 0) tb.t = -tb.t;
    for (int j = 0; j < NCDE_MAX_STAGES; ++j) tb.coef[j] = j == 0 ? h0 : 0.f;
    c.nfe += 1;
}

__global__ void aug_init_step2_kernel(const __grid_constant__ AugCtrlArgs a) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x != 0) return;
    AdaptCtrl& c = *a.ctrl;
    float d2;
    seg_norms(a, 0, &d2);
    d2 = __fdiv_rn(d2, c.h0);
    float h1;
    if (c.d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, __fmul_rn(c.h0, 1e-3f));
    else h1 = powf(__fdiv_rn(0.01f, fmaxf(c.d1, d2)), 0.2f);
    c.dt = (double)fminf(__fmul_rn(100.f, c.h0), h1);
    c.d2 = d2; c.dt_init = c.dt;
    prep_stage_tabs(c);
}

// error-ratio partial sums of one segment: err = k . (dt c_err), tol = atol + rtol max(|s0|, |s1|)
__global__ void aug_err_kernel(const AdaptCtrl* ctrl, const float* __restrict__ s0, const float* __restrict__ s1,
                               const float* k0, const float* k1, const float* k2, const float* k3, const float* k4,
                               const float* k5, const float* k6, int64_t n, int Bp, int B, double* __restrict__ partials) {
    pdl_trigger();
    pdl_wait();
    const AdaptCtrl& c = *ctrl;
    if (c.done) return;
    const float atol = (float)c.atol, rtol = (float)c.rtol, dtf = (float)c.step_dt;
    const float* k[7] = {k0, k1, k2, k3, k4, k5, k6};
    float ce[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) ce[j] = __fmul_rn(dtf, (float)kDpErr[j]);
    double s = 0, bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (Bp > 0 && (int)(i % Bp) >= B) continue;
        float err = 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) err = fmaf(k[j][i], ce[j], err);
        const float a0 = s0[i], a1 = s1[i];
        const float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(a0), fabsf(a1))));
        const float r = err / tol;
        s += (double)(r * r);
        if (!isfinite(a0)) bad += 1.0;
    }
    block_sum2(s, bad);
    if (threadIdx.x == 0) { partials[blockIdx.x] = s; partials[gridDim.x + blockIdx.x] = bad; }
}

__global__ void aug_ctrl_kernel(const __grid_constant__ AugCtrlArgs a) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x != 0) return;
    AdaptCtrl& c = *a.ctrl;
    if (c.done) { c.accept = 0; return; }
    float ratio;
    const int arg = seg_norms(a, 0, &ratio);
    const long long slot = c.attempted;
    double bad = 0;
    for (int i = 0; i < a.nblocks; ++i) bad += a.partials[(size_t)1 * a.nblocks + i];   // non-finite y
    if (bad > 0) c.flags |= NCDE_FLAG_NONFINITE;
    controller_update(c, ratio, a.out_t);
    if (slot < 64) c.trace[slot][2] += 10.0 * arg;   // diagnostics: which segment set the error ratio
}

// accepted step of one segment: s <- s1, k0 <- k6; when the step reaches the interval end, the dense output at the
// target time is written to `out` first (interp.py:1-48)
__global__ void aug_accept_kernel(const AdaptCtrl* ctrl, float* __restrict__ s, const float* __restrict__ s1, float* k0,
                                  const float* k1, const float* k2, const float* k3, const float* k4, const float* k5,
                                  const float* k6, float* __restrict__ out, int64_t n, const double* __restrict__ out_t) {
    pdl_trigger();
    pdl_wait();
    const AdaptCtrl& c = *ctrl;
    if (!c.accept) return;
    const float dtf = (float)c.acc_dt;
    const bool emit = out != nullptr && c.done && c.j_end > c.j_begin;
    const float x = emit ? (float)((out_t[c.j_end - 1] - c.t_lo) / (c.t_hi - c.t_lo)) : 0.f;
    const float* k[7] = {k0, k1, k2, k3, k4, k5, k6};
    float cm[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) cm[j] = __fmul_rn(dtf, (float)kDpMid[j]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float y0 = s[i], y1 = s1[i], f0 = k0[i], f1 = k6[i];
        if (emit) {
            float ym = 0.f;
#pragma unroll
            for (int j = 0; j < 7; ++j) ym = fmaf(k[j][i], cm[j], ym);
            ym = y0 + ym;
            const float e4 = 2.f * dtf * (f1 - f0) - 8.f * (y1 + y0) + 16.f * ym;
            const float e3 = dtf * (5.f * f0 - 3.f * f1) + 18.f * y0 + 14.f * y1 - 32.f * ym;
            const float e2 = dtf * (f1 - 4.f * f0) - 11.f * y0 - 5.f * y1 + 16.f * ym;
            const float e1 = dtf * f0;
            float total = y0 + x * e1;
            float xp = x * x;
            total = total + xp * e2;
            xp = xp * x;
            total = total + xp * e3;
            xp = xp * x;
            total = total + xp * e4;
            out[i] = total;
        }
        s[i] = y1;
        k0[i] = f1;
    }
}

// time-gradient component of the augmented derivative: out = sum_{b<B,h} a[h][b] q[h][b], one CTA, fixed summation order
__global__ void aug_dot_kernel(const AdaptCtrl* ctrl, const float* __restrict__ aT, const float* __restrict__ qT, int B, int Bp,
                               int H, float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    if (ctrl->done) return;
    double s = 0, unused = 0;
    const int64_t n = (int64_t)H * Bp;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        if ((int)(i % Bp) >= B) continue;
        s += (double)__fmul_rn(aT[i], qT[i]);
    }
    block_sum2(s, unused);
    if (threadIdx.x == 0) *out = (float)s;
}

// dst = -src (the y component of the augmented derivative is -f)
__global__ void negate_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = -src[i];
}

}  // namespace ncde
