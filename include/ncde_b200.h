/*
 * ncde_b200.h — C ABI of the B200-native Neural-CDE solve path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every pointer marked "device" is a
 * CUDA device pointer on the current device; `stream` is a cudaStream_t passed as void*.  All entry points are
 * asynchronous on `stream` (no host synchronisation inside) and return NCDE_OK or a negative error code;
 * ncde_last_error() gives a thread-local message.  The caller owns every buffer.
 *
 * The reference (jambo6/online-neural-cdes) has no native layer: the functions replaced here are the Python
 * hot path of its vendored torchcde 0.2.0 / torchdiffeq 0.2.1.  Each entry point cites the reference code it
 * replaces (paths relative to the reference root).
 */
#ifndef NCDE_B200_H
#define NCDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCDE_ABI_VERSION 4
#define NCDE_MAX_LAYERS 8
#define NCDE_MAX_STAGES 7

enum ncde_status {
    NCDE_OK = 0,
    NCDE_ERR_INVALID = -1,      /* bad argument (maps to ValueError) */
    NCDE_ERR_CUDA = -2,         /* a CUDA runtime call failed */
    NCDE_ERR_UNSUPPORTED = -3,  /* shape outside what the kernels were built for */
    NCDE_ERR_WORKSPACE = -4     /* caller-provided workspace too small */
};

enum ncde_dtype { NCDE_F32 = 0, NCDE_F64 = 1 };
enum ncde_path_kind { NCDE_PATH_LINEAR = 0, NCDE_PATH_CUBIC = 1 };
enum ncde_method { NCDE_EULER = 0, NCDE_RK4_38 = 1, NCDE_DOPRI5 = 2 };
/* NCDE_ACT_GATE_IN (hidden layers only, out_dim == 2 * in_dim): out[i] = pre[i] for i < in_dim, out[in_dim + i] =
 * sigmoid(pre[in_dim + i]) * in[i] — with W = [I ; W_f] this layer turns x into [x ; sigmoid(W_f x + b_f) * x], the reset-gated
 * second input of GRUGatedVectorField (src/ncde/vector_fields/gating.py:35-61).  Fixed-grid fp32 path. */
enum ncde_act { NCDE_ACT_NONE = 0, NCDE_ACT_RELU = 1, NCDE_ACT_TANH = 2, NCDE_ACT_GATE_IN = 3 };
/* how the control enters the vector field (`vector_field_type` of torchcde.cdeint, modules/torchcde/torchcde/solver.py:112-137):
 * MATMUL      dz/dt = f(z) . dX/dt                 f: H -> H*C, contracted with the path derivative
 * EVALUATE    dz/dt = f([z, X(t)])                 f: H+C -> H, no contraction
 * DERIVATIVE  dz/dt = f([z, dX/dt(t)])             f: H+C -> H, no contraction
 * EVALUATE / DERIVATIVE run on the fixed-grid fp32 path (ncde_solve_fwd / ncde_solve_bwd). */
enum ncde_vf_type { NCDE_VF_MATMUL = 0, NCDE_VF_EVALUATE = 1, NCDE_VF_DERIVATIVE = 2 };
/* arithmetic of the vector-field GEMMs: fp32 FFMA; bf16 tcgen05 tensor-core tiles with fp32 accumulate; or BF16X3, the
 * parity-grade tensor-core mode: every operand is a (hi, lo) pair of bf16 tiles and every GEMM is three tcgen05 MMAs
 * hi*hi + lo*hi + hi*lo into one fp32 accumulator (16 mantissa bits per operand, bf16 exponent range).  BF16X3 runs on the
 * persistent fixed-grid kernels only (euler / rk4, vector_field_type matmul, outputs on grid points). */
enum ncde_precision { NCDE_PREC_FP32 = 0, NCDE_PREC_BF16 = 1, NCDE_PREC_BF16X3 = 2 };

/* device status words written by kernels (read them after synchronising the stream) */
enum ncde_flag_bits {
    NCDE_FLAG_NAN_TIME = 1,       /* rectilinear: NaN in the time channel (AssertionError in the reference) */
    NCDE_FLAG_NONFINITE = 2,      /* dopri5: non-finite state (rk_common.py:233) */
    NCDE_FLAG_DT_UNDERFLOW = 4,   /* dopri5: t0 + dt == t0 (rk_common.py:232) */
    NCDE_FLAG_MAX_STEPS = 8       /* dopri5: attempt budget exhausted before reaching t[-1] */
};

const char* ncde_version(void);
const char* ncde_last_error(void);
int ncde_abi_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * Interpolation constructors.  x is (n_series, L, C) contiguous, dtype f32 or f64.
 * ---------------------------------------------------------------------------------------------------------- */

/* Forward fill along the length axis.  Replaces torchcde.misc.forward_fill
 * (modules/torchcde/torchcde/misc.py:103-126).  out may alias x. */
int ncde_forward_fill(int dtype, const void* x, void* out, int64_t n_series, int64_t L, int64_t C, void* stream);

/* Rectilinear preparation: ffill, duplicate rows, advance the time column, drop the last row ->
 * out (n_series, 2L-1, C).  Replaces _prepare_rectilinear_interpolation
 * (modules/torchcde/torchcde/interpolation_linear.py:87-128).  Sets NCDE_FLAG_NAN_TIME in *flags (device int32,
 * caller zero-initialised) if the time channel holds a NaN. */
int ncde_rectilinear_prepare(int dtype, const void* x, void* out, int64_t n_series, int64_t L, int64_t C,
                             int time_index, int32_t* flags, void* stream);

/* In-place linear fill of missing (NaN) values per (series, channel): ends imputed with the first / last
 * observation, interior by linear interpolation in t, all-NaN channel -> zeros.  Replaces
 * _linear_interpolation_coeffs_with_missing_values (modules/torchcde/torchcde/interpolation_linear.py:13-84).
 * t: device (L). */
int ncde_linear_fill_missing(int dtype, void* x, const void* t, int64_t n_series, int64_t L, int64_t C,
                             void* stream);

/* Natural cubic spline coefficients, out (n_series, L-1, 4C) = [a | b | 2c | 3d].  NaN-aware (version 0/1 end
 * handling).  Replaces _natural_cubic_spline_coeffs (modules/torchcde/torchcde/interpolation_cubic.py:7-190) and
 * tridiagonal_solve (modules/torchcde/torchcde/misc.py:13-67).  scratch: device, ncde_cubic_scratch_bytes(). */
size_t ncde_cubic_scratch_bytes(int dtype, int64_t n_series, int64_t L, int64_t C);
int ncde_natural_cubic_coeffs(int dtype, const void* x, const void* t, void* out, int64_t n_series, int64_t L,
                              int64_t C, int version, void* scratch, void* stream);

/* derivs[s,i,c] = (coeffs[s,i+1,c]-coeffs[s,i,c]) / (t[i+1]-t[i]); the LinearInterpolation constructor
 * (modules/torchcde/torchcde/interpolation_linear.py:198). */
int ncde_linear_derivs(int dtype, const void* coeffs, const void* t, void* derivs, int64_t n_series, int64_t K,
                       int64_t C, void* stream);

/* Evaluate a path or its derivative at n_t query times (device array tq).  kind LINEAR: coeffs (n_series,K,C),
 * derivs (n_series,K-1,C) (may be NULL when deriv == 0); kind CUBIC: coeffs (n_series,K-1,4C).
 * out: (n_series, n_t, C).  index_out (device int64, n_t, nullable) receives the knot index
 * bucketize(t, knots) - 1 clamped to [0, K-2].  Replaces LinearInterpolation / NaturalCubicSpline
 * ._interpret_t/.evaluate/.derivative (interpolation_linear.py:212-234, interpolation_cubic.py:315-336). */
int ncde_path_eval(int kind, int dtype, const void* coeffs, const void* derivs, const void* knots,
                   int64_t n_series, int64_t K, int64_t C, const void* tq, int64_t n_t, int deriv, void* out,
                   int64_t* index_out, void* stream);

/* Ragged batches (offline preprocessing and loader, get_data/transformers.py:50-85, get_data/common.py:59-80,
 * experiments/ingredients/loader.py:100-113,190-196).  x: (n_series, Lmax, C) with series s occupying its first lengths[s]
 * rows (lengths: device int32, every entry >= 2; the remaining rows are ignored).  One call does for ALL series what the reference
 * does in a Python loop per series:
 *   initial_nan_to_zero: missing values of the first row become 0 (transformers.py:52-55);
 *   NCDE_RAGGED_LINEAR       out (n, Lmax, C):       linear_interpolation_coeffs(d)                 knots t = 0..L_s-1
 *   NCDE_RAGGED_RECTILINEAR  out (n, 2Lmax-1, Cout): linear_interpolation_coeffs(d, rectilinear=time_index)
 *                            intensity != 0 appends C-1 channels (Cout = 2C-1, time_index must be 0): the running count of
 *                            observations of channels 1..C-1, duplicated like the values (loader.py:100-113; a first-row value
 *                            that is missing or exactly 0 does not count, as there)
 *   NCDE_RAGGED_CUBIC        out (n, Lmax-1, 4C):    natural_cubic_coeffs(d)
 *   pad != 0: rows past a series' own K_s = L_s | 2L_s-1 | L_s-1 repeat its last row — what PadRaggedTensors + ForwardFill give
 *   when the loader batches the coefficient lists (loader.py:190-196); pad == 0 leaves them NaN.
 * Results for the valid rows are bit-identical to the per-series reference calls.  flags: NCDE_FLAG_NAN_TIME as in
 * ncde_rectilinear_prepare.  scratch: device, ncde_ragged_scratch_bytes(). */
enum ncde_ragged_method { NCDE_RAGGED_LINEAR = 0, NCDE_RAGGED_RECTILINEAR = 1, NCDE_RAGGED_CUBIC = 2 };
size_t ncde_ragged_scratch_bytes(int method, int dtype, int64_t n_series, int64_t Lmax, int64_t C);
int ncde_ragged_interpolate(int method, int dtype, const void* x, const int32_t* lengths, void* out, int64_t n_series,
                            int64_t Lmax, int64_t C, int time_index, int initial_nan_to_zero, int intensity, int pad,
                            void* scratch, int32_t* flags, void* stream);

/* Backward of ncde_path_eval with respect to the coefficients (what autograd does through
 * LinearInterpolation / NaturalCubicSpline .evaluate/.derivative in the reference, e.g. for h0 = Linear(X.evaluate(0)) of a
 * stacked Neural CDE, src/ncde/ncde.py:179-181).  grad_out (n_series, n_t, C); grad_coeffs has the shape of coeffs and is
 * ACCUMULATED into (caller zero-initialises).  One thread per (series, channel) walks the query times in order: no atomics. */
int ncde_path_eval_bwd(int kind, int dtype, const void* knots, int64_t n_series, int64_t K, int64_t C, const void* tq,
                       int64_t n_t, int deriv, const void* grad_out, void* grad_coeffs, void* stream);

/* Log-ODE transform, depth 1 or 2: log-signatures of the piecewise-linear path x (n_series, Lp, d) over W windows
 * (window w = rows idx[w]..idx[w+1], idx a device int32[W+1]), cumulatively summed, with the first row set to x[:,0,:] padded
 * with zeros -> out (n_series, W+1, d + d(d-1)/2).  Channel order as Signatory's "words" mode.  wscale (device, W values of
 * the same dtype, nullable) scales each window first.  Replaces the signatory.Logsignature / stack / cumsum part of
 * torchcde.log_ode._logsignature_windows (modules/torchcde/torchcde/log_ode.py:49-70); the windowing and the linear fill
 * (:15-47) stay on the host side / go through ncde_linear_fill_missing. */
int ncde_logsig_windows(int dtype, const void* x, const int32_t* idx, const void* wscale, void* out, int64_t n_series,
                        int64_t Lp, int d, int depth, int W, void* stream);

/* Linear / rectilinear hybrid, compaction step (src/ncde/interpolation.py:224-253): `full` (n_series, K, C) is the
 * rectilinear path of the series whose linear channels were filled beforehand; linear channels (chan_kind 0) are shifted up
 * by one row, a row is kept only if a time / rectilinear channel (chan_kind 1) differs from the previous row, and each series is
 * padded to K rows by repeating its last kept row.  counts (device int32[n_series]) receives the kept rows per series; the
 * caller slices `out` to max(counts) rows (the reference's pad_sequence + forward_fill). */
int ncde_hybrid_compact(int dtype, const void* full, const int32_t* chan_kind, void* out, int32_t* counts, int64_t n_series,
                        int64_t K, int64_t C, void* stream);

/* SmoothLinearInterpolation (src/ncde/interpolation.py:6-183; unit knot spacing, as the reference requires): coefficients of
 * the cubic (terms = 4: A,B,C,D) or quintic (terms = 6: A..F) that replaces the linear piece on [k, k + eps) after every
 * interior knot so that the first (and second) derivatives are continuous -> out (n_series, K-2, C, terms).
 * _setup_cubic_matching_coefficients / _setup_quintic_matching_coefficients (:146-183). */
int ncde_smooth_matching_coeffs(int dtype, const void* coeffs, void* out, int64_t n_series, int64_t K, int64_t C, double eps,
                                int terms, void* stream);
/* evaluate (deriv = 0) / derivative (deriv = 1) of the smoothed path at n_t times -> out (n_series, n_t, C)
 * (SmoothLinearInterpolation._interpret_t / evaluate / derivative, :72-123). */
int ncde_path_eval_smooth(int dtype, const void* coeffs, const void* derivs, const void* knots, const void* match, int terms,
                          double eps, int64_t n_series, int64_t K, int64_t C, const void* tq, int64_t n_t, int deriv, void* out,
                          void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * The solve: z_t = z_0 + int f_theta(z_s) dX_s, replacing torchcde.cdeint -> torchdiffeq.odeint[_adjoint]
 * (modules/torchcde/torchcde/solver.py:102-238; modules/torchdiffeq/torchdiffeq/_impl/solvers.py:48-119,
 * fixed_grid.py:6-29, rk_common.py:41-313, adjoint.py:9-215).  fp32 state, fp64 step control.
 * ---------------------------------------------------------------------------------------------------------- */

/* f_theta: a chain of Linear layers; layer l maps in_dim[l] -> out_dim[l] with activation act[l].  The last
 * layer has out_dim == H*C and its output row h*C + c is the (h, c) entry of the vector-field matrix
 * (src/ncde/vector_fields/base.py:83-104).  Layers with equal slot[] share one weight/bias tensor (the
 * reference's list-multiplication of one nn.Linear, base.py:64-69): give them the same W/bias pointers; their
 * gradients accumulate into the same gW/gbias. */
typedef struct ncde_mlp {
    int32_t n_layers;
    int32_t in_dim[NCDE_MAX_LAYERS];
    int32_t out_dim[NCDE_MAX_LAYERS];
    int32_t act[NCDE_MAX_LAYERS];
    int32_t slot[NCDE_MAX_LAYERS];
    const float* W[NCDE_MAX_LAYERS];    /* device, (out_dim, in_dim) row-major — torch.nn.Linear layout */
    const float* bias[NCDE_MAX_LAYERS]; /* device, (out_dim) or NULL */
    /* Optional multiplicative gate on the last layer (MinimalGatedVectorField, src/ncde/vector_fields/gating.py:7-32):
     * output = sigmoid(W_gate a + bias_gate) * tanh(W[last] a + bias[last]), W_gate shaped like W[last].  NULL: no gate.
     * Needs n_layers < NCDE_MAX_LAYERS: ncde_solve_bwd accumulates the gate's gradients into gW[n_layers] / gbias[n_layers].
     * Fixed-grid fp32 path (ncde_solve_fwd / ncde_solve_bwd), at most 64 channels. */
    const float* W_gate;
    const float* bias_gate;
} ncde_mlp_t;

typedef struct ncde_path {
    int32_t kind;        /* ncde_path_kind */
    int64_t K;           /* number of knots */
    const float* knots;  /* device (K) */
    const float* coeffs; /* device; LINEAR (B,K,C), CUBIC (B,K-1,4C) */
    const float* derivs; /* device; LINEAR (B,K-1,C) precomputed by ncde_linear_derivs, or NULL; CUBIC NULL */
    /* LINEAR only, optional (SmoothLinearInterpolation, src/ncde/interpolation.py:6-123): polynomial pieces that replace the
     * linear ones on [t_k, t_k + match_eps) for every interior knot k = 1..K-2 */
    const float* match;  /* device (B, K-2, C, match_terms) from ncde_smooth_matching_coeffs, or NULL */
    int32_t match_terms; /* 4: cubic (first derivatives matched), 6: quintic (second derivatives too) */
    float match_eps;     /* gradient_matching_eps */
} ncde_path_t;

/* Fixed-grid schedule, built by the host exactly as torchdiffeq does (solvers.py:77-119): everything here lives
 * in HOST memory and is consumed while enqueuing. */
typedef struct ncde_fixed_grid {
    int64_t n_steps;
    const float* stage_t; /* (n_steps, n_stages) stage times after the cast to the state dtype (misc.py:181) */
    const float* dt;      /* (n_steps) t1 - t0 rounded to fp32 */
    int64_t n_out;        /* number of output times T (t[0] included) */
    const int64_t* out_step; /* (T) grid step whose interval emits output j; out_step[0] is ignored */
    const int32_t* out_mode; /* (T) 0: state at step start, 1: state at step end, 2: linear interpolation */
    const float* out_slope;  /* (T) (t[j]-t0)/(t1-t0) for mode 2 */
} ncde_fixed_grid_t;

/* dopri5 controller parameters (rk_common.py:122-160; misc.py:32-89). */
typedef struct ncde_adaptive {
    double rtol, atol, min_step, max_step, first_step /* < 0: select automatically */, safety, ifactor, dfactor;
    int64_t max_attempts; /* attempt budget enqueued without host sync */
    int64_t n_out;        /* T */
    const double* out_t;  /* HOST (T) increasing output times */
} ncde_adaptive_t;

typedef struct ncde_problem {
    int64_t B;
    int32_t H, C;
    int32_t method;    /* ncde_method */
    int32_t precision; /* ncde_precision */
    ncde_mlp_t mlp;
    ncde_path_t path;
    ncde_fixed_grid_t grid;   /* EULER / RK4_38 */
    ncde_adaptive_t adaptive; /* DOPRI5 */
    int32_t vf_type;          /* ncde_vf_type; for EVALUATE / DERIVATIVE mlp.in_dim[0] == H + C and the last layer has H outputs */
} ncde_problem_t;

/* Byte sizes of the two caller-provided device buffers: `saved` carries what the backward pass needs (stage
 * inputs, hidden activations, path derivatives per stage) and must stay alive until ncde_solve_bwd;
 * `workspace` is scratch valid for one call. */
size_t ncde_solve_saved_bytes(const ncde_problem_t* p, int need_grad);
/* backward: 0 = forward pass, 1 = ncde_solve_bwd, 2 = ncde_solve_bwd with grad_coeffs != NULL (returns 0 if that is
 * not supported for this problem). */
size_t ncde_solve_workspace_bytes(const ncde_problem_t* p, int backward);

/* Forward.  z0 (B,H) device; z_out (T,B,H) device (torchdiffeq's layout; cdeint returns its (B,T,H) permuted
 * view, solver.py:227-229).  saved may be NULL when need_grad == 0.  flags: device int32 (nullable).
 * stats (device int64[4], nullable): attempted steps, accepted steps, vector-field evaluations, kernels launched.
 * launches (host, nullable): number of kernels this call enqueued. */
int ncde_solve_fwd(const ncde_problem_t* p, const float* z0, float* z_out, void* saved, int need_grad,
                   void* workspace, size_t workspace_bytes, int32_t* flags, int64_t* stats, int64_t* launches,
                   void* stream);

/* dopri5 forward (NCDE_DOPRI5).  Error norm, accept/reject, next step size, initial step selection and dense output
 * all run on a control block in device memory (replaces RKAdaptiveStepsizeODESolver, rk_common.py:117-313,
 * misc.py:32-89, interp.py).  The host enqueues attempts in chunks of 32 and, between chunks, reads a completion
 * flag the GPU wrote one chunk earlier, so the device never waits for the host; at most adaptive.max_attempts
 * attempts are enqueued (NCDE_FLAG_MAX_STEPS if that was not enough).  stats (device int64[200], nullable):
 * attempted steps, accepted steps, vector-field evaluations, ncde_flag_bits, then the selected initial step (fp64
 * bits) and h0, d0, d1, d2 of its selection (fp32 bits), then (dt, error ratio, accepted) as fp64 for the first 64 attempts.  No saved state: gradients of the
 * adaptive solve are not implemented in this revision. */
int ncde_solve_adaptive_fwd(const ncde_problem_t* p, const float* z0, float* z_out, void* workspace,
                            size_t workspace_bytes, int64_t* stats, int64_t* launches, void* stream);

/* Backward of the fixed-grid solve (discretise-then-optimise: the exact gradient autograd produces through the
 * reference's step loop).  grad_out (T,B,H).  Writes grad_z0 (B,H); ACCUMULATES into gW[l]/gbias[l] (torch
 * layout, one pointer per layer; layers sharing a slot must pass the same pointer).  grad_coeffs (nullable):
 * gradient w.r.t. path.coeffs — what autograd sends into X's coefficient buffer through X.derivative(t) at every stage
 * (modules/torchcde/torchcde/solver.py:128-132; interpolation_linear.py:198,231-234; interpolation_cubic.py:331-336) —
 * same shape as path.coeffs, ACCUMULATED (the caller zero-initialises).  This is what stacked Neural CDEs need
 * (src/ncde/stacked.py:120-128; modules/torchcde/test/test_tricks.py:54-106).  fp32 precision, un-smoothed paths. */
int ncde_solve_bwd(const ncde_problem_t* p, const float* grad_out, const void* saved, float* grad_z0,
                   float* const* gW, float* const* gbias, float* grad_coeffs, void* workspace,
                   size_t workspace_bytes, int64_t* launches, void* stream);

/* Continuous adjoint with a fixed-grid adjoint method (NCDE_EULER / NCDE_RK4_38); replaces OdeintAdjointMethod.backward
 * (modules/torchdiffeq/torchdiffeq/_impl/adjoint.py:36-145).  p->grid describes the BACKWARD schedule, interval by
 * interval from the last output interval to the first: n_steps = total steps, dt = step sizes (> 0, reversed time),
 * stage_t = the stage times expressed in forward time (already negated back and cast to fp32).  interval_steps
 * (host, n_out - 1 entries, same order) gives the number of steps of each interval.  y_out, grad_out: (T,B,H)
 * forward outputs and their gradients.  Writes grad_z0; ACCUMULATES into gW / gbias like ncde_solve_bwd. */
size_t ncde_solve_adjoint_workspace_bytes(const ncde_problem_t* p);
int ncde_solve_adjoint_bwd(const ncde_problem_t* p, const int64_t* interval_steps, int64_t n_out, const float* y_out,
                           const float* grad_out, float* grad_z0, float* const* gW, float* const* gbias,
                           void* workspace, size_t workspace_bytes, int64_t* launches, void* stream);

/* Continuous adjoint with dopri5 as the adjoint method.  p->adaptive carries the ADJOINT tolerances / options and the
 * forward output times (host).  One device-controlled adaptive solve of the augmented state (y, a, g_theta) per output
 * interval in reversed time, with the reference's mixed error norm max(|vjp_t|, rms(y), rms(a), max_p rms(g_theta_p))
 * (adjoint.py:235-246; vjp_t, the time-gradient scalar, is integrated for cubic paths and is exactly zero for linear ones).  The host reads a completion flag after
 * every chunk of 4 attempts (one stream synchronisation per chunk).  stats: device int64[200] = attempted, accepted,
 * evaluations, flags summed over the intervals, [8..199] = (dt, error ratio, accepted) as doubles for the first 64 attempts. */
size_t ncde_solve_adjoint_adaptive_workspace_bytes(const ncde_problem_t* p);
int ncde_solve_adjoint_adaptive_bwd(const ncde_problem_t* p, const float* y_out, const float* grad_out, float* grad_z0,
                                    float* const* gW, float* const* gbias, void* workspace, size_t workspace_bytes,
                                    int64_t* stats, int64_t* launches, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Per-kernel timing with CUDA events on the launching stream (measurement support for bench.py; no reference
 * counterpart).  While enabled, every launch of the kernel classes in `class_mask` is bracketed by an event
 * pair.  ncde_profile_read synchronises on the recorded events, sums elapsed milliseconds and launch counts per
 * class into ms[NCDE_PROF_CLASSES] / count[NCDE_PROF_CLASSES], and clears the record.
 * ---------------------------------------------------------------------------------------------------------- */
enum ncde_prof_class {
    NCDE_PROF_HIDDEN_FWD = 0, NCDE_PROF_FIELD_FWD = 1, NCDE_PROF_FIELD_BWD = 2, NCDE_PROF_HIDDEN_BWD = 3,
    NCDE_PROF_HIDDEN_WGRAD = 4, NCDE_PROF_OTHER = 5 /* dx_all: dX/dt of every stage, one launch per solve */,
    NCDE_PROF_SOLVE_FWD = 6, NCDE_PROF_SOLVE_BWD = 7 /* the persistent whole-pass kernels */, NCDE_PROF_CLASSES = 8
};
int ncde_profile_enable(int class_mask);
int ncde_profile_read(double* ms, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* NCDE_B200_H */
