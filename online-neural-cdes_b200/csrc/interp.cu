// Interpolation constructors and evaluators as sm_100a CUDA.
//
// These are HBM-bound copy / scan kernels: one thread owns one (series, channel) scalar series and walks the
// length axis; consecutive threads own consecutive channels, so every step of the walk is a coalesced row access.
// Compiled with -fmad=false: the reference evaluates each torch op with its own rounding (no fused
// multiply-add), and matching that is what makes the outputs bit-identical to the reference on the same inputs.
#include <stdlib.h>

#include "common.cuh"

namespace ncde {

template <typename T>
__device__ __forceinline__ bool is_nan(T v) { return v != v; }

// ---------------------------------------------------------------------------------------------------------------
// forward fill — torchcde/misc.py:103-126
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void forward_fill_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t n_series, int64_t L,
                                    int64_t C) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * C) return;
    int64_t s = tid / C, c = tid % C;
    const T* xs = x + s * L * C + c;
    T* os = out + s * L * C + c;
    T last = xs[0];  // stays NaN until the first observation, like the gather through cummax index 0
    for (int64_t i = 0; i < L; ++i) {
        T v = xs[i * C];
        if (!is_nan(v)) last = v;
        os[i * C] = last;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// rectilinear preparation — torchcde/interpolation_linear.py:87-128
//   value channels: out[2i] = out[2i+1] = ffill(x)[i]        time channel: out[2i] = t[i], out[2i+1] = t[i+1]
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void rectilinear_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t n_series, int64_t L,
                                   int64_t C, int time_index, int32_t* __restrict__ flags) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * C) return;
    int64_t s = tid / C, c = tid % C;
    const T* xs = x + s * L * C + c;
    T* os = out + s * (2 * L - 1) * C + c;
    const bool is_time = (c == time_index);
    T last = xs[0];
    bool bad_time = false;
    for (int64_t i = 0; i < L; ++i) {
        T v = xs[i * C];
        if (is_nan(v)) bad_time = bad_time || is_time; else last = v;
        if (is_time) {
            os[(2 * i) * C] = last;
            if (i > 0) os[(2 * i - 1) * C] = last;
        } else {
            os[(2 * i) * C] = last;
            if (i < L - 1) os[(2 * i + 1) * C] = last;
        }
    }
    if (bad_time && flags) atomicOr(flags, NCDE_FLAG_NAN_TIME);
}

// ---------------------------------------------------------------------------------------------------------------
// linear fill of missing values, in place — torchcde/interpolation_linear.py:13-84
// ---------------------------------------------------------------------------------------------------------------
// one scalar series xs[0], xs[C], ..., xs[(L-1)*C] (C = element stride between consecutive rows)
template <typename T>
__device__ void linear_fill_series(T* xs, int64_t C, int64_t L, const T* __restrict__ t) {
    int64_t first = -1, last = -1;
    for (int64_t i = 0; i < L; ++i) {
        if (!is_nan(xs[i * C])) { if (first < 0) first = i; last = i; }
    }
    if (first < 0) {  // nothing observed: constant zero path
        for (int64_t i = 0; i < L; ++i) xs[i * C] = T(0);
        return;
    }
    if (first == 0 && last == L - 1) {
        bool any = false;
        for (int64_t i = 0; i < L; ++i) any = any || is_nan(xs[i * C]);
        if (!any) return;
    }
    // impute the two ends with the first / last observation, then interpolate interior gaps between the
    // surrounding observed (or imputed) points
    if (first > 0) xs[0] = xs[first * C];
    if (last < L - 1) xs[(L - 1) * C] = xs[last * C];
    int64_t p = 0;
    int64_t i = 1;
    while (i < L) {
        if (!is_nan(xs[i * C])) { p = i; ++i; continue; }
        int64_t n = i + 1;
        while (is_nan(xs[n * C])) ++n;  // terminates: xs[L-1] is observed
        const T xp = xs[p * C], xn = xs[n * C], tp = t[p], tn = t[n];
        for (int64_t j = i; j < n; ++j) {
            T ratio = (t[j] - tp) / (tn - tp);
            xs[j * C] = xp + ratio * (xn - xp);
        }
        p = n;
        i = n + 1;
    }
}

template <typename T>
__global__ void linear_fill_kernel(T* __restrict__ x, const T* __restrict__ t, int64_t n_series, int64_t L,
                                   int64_t C) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * C) return;
    int64_t s = tid / C, c = tid % C;
    linear_fill_series<T>(x + s * L * C + c, C, L, t);
}

// ---------------------------------------------------------------------------------------------------------------
// linear derivs — LinearInterpolation.__init__, torchcde/interpolation_linear.py:198
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void linear_derivs_kernel(const T* __restrict__ coeffs, const T* __restrict__ t, T* __restrict__ derivs,
                                     int64_t n_series, int64_t K, int64_t C) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = n_series * (K - 1) * C;
    if (tid >= total) return;
    int64_t c = tid % C;
    int64_t i = (tid / C) % (K - 1);
    int64_t s = tid / (C * (K - 1));
    const T* cs = coeffs + s * K * C;
    derivs[tid] = (cs[(i + 1) * C + c] - cs[i * C + c]) / (t[i + 1] - t[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// evaluate / derivative — interpolation_linear.py:212-234, interpolation_cubic.py:315-336
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void path_eval_kernel(int kind, const T* __restrict__ coeffs, const T* __restrict__ derivs,
                                 const T* __restrict__ knots, int64_t n_series, int64_t K, int64_t C,
                                 const T* __restrict__ tq, int64_t n_t, int deriv, T* __restrict__ out,
                                 int64_t* __restrict__ index_out) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = n_series * n_t * C;
    if (tid >= total) return;
    int64_t c = tid % C;
    int64_t q = (tid / C) % n_t;
    int64_t s = tid / (C * n_t);
    const T t = tq[q];
    const int idx = knot_index<T>(knots, (int)K, t);
    if (index_out && s == 0 && c == 0) index_out[q] = idx;
    const T frac = t - knots[idx];
    T r;
    if (kind == NCDE_PATH_LINEAR) {
        if (deriv) {
            r = derivs[(s * (K - 1) + idx) * C + c];
        } else {
            const T lo = coeffs[(s * K + idx) * C + c];
            const T hi = coeffs[(s * K + idx + 1) * C + c];
            const T width = knots[idx + 1] - knots[idx];
            r = lo + frac * (hi - lo) / width;
        }
    } else {
        const T* row = coeffs + (s * (K - 1) + idx) * 4 * C;
        const T a = row[c], b = row[C + c], two_c = row[2 * C + c], three_d = row[3 * C + c];
        if (deriv) {
            T inner = two_c + three_d * frac;
            r = b + inner * frac;
        } else {
            T inner = T(0.5) * two_c + three_d * frac / T(3);
            inner = b + inner * frac;
            r = a + inner * frac;
        }
    }
    out[tid] = r;
}

// gradient of path_eval_kernel w.r.t. the coefficients; thread = (series, channel), sequential over the query times
template <typename T>
__global__ void path_eval_bwd_kernel(int kind, const T* __restrict__ knots, int64_t n_series, int64_t K, int64_t C,
                                     const T* __restrict__ tq, int64_t n_t, int deriv, const T* __restrict__ grad_out,
                                     T* __restrict__ grad_coeffs) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * C) return;
    const int64_t c = tid % C, s = tid / C;
    for (int64_t q = 0; q < n_t; ++q) {
        const T t = tq[q];
        const int idx = knot_index<T>(knots, (int)K, t);
        const T frac = t - knots[idx];
        const T g = grad_out[(s * n_t + q) * C + c];
        if (kind == NCDE_PATH_LINEAR) {
            const T width = knots[idx + 1] - knots[idx];
            T* lo = grad_coeffs + (s * K + idx) * C + c;
            if (deriv) {
                const T v = g / width;
                lo[0] = lo[0] - v;
                lo[C] = lo[C] + v;
            } else {
                // r = lo + frac * (hi - lo) / width
                const T v = g * frac / width;
                lo[0] = lo[0] + (g - v);
                lo[C] = lo[C] + v;
            }
        } else {
            T* row = grad_coeffs + (s * (K - 1) + idx) * 4 * C;
            if (deriv) {
                // r = b + (two_c + three_d * frac) * frac
                const T gi = g * frac;
                row[C + c] = row[C + c] + g;
                row[2 * C + c] = row[2 * C + c] + gi;
                row[3 * C + c] = row[3 * C + c] + gi * frac;
            } else {
                // r = a + (b + (0.5 * two_c + three_d * frac / 3) * frac) * frac
                const T g1 = g * frac, g2 = g1 * frac;
                row[c] = row[c] + g;
                row[C + c] = row[C + c] + g1;
                row[2 * C + c] = row[2 * C + c] + T(0.5) * g2;
                row[3 * C + c] = row[3 * C + c] + g2 * frac / T(3);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// natural cubic spline coefficients — interpolation_cubic.py:7-190, misc.py:13-67 (Thomas algorithm)
// One thread per scalar series; scratch arrays are laid out [L][n_threads] so every sweep step is coalesced.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ void cubic_series(const T* xs, int64_t XS, T* __restrict__ os, int64_t L, int64_t C, int version,
                             const T* __restrict__ t, T* __restrict__ s_x, T* __restrict__ s_d, T* __restrict__ s_r,
                             int32_t* __restrict__ s_i, int64_t N, int64_t tid) {
#define SX(i) s_x[(int64_t)(i) * N + tid]
#define SD(i) s_d[(int64_t)(i) * N + tid]
#define SR(i) s_r[(int64_t)(i) * N + tid]
#define SI(i) s_i[(int64_t)(i) * N + tid]
    // 1. observed points after end imputation (version 0: ends only; version 1: fill outward from first/last)
    int64_t first = -1, last = -1;
    for (int64_t i = 0; i < L; ++i) {
        if (!is_nan(xs[i * XS])) { if (first < 0) first = i; last = i; }
    }
    if (first < 0) {
        for (int64_t i = 0; i < L - 1; ++i) {
            os[i * 4 * C] = T(0); os[i * 4 * C + C] = T(0); os[i * 4 * C + 2 * C] = T(0); os[i * 4 * C + 3 * C] = T(0);
        }
        return;
    }
    const T x_first = xs[first * XS], x_last = xs[last * XS];
    int m = 0;
    for (int64_t i = 0; i < L; ++i) {
        T v = xs[i * XS];
        bool have = !is_nan(v);
        if (!have) {
            if (version == 0) {
                if (i == 0) { v = x_first; have = true; }
                else if (i == L - 1) { v = x_last; have = true; }
            } else {
                if (i < first) { v = x_first; have = true; }
                else if (i > last) { v = x_last; have = true; }
            }
        }
        if (have) { SX(m) = v; SI(m) = (int32_t)i; ++m; }
    }
    // 2. knot derivatives kd (stored back into s_r) on the observed grid
    if (m > 2) {
        // forward sweep
        T hinv_prev = T(0), dxs_prev = T(0);
        T d_prev = T(0), r_prev = T(0);
        for (int i = 0; i < m; ++i) {
            T hinv = T(0), dxs = T(0);
            if (i < m - 1) {
                hinv = T(1) / (t[SI(i + 1)] - t[SI(i)]);
                T three_dx = T(3) * (SX(i + 1) - SX(i));
                dxs = three_dx * (hinv * hinv);
            }
            T diag = (i == 0) ? hinv : (hinv + hinv_prev);
            diag = diag * T(2);
            T rhs = (i == 0) ? dxs : (dxs + dxs_prev);
            T d, r;
            if (i == 0) { d = diag; r = rhs; }
            else {
                T w = hinv_prev / d_prev;       // lower[i-1] / d[i-1]
                d = diag - w * hinv_prev;        // diag[i] - w * upper[i-1]
                r = rhs - w * r_prev;
            }
            SD(i) = d; SR(i) = r;
            d_prev = d; r_prev = r; hinv_prev = hinv; dxs_prev = dxs;
        }
        // back substitution (kd overwrites r)
        T nxt = SR(m - 1) / SD(m - 1);
        SR(m - 1) = nxt;
        for (int i = m - 2; i >= 0; --i) {
            T upper = T(1) / (t[SI(i + 1)] - t[SI(i)]);
            nxt = (SR(i) - upper * nxt) / SD(i);
            SR(i) = nxt;
        }
    }
    // 3. every original interval takes the polynomial of the observed piece it lies in, re-centred at its own
    //    left end (offset = observed-left-time - own time)
    int j = 0;  // current observed piece [SI(j), SI(j+1)]
    for (int64_t i = 0; i < L - 1; ++i) {
        while (j < m - 2 && (int64_t)SI(j + 1) <= i) ++j;
        const T tl = t[SI(j)], tr = t[SI(j + 1)];
        const T xl = SX(j), xr = SX(j + 1);
        T pa, pb, pc, pd;
        if (m == 2) {
            pa = xl; pb = (xr - xl) / (tr - tl); pc = T(0); pd = T(0);
        } else {
            const T hinv = T(1) / (tr - tl);
            const T hinv2 = hinv * hinv;
            const T six_dx = T(2) * (T(3) * (xr - xl));
            const T kl = SR(j), kr = SR(j + 1);
            pa = xl; pb = kl;
            pc = (six_dx * hinv - T(4) * kl - T(2) * kr) * hinv;
            pd = (-six_dx * hinv + T(3) * (kl + kr)) * hinv2;
        }
        const T off = tl - t[i];
        const T inner = (T(0.5) * pc - pd * off / T(3)) * off;
        os[i * 4 * C] = pa + (inner - pb) * off;
        os[i * 4 * C + C] = pb + (pd * off - pc) * off;
        os[i * 4 * C + 2 * C] = pc - T(2) * pd * off;
        os[i * 4 * C + 3 * C] = pd;
    }
#undef SX
#undef SD
#undef SR
#undef SI
}

template <typename T>
__global__ void cubic_coeffs_kernel(const T* __restrict__ x, const T* __restrict__ t, T* __restrict__ out,
                                    int64_t n_series, int64_t L, int64_t C, int version, T* __restrict__ s_x,
                                    T* __restrict__ s_d, T* __restrict__ s_r, int32_t* __restrict__ s_i) {
    const int64_t N = n_series * C;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= N) return;
    int64_t s = tid / C, c = tid % C;
    cubic_series<T>(x + s * L * C + c, C, out + s * (L - 1) * 4 * C + c, L, C, version, t, s_x, s_d, s_r, s_i, N, tid);
}

// ---------------------------------------------------------------------------------------------------------------
// Ragged batches: n series of different lengths packed into one NaN-padded (n, Lmax, C) tensor, series s
// having lengths[s] valid rows.  What the reference does with a Python loop over the series
// (get_data/transformers.py:50-85: d[:1][isnan] = 0, then linear_interpolation_coeffs / natural_cubic_coeffs per series) and,
// later, per batch (experiments/ingredients/loader.py:100-113 intensity channels, :190-196 PadRaggedTensors + ForwardFill)
// runs here as one launch per stage over all series; thread = (series, channel) as above.
// ---------------------------------------------------------------------------------------------------------------
// work = x with the first row's missing values set to zero ("causality", transformers.py:52-55); t = 0, 1, 2, ...
template <typename T>
__global__ void ragged_init_kernel(const T* __restrict__ x, T* __restrict__ work, T* __restrict__ t, int64_t n_series,
                                   int64_t Lmax, int64_t C, int64_t Kmax, int init_zero) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < Kmax) t[tid] = (T)tid;
    if (tid >= n_series * Lmax * C) return;
    T v = x[tid];
    if (init_zero && (tid / C) % Lmax == 0 && is_nan(v)) v = T(0);
    work[tid] = v;
}

template <typename T>
__global__ void arange_kernel(T* __restrict__ t, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i] = (T)i;
}

// rows [K, Kmax) of a finished series: repeat its last row (PadRaggedTensors + ForwardFill, loader.py:190-196) or NaN
template <typename T>
__device__ void ragged_tail(T* os, int64_t stride, int64_t K, int64_t Kmax, int pad) {
    const T fill = pad ? os[(K - 1) * stride] : (T)NAN;
    for (int64_t i = K; i < Kmax; ++i) os[i * stride] = fill;
}

template <typename T>
__global__ void ragged_linear_kernel(const T* __restrict__ work, const int32_t* __restrict__ lengths, const T* __restrict__ t,
                                     T* __restrict__ out, int64_t n_series, int64_t Lmax, int64_t C, int pad) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * C) return;
    const int64_t s = tid / C, c = tid % C, L = lengths[s];
    const T* xs = work + s * Lmax * C + c;
    T* os = out + s * Lmax * C + c;
    for (int64_t i = 0; i < L; ++i) os[i * C] = xs[i * C];
    linear_fill_series<T>(os, C, L, t);
    ragged_tail<T>(os, C, L, Lmax, pad);
}

// rectilinear (time channel 0 .. time_index) + optional observation-intensity channels: channel C + c - 1 of the output counts
// the observations of channel c >= 1 up to each row (loader.py:100-113: a first-row value that is missing OR exactly zero does
// not count; rows duplicated like the values, last one dropped)
template <typename T>
__global__ void ragged_rectilinear_kernel(const T* __restrict__ x, const T* __restrict__ work,
                                          const int32_t* __restrict__ lengths, const T* __restrict__ t, T* __restrict__ out,
                                          int64_t n_series, int64_t Lmax, int64_t C, int64_t Cout, int time_index,
                                          int pad, int32_t* __restrict__ flags) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * C) return;
    const int64_t s = tid / C, c = tid % C, L = lengths[s];
    const int64_t Kmax = 2 * Lmax - 1, K = 2 * L - 1;
    const T* xs = work + s * Lmax * C + c;
    T* os = out + s * Kmax * Cout + c;
    const bool is_time = (c == time_index);
    T last = xs[0];
    bool bad_time = false;
    for (int64_t i = 0; i < L; ++i) {
        T v = xs[i * C];
        if (is_nan(v)) bad_time = bad_time || is_time; else last = v;
        if (is_time) {
            os[(2 * i) * Cout] = last;
            if (i > 0) os[(2 * i - 1) * Cout] = last;
        } else {
            os[(2 * i) * Cout] = last;
            if (i < L - 1) os[(2 * i + 1) * Cout] = last;
        }
    }
    if (bad_time && flags) atomicOr(flags, NCDE_FLAG_NAN_TIME);
    linear_fill_series<T>(os, Cout, K, t);
    ragged_tail<T>(os, Cout, K, Kmax, pad);
    if (Cout > C && c >= 1) {
        const T* xr = x + s * Lmax * C + c;
        T* oi = out + s * Kmax * Cout + C + (c - 1);
        int64_t cnt = 0;
        for (int64_t i = 0; i < L; ++i) {
            const T v = xr[i * C];
            if (!is_nan(v) && !(i == 0 && v == T(0))) ++cnt;
            oi[(2 * i) * Cout] = (T)cnt;
            if (i < L - 1) oi[(2 * i + 1) * Cout] = (T)cnt;
        }
        ragged_tail<T>(oi, Cout, K, Kmax, pad);
    }
}

template <typename T>
__global__ void ragged_cubic_kernel(const T* __restrict__ work, const int32_t* __restrict__ lengths, const T* __restrict__ t,
                                    T* __restrict__ out, int64_t n_series, int64_t Lmax, int64_t C, int pad,
                                    T* __restrict__ s_x, T* __restrict__ s_d, T* __restrict__ s_r, int32_t* __restrict__ s_i) {
    const int64_t N = n_series * C;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= N) return;
    const int64_t s = tid / C, c = tid % C, L = lengths[s];
    T* os = out + s * (Lmax - 1) * 4 * C + c;
    cubic_series<T>(work + s * Lmax * C + c, C, os, L, C, 1, t, s_x, s_d, s_r, s_i, N, tid);
    for (int q = 0; q < 4; ++q) ragged_tail<T>(os + q * C, 4 * C, L - 1, Lmax - 1, pad);
}

// Branch-uniform restatement of linear_fill_series for a series held in a shared-memory column: a forward pass records, per
// row, the previous observation (16-bit index column `prev`), a backward pass carries the next one in registers and writes the
// interpolated value.  Every lane runs the same L iterations (the gap-chasing while-loops of linear_fill_series diverge within a
// warp because each (series, channel) has its own gap pattern).  Same operands, same expression, same rounding as above:
// results are bit-identical.  L < 65535.
template <typename T>
__device__ void linear_fill_uniform(T* col, int stride, unsigned short* prev, int64_t L, const T* __restrict__ t) {
    int first = -1, p = -1;
    bool any_nan = false;
    for (int i = 0; i < (int)L; ++i) {
        const bool obs = !is_nan(col[(int64_t)i * stride]);
        if (obs) { if (first < 0) first = i; p = i; } else any_nan = true;
        prev[(int64_t)i * stride] = (unsigned short)(p < 0 ? 0xFFFF : p);
    }
    const int last = p;
    if (first < 0) {  // nothing observed: constant zero path
        for (int i = 0; i < (int)L; ++i) col[(int64_t)i * stride] = T(0);
        return;
    }
    if (!any_nan) return;
    const T x_first = col[(int64_t)first * stride], x_last = col[(int64_t)last * stride];
    int n = (int)L - 1;   // next observed point; the right end counts as one (imputed with the last observation)
    T xn = x_last;
    for (int j = (int)L - 1; j >= 0; --j) {
        const T v = col[(int64_t)j * stride];
        const bool obs = !is_nan(v);
        T r = v;
        if (!obs) {
            if (j == (int)L - 1) r = x_last;
            else if (j == 0) r = x_first;
            else {
                int pi = prev[(int64_t)j * stride];
                T xp;
                if (pi == 0xFFFF) { pi = 0; xp = x_first; } else xp = col[(int64_t)pi * stride];
                const T tp = t[pi], tn = t[n];
                const T ratio = (t[j] - tp) / (tn - tp);
                r = xp + ratio * (xn - xp);
            }
            col[(int64_t)j * stride] = r;
        } else {
            n = j; xn = v;
        }
    }
}

// fixed-length linear fill through shared memory (ncde_linear_fill_missing fast path): sm = [L][blockDim] values + 16-bit indices
template <typename T>
__global__ void linear_fill_staged_kernel(T* __restrict__ x, const T* __restrict__ t, int64_t n_series, int64_t L, int64_t C) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T* col = reinterpret_cast<T*>(sm_raw) + threadIdx.x;
    unsigned short* prev = reinterpret_cast<unsigned short*>(sm_raw + (size_t)L * blockDim.x * sizeof(T)) + threadIdx.x;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * C) return;
    const int64_t s = tid / C, c = tid % C;
    T* xs = x + s * L * C + c;
    int64_t i = 0;
    for (; i + 8 <= L; i += 8) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = xs[(i + u) * C];
#pragma unroll
        for (int u = 0; u < 8; ++u) col[(i + u) * blockDim.x] = v[u];
    }
    for (; i < L; ++i) col[i * blockDim.x] = xs[i * C];
    linear_fill_uniform<T>(col, blockDim.x, prev, L, t);
    for (i = 0; i < L; ++i) xs[i * C] = col[i * blockDim.x];
}

// ---- shared-memory staged variants (the fast path) -----------------------------------------------------------------------------
// The walks above chase one dependent global load per row.  Here a thread first pulls its whole series into a shared-memory
// column with independent (unrolled, coalesced 128-byte-per-warp) loads — first-row zeroing applied on the way, so the work
// copy of ragged_init_kernel is not needed — runs the same device functions on shared memory, and streams the result out with
// independent stores.  HBM traffic: raw set read once, coefficients written once.  sm: [Lmax][blockDim.x] (+ the same again for
// the rectilinear scheme's second operand is not needed: values are written straight from the forward-filled column).
template <typename T>
__device__ __forceinline__ void stage_series(const T* __restrict__ xs, int64_t C, int64_t L, T* col, int stride, int init_zero) {
    int64_t i = 0;
    for (; i + 8 <= L; i += 8) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = xs[(i + u) * C];
#pragma unroll
        for (int u = 0; u < 8; ++u) col[(i + u) * stride] = v[u];
    }
    for (; i < L; ++i) col[i * stride] = xs[i * C];
    if (init_zero && is_nan(col[0])) col[0] = T(0);
}

template <typename T>
__global__ void ragged_linear_staged_kernel(const T* __restrict__ x, const int32_t* __restrict__ lengths, const T* __restrict__ t,
                                            T* __restrict__ out, int64_t n_series, int64_t Lmax, int64_t C, int init_zero, int pad) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T* col = reinterpret_cast<T*>(sm_raw) + threadIdx.x;
    const int stride = blockDim.x;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * C) return;
    const int64_t s = tid / C, c = tid % C, L = lengths[s];
    unsigned short* prev = reinterpret_cast<unsigned short*>(sm_raw + (size_t)Lmax * blockDim.x * sizeof(T)) + threadIdx.x;
    stage_series<T>(x + s * Lmax * C + c, C, L, col, stride, init_zero);
    linear_fill_uniform<T>(col, stride, prev, L, t);
    T* os = out + s * Lmax * C + c;
    const T fill = pad ? col[(L - 1) * stride] : (T)NAN;
    for (int64_t i = 0; i < Lmax; ++i) os[i * C] = i < L ? col[i * stride] : fill;
}

template <typename T>
__global__ void ragged_rectilinear_staged_kernel(const T* __restrict__ x, const int32_t* __restrict__ lengths,
                                                 const T* __restrict__ t, T* __restrict__ out, int64_t n_series, int64_t Lmax,
                                                 int64_t C, int64_t Cout, int time_index, int init_zero, int pad,
                                                 int32_t* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T* col = reinterpret_cast<T*>(sm_raw) + threadIdx.x;
    const int stride = blockDim.x;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * C) return;
    const int64_t s = tid / C, c = tid % C, L = lengths[s];
    const int64_t Kmax = 2 * Lmax - 1, K = 2 * L - 1;
    const T* xr = x + s * Lmax * C + c;
    stage_series<T>(xr, C, L, col, stride, 0);
    T* os = out + s * Kmax * Cout + c;
    // observation counts first (they need the raw values: loader.py:100-113), then forward fill in place
    if (Cout > C && c >= 1) {
        T* oi = out + s * Kmax * Cout + C + (c - 1);
        int64_t cnt = 0;
        T lastc = T(0);
        for (int64_t i = 0; i < L; ++i) {
            const T v = col[i * stride];
            if (!is_nan(v) && !(i == 0 && v == T(0))) ++cnt;
            lastc = (T)cnt;
            oi[(2 * i) * Cout] = lastc;
            if (i < L - 1) oi[(2 * i + 1) * Cout] = lastc;
        }
        const T fill = pad ? lastc : (T)NAN;
        for (int64_t i = K; i < Kmax; ++i) oi[i * Cout] = fill;
    }
    if (init_zero && is_nan(col[0])) col[0] = T(0);
    const bool is_time = (c == time_index);
    const bool lead_nan = is_nan(col[0]);
    T last = col[0];
    bool bad_time = false;
    for (int64_t i = 0; i < L; ++i) {
        const T v = col[i * stride];
        if (is_nan(v)) bad_time = bad_time || is_time; else last = v;
        col[i * stride] = last;
    }
    if (bad_time && flags) atomicOr(flags, NCDE_FLAG_NAN_TIME);
    // value channels: out[2i] = out[2i+1] = ffill[i]; time channel: out[2i] = t[i], out[2i+1] = t[i+1]
    for (int64_t i = 0; i < L; ++i) {
        const T v = col[i * stride];
        os[(2 * i) * Cout] = v;
        if (is_time) { if (i > 0) os[(2 * i - 1) * Cout] = v; }
        else if (i < L - 1) os[(2 * i + 1) * Cout] = v;
    }
    // a series that starts with missing values (only possible without the first-row rule) is back-filled by the generic path
    if (lead_nan) linear_fill_series<T>(os, Cout, K, t);
    const T fill = pad ? os[(K - 1) * Cout] : (T)NAN;
    for (int64_t i = K; i < Kmax; ++i) os[i * Cout] = fill;
}

template <typename T>
__global__ void ragged_cubic_staged_kernel(const T* __restrict__ x, const int32_t* __restrict__ lengths, const T* __restrict__ t,
                                           T* __restrict__ out, int64_t n_series, int64_t Lmax, int64_t C, int init_zero, int pad,
                                           T* __restrict__ s_x, T* __restrict__ s_d, T* __restrict__ s_r, int32_t* __restrict__ s_i) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T* col = reinterpret_cast<T*>(sm_raw) + threadIdx.x;
    const int stride = blockDim.x;
    const int64_t N = n_series * C;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= N) return;
    const int64_t s = tid / C, c = tid % C, L = lengths[s];
    stage_series<T>(x + s * Lmax * C + c, C, L, col, stride, init_zero);
    T* os = out + s * (Lmax - 1) * 4 * C + c;
    cubic_series<T>(col, stride, os, L, C, 1, t, s_x, s_d, s_r, s_i, N, tid);
    for (int q = 0; q < 4; ++q) ragged_tail<T>(os + q * C, 4 * C, L - 1, Lmax - 1, pad);
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)ceil_div(n, block); }

}  // namespace ncde

using namespace ncde;

#define DISPATCH_DTYPE(dtype, ...)                                                    \
    if ((dtype) == NCDE_F32) { typedef float T; __VA_ARGS__; }                        \
    else if ((dtype) == NCDE_F64) { typedef double T; __VA_ARGS__; }                  \
    else { set_error("unsupported dtype %d", (int)(dtype)); return NCDE_ERR_INVALID; }

extern "C" int ncde_forward_fill(int dtype, const void* x, void* out, int64_t n_series, int64_t L, int64_t C,
                                 void* stream) {
    ncde::DeviceGuard device_guard(x);
    NCDE_REQUIRE(x && out && n_series >= 0 && L >= 1 && C >= 1, NCDE_ERR_INVALID, "forward_fill: bad arguments");
    if (n_series == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, (forward_fill_kernel<T><<<grid_for(n_series * C, 128), 128, 0, st>>>(
                              (const T*)x, (T*)out, n_series, L, C)));
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

extern "C" int ncde_rectilinear_prepare(int dtype, const void* x, void* out, int64_t n_series, int64_t L,
                                        int64_t C, int time_index, int32_t* flags, void* stream) {
    ncde::DeviceGuard device_guard(x);
    NCDE_REQUIRE(x && out && n_series >= 0 && L >= 1 && C >= 1, NCDE_ERR_INVALID, "rectilinear: bad arguments");
    NCDE_REQUIRE(time_index >= 0 && time_index < C, NCDE_ERR_INVALID, "Time index must be in [0, %lld], was given %d.",
                 (long long)C - 1, time_index);
    if (n_series == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, (rectilinear_kernel<T><<<grid_for(n_series * C, 128), 128, 0, st>>>(
                              (const T*)x, (T*)out, n_series, L, C, time_index, flags)));
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

extern "C" int ncde_linear_fill_missing(int dtype, void* x, const void* t, int64_t n_series, int64_t L, int64_t C,
                                        void* stream) {
    ncde::DeviceGuard device_guard(x);
    NCDE_REQUIRE(x && t && n_series >= 0 && L >= 2 && C >= 1, NCDE_ERR_INVALID, "linear_fill_missing: bad arguments");
    if (n_series == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    static const bool allow_staged = getenv("NCDE_RAGGED_NO_STAGING") == nullptr;
    const size_t smem = (size_t)L * 128 * ((dtype == NCDE_F64 ? 8 : 4) + 2);
    if (allow_staged && smem <= 200 * 1024 && L < 65535) {
        DISPATCH_DTYPE(dtype, {
            NCDE_CUDA_OK(cudaFuncSetAttribute(linear_fill_staged_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            linear_fill_staged_kernel<T><<<grid_for(n_series * C, 128), 128, smem, st>>>((T*)x, (const T*)t, n_series, L, C);
        });
        NCDE_CUDA_OK(cudaGetLastError());
        return NCDE_OK;
    }
    DISPATCH_DTYPE(dtype, (linear_fill_kernel<T><<<grid_for(n_series * C, 128), 128, 0, st>>>(
                              (T*)x, (const T*)t, n_series, L, C)));
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

extern "C" int ncde_linear_derivs(int dtype, const void* coeffs, const void* t, void* derivs, int64_t n_series,
                                  int64_t K, int64_t C, void* stream) {
    ncde::DeviceGuard device_guard(coeffs);
    NCDE_REQUIRE(coeffs && t && derivs && K >= 2 && C >= 1, NCDE_ERR_INVALID, "linear_derivs: bad arguments");
    if (n_series == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, (linear_derivs_kernel<T><<<grid_for(n_series * (K - 1) * C, 256), 256, 0, st>>>(
                              (const T*)coeffs, (const T*)t, (T*)derivs, n_series, K, C)));
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

extern "C" int ncde_path_eval(int kind, int dtype, const void* coeffs, const void* derivs, const void* knots,
                              int64_t n_series, int64_t K, int64_t C, const void* tq, int64_t n_t, int deriv,
                              void* out, int64_t* index_out, void* stream) {
    ncde::DeviceGuard device_guard(coeffs);
    NCDE_REQUIRE(coeffs && knots && tq && out && K >= 2 && C >= 1, NCDE_ERR_INVALID, "path_eval: bad arguments");
    NCDE_REQUIRE(kind == NCDE_PATH_LINEAR || kind == NCDE_PATH_CUBIC, NCDE_ERR_INVALID, "path_eval: bad kind");
    NCDE_REQUIRE(!(kind == NCDE_PATH_LINEAR && deriv && !derivs), NCDE_ERR_INVALID,
                 "path_eval: linear derivative needs derivs");
    if (n_series == 0 || n_t == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, (path_eval_kernel<T><<<grid_for(n_series * n_t * C, 256), 256, 0, st>>>(
                              kind, (const T*)coeffs, (const T*)derivs, (const T*)knots, n_series, K, C,
                              (const T*)tq, n_t, deriv, (T*)out, index_out)));
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

static size_t ragged_scratch_layout(int dtype, int method, int64_t n, int64_t Lmax, int64_t C, size_t* off_t, size_t* off_cubic) {
    const size_t el = dtype == NCDE_F64 ? 8 : 4;
    size_t off = round_up((size_t)n * Lmax * C * el, 256);          // work copy
    *off_t = off;
    off += round_up((size_t)(2 * Lmax) * el, 256);                  // t = 0, 1, ...
    *off_cubic = off;
    if (method == NCDE_RAGGED_CUBIC) off += ncde_cubic_scratch_bytes(dtype, n, Lmax, C);
    return off + 256;
}

extern "C" size_t ncde_ragged_scratch_bytes(int method, int dtype, int64_t n_series, int64_t Lmax, int64_t C) {
    size_t a, b;
    return ragged_scratch_layout(dtype, method, n_series, Lmax, C, &a, &b);
}

extern "C" int ncde_ragged_interpolate(int method, int dtype, const void* x, const int32_t* lengths, void* out,
                                       int64_t n_series, int64_t Lmax, int64_t C, int time_index, int initial_nan_to_zero,
                                       int intensity, int pad, void* scratch, int32_t* flags, void* stream) {
    ncde::DeviceGuard device_guard(x);
    NCDE_REQUIRE(x && lengths && out && scratch, NCDE_ERR_INVALID, "ragged_interpolate: null pointer");
    NCDE_REQUIRE(method == NCDE_RAGGED_LINEAR || method == NCDE_RAGGED_RECTILINEAR || method == NCDE_RAGGED_CUBIC,
                 NCDE_ERR_INVALID, "ragged_interpolate: unknown method %d", method);
    NCDE_REQUIRE(n_series >= 0 && Lmax >= 2 && C >= 1, NCDE_ERR_INVALID, "Must have a time dimension of size at least 2.");
    NCDE_REQUIRE(method != NCDE_RAGGED_RECTILINEAR || (time_index >= 0 && time_index < C), NCDE_ERR_INVALID,
                 "Time index must be in [0, %lld], was given %d.", (long long)C - 1, time_index);
    NCDE_REQUIRE(!intensity || (method == NCDE_RAGGED_RECTILINEAR && time_index == 0), NCDE_ERR_INVALID,
                 "ragged_interpolate: intensity channels belong to the rectilinear scheme with the time channel first");
    if (n_series == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    size_t off_t, off_cubic;
    ragged_scratch_layout(dtype, method, n_series, Lmax, C, &off_t, &off_cubic);
    const int64_t Kmax = 2 * Lmax;
    const int64_t n_init = n_series * Lmax * C > Kmax ? n_series * Lmax * C : Kmax;
    static const bool allow_staged = getenv("NCDE_RAGGED_NO_STAGING") == nullptr;   // A/B switch for tools/bench_ragged.py
    const size_t el = dtype == NCDE_F64 ? 8 : 4;
    const int bs = 128;
    const size_t smem = (size_t)Lmax * bs * el;
    // measured (profiles/README.md): staging pays for linear (0.59 -> 0.32 ms before the uniform fill) and rectilinear
    // (0.54 -> 0.19 ms); the cubic kernel is bound by its Thomas-sweep scratch, not by the input walk (1.29 -> 1.59 ms), so it stays
    if (allow_staged && method != NCDE_RAGGED_CUBIC && smem + (size_t)Lmax * bs * 2 <= 200 * 1024 && Lmax < 65535) {
        DISPATCH_DTYPE(dtype, {
            T* t = (T*)((char*)scratch + off_t);
            arange_kernel<T><<<grid_for(Kmax, 256), 256, 0, st>>>(t, Kmax);
            if (method == NCDE_RAGGED_LINEAR) {
                const size_t smem_l = smem + (size_t)Lmax * bs * 2;   // + the 16-bit previous-observation column
                NCDE_CUDA_OK(cudaFuncSetAttribute(ragged_linear_staged_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
                ragged_linear_staged_kernel<T><<<grid_for(n_series * C, bs), bs, smem_l, st>>>((const T*)x, lengths, t, (T*)out, n_series,
                                                                                             Lmax, C, initial_nan_to_zero, pad);
            } else if (method == NCDE_RAGGED_RECTILINEAR) {
                const int64_t Cout = C + (intensity ? C - 1 : 0);
                NCDE_CUDA_OK(cudaFuncSetAttribute(ragged_rectilinear_staged_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ragged_rectilinear_staged_kernel<T><<<grid_for(n_series * C, bs), bs, smem, st>>>(
                    (const T*)x, lengths, t, (T*)out, n_series, Lmax, C, Cout, time_index, initial_nan_to_zero, pad, flags);
            } else {
                const size_t n = (size_t)n_series * (size_t)C * (size_t)Lmax;
                T* sx = (T*)((char*)scratch + off_cubic);
                T* sd = sx + n;
                T* sr = sd + n;
                int32_t* si = (int32_t*)(sr + n);
                NCDE_CUDA_OK(cudaFuncSetAttribute(ragged_cubic_staged_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ragged_cubic_staged_kernel<T><<<grid_for(n_series * C, bs), bs, smem, st>>>((const T*)x, lengths, t, (T*)out, n_series, Lmax,
                                                                                            C, initial_nan_to_zero, pad, sx, sd, sr, si);
            }
        });
        NCDE_CUDA_OK(cudaGetLastError());
        return NCDE_OK;
    }
    DISPATCH_DTYPE(dtype, {
        T* work = (T*)scratch;
        T* t = (T*)((char*)scratch + off_t);
        ragged_init_kernel<T><<<grid_for(n_init, 256), 256, 0, st>>>((const T*)x, work, t, n_series, Lmax, C, Kmax,
                                                                      initial_nan_to_zero);
        if (method == NCDE_RAGGED_LINEAR) {
            ragged_linear_kernel<T><<<grid_for(n_series * C, 128), 128, 0, st>>>(work, lengths, t, (T*)out, n_series, Lmax, C, pad);
        } else if (method == NCDE_RAGGED_RECTILINEAR) {
            const int64_t Cout = C + (intensity ? C - 1 : 0);
            ragged_rectilinear_kernel<T><<<grid_for(n_series * C, 128), 128, 0, st>>>((const T*)x, work, lengths, t, (T*)out, n_series,
                                                                                      Lmax, C, Cout, time_index, pad, flags);
        } else {
            const size_t n = (size_t)n_series * (size_t)C * (size_t)Lmax;
            T* sx = (T*)((char*)scratch + off_cubic);
            T* sd = sx + n;
            T* sr = sd + n;
            int32_t* si = (int32_t*)(sr + n);
            ragged_cubic_kernel<T><<<grid_for(n_series * C, 128), 128, 0, st>>>(work, lengths, t, (T*)out, n_series, Lmax, C, pad,
                                                                                sx, sd, sr, si);
        }
    });
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

extern "C" int ncde_path_eval_bwd(int kind, int dtype, const void* knots, int64_t n_series, int64_t K, int64_t C,
                                  const void* tq, int64_t n_t, int deriv, const void* grad_out, void* grad_coeffs,
                                  void* stream) {
    ncde::DeviceGuard device_guard(knots);
    NCDE_REQUIRE(knots && tq && grad_out && grad_coeffs && K >= 2 && C >= 1, NCDE_ERR_INVALID, "path_eval_bwd: bad arguments");
    NCDE_REQUIRE(kind == NCDE_PATH_LINEAR || kind == NCDE_PATH_CUBIC, NCDE_ERR_INVALID, "path_eval_bwd: bad kind");
    if (n_series == 0 || n_t == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, (path_eval_bwd_kernel<T><<<grid_for(n_series * C, 128), 128, 0, st>>>(
                              kind, (const T*)knots, n_series, K, C, (const T*)tq, n_t, deriv, (const T*)grad_out,
                              (T*)grad_coeffs)));
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Depth-<=3 log-signatures over windows of a piecewise-linear path, cumulatively summed (log-ODE transform).
// Replaces the signatory.Logsignature(depth) + stack + cumsum part of torchcde.log_ode._logsignature_windows
// (modules/torchcde/torchcde/log_ode.py:49-70).  x is the filled path (n_series, Lp, d); window w covers rows
// idx[w] .. idx[w+1] inclusive.  Output row 0 is the "first increment" x[:, 0, :] padded with zeros (:53-55), row w+1 the
// running sum of the window log-signatures.  Channel order = Signatory's default "words" mode: the coefficients of
// log(signature), taken in the tensor algebra, at the Lyndon words of length 1, 2, 3 (each length in lexicographic order):
//   length 1  (i)       the increment
//   length 2  (i, j), i < j                                  L2_ij  = S2_ij - S1_i S1_j / 2            (= the Levy area)
//   length 3  (i, j, k), j >= i, k > i                       L3_ijk = S3_ijk - (S1_i S2_jk + S2_ij S1_k) / 2 + S1_i S1_j S1_k / 3
// with S the signature of the window, built segment by segment with Chen's identity (exp of a linear segment D is
// 1 + D + D(x)D / 2 + D(x)D(x)D / 6):  S3_ijk += S2_ij D_k + S1_i D_j D_k / 2 + D_i D_j D_k / 6,  S2_ij += S1_i D_j + D_i D_j / 2.
// One thread per (series, output channel) carries only the signature entries its word needs and walks the windows in order,
// so the cumulative sum needs no second pass.
// wscale (device, W entries, nullable) multiplies each window's log-signature before the sum (_version 0: window duration).
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int logsig_channels(int d, int depth) {
    int ch = d;
    if (depth >= 2) ch += d * (d - 1) / 2;
    if (depth >= 3) ch += (d * d * d - d) / 3;
    return ch;
}
template <typename T>
__global__ void logsig_windows_kernel(const T* __restrict__ x, const int32_t* __restrict__ idx, const T* __restrict__ wscale,
                                      T* __restrict__ out, int64_t n_series, int64_t Lp, int d, int depth, int W) {
    const int ch = logsig_channels(d, depth);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * ch) return;
    const int64_t s = tid / ch;
    const int c = (int)(tid % ch);
    const T* xs = x + s * Lp * d;
    T* os = out + s * (int64_t)(W + 1) * ch;
    const int n2 = d * (d - 1) / 2;
    int len = 1, i = c, j = 0, k = 0;
    if (c >= d + n2) {          // Lyndon word (i, j, k): first letter i has (d - i)(d - 1 - i) words: j in [i, d), k in (i, d)
        len = 3;
        int r = c - d - n2;
        i = 0;
        while (r >= (d - i) * (d - 1 - i)) { r -= (d - i) * (d - 1 - i); ++i; }
        j = i + r / (d - 1 - i);
        k = i + 1 + r % (d - 1 - i);
    } else if (c >= d) {        // Lyndon word (i, j): c - d = i (2d - i - 1) / 2 + (j - i - 1)
        len = 2;
        int r = c - d;
        i = 0;
        while (r >= d - 1 - i) { r -= d - 1 - i; ++i; }
        j = i + 1 + r;
    }
    T run = len == 1 ? xs[i] : (T)0;
    os[c] = run;
    for (int w = 0; w < W; ++w) {
        const int lo = idx[w], hi = idx[w + 1];
        T v;
        if (len == 1) {
            v = xs[(int64_t)hi * d + i] - xs[(int64_t)lo * d + i];
        } else if (len == 2) {
            const T x0i = xs[(int64_t)lo * d + i], x0j = xs[(int64_t)lo * d + j];
            T acc = 0;
            T pi = x0i, pj = x0j;
            for (int q = lo; q < hi; ++q) {
                const T ni = xs[(int64_t)(q + 1) * d + i], nj = xs[(int64_t)(q + 1) * d + j];
                acc += (pi - x0i) * (nj - pj) - (pj - x0j) * (ni - pi);
                pi = ni; pj = nj;
            }
            v = (T)0.5 * acc;
        } else {
            T pi = xs[(int64_t)lo * d + i], pj = xs[(int64_t)lo * d + j], pk = xs[(int64_t)lo * d + k];
            T s1i = 0, s1j = 0, s1k = 0, s2ij = 0, s2jk = 0, s3 = 0;
            for (int q = lo; q < hi; ++q) {
                const T ni = xs[(int64_t)(q + 1) * d + i], nj = xs[(int64_t)(q + 1) * d + j], nk = xs[(int64_t)(q + 1) * d + k];
                const T di = ni - pi, dj = nj - pj, dk = nk - pk;
                s3 += s2ij * dk + s1i * dj * dk * (T)0.5 + di * dj * dk * (T)(1.0 / 6.0);
                s2ij += s1i * dj + di * dj * (T)0.5;
                s2jk += s1j * dk + dj * dk * (T)0.5;
                s1i += di; s1j += dj; s1k += dk;
                pi = ni; pj = nj; pk = nk;
            }
            v = s3 - (T)0.5 * (s1i * s2jk + s2ij * s1k) + s1i * s1j * s1k * (T)(1.0 / 3.0);
        }
        if (wscale) v *= wscale[w];
        run += v;
        os[(int64_t)(w + 1) * ch + c] = run;
    }
}

extern "C" int ncde_logsig_windows(int dtype, const void* x, const int32_t* idx, const void* wscale, void* out, int64_t n_series,
                                   int64_t Lp, int d, int depth, int W, void* stream) {
    ncde::DeviceGuard device_guard(x);
    NCDE_REQUIRE(x && idx && out && Lp >= 1 && d >= 1 && W >= 0, NCDE_ERR_INVALID, "logsig_windows: bad arguments");
    NCDE_REQUIRE(depth >= 1 && depth <= 3, NCDE_ERR_UNSUPPORTED, "logsig_windows: depth %d is not implemented (1, 2 and 3 are)", depth);
    NCDE_REQUIRE(d <= 1024, NCDE_ERR_UNSUPPORTED, "logsig_windows: %d channels > 1024 not supported", d);
    if (n_series == 0) return NCDE_OK;
    const int ch = logsig_channels(d, depth);
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, (logsig_windows_kernel<T><<<grid_for(n_series * ch, 128), 128, 0, st>>>(
                              (const T*)x, idx, (const T*)wscale, (T*)out, n_series, Lp, d, depth, W)));
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Linear / rectilinear hybrid (src/ncde/interpolation.py:186-253, steps after the two linear_interpolation_coeffs calls):
// shift the linearly interpolated channels up by one row, keep a row only if a time / rectilinear channel differs from
// the previous row, pad every series to the common length by repeating its last kept row (= NaN padding + forward fill).
// One CTA per series walks the rows in order; lanes = channels, so every row is one coalesced access.
//   chan_kind[c] = 1: time or rectilinear channel (decides whether a row is kept), 0: linear channel (shifted)
//   out (n_series, K, C): rows [0, counts[s]) are the kept rows, the rest repeats the last one; the caller slices to max(counts)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void hybrid_compact_kernel(const T* __restrict__ full, const int32_t* __restrict__ chan_kind, T* __restrict__ out,
                                      int32_t* __restrict__ counts, int64_t K, int64_t C) {
    const int64_t s = blockIdx.x;
    const T* fs = full + s * K * C;
    T* os = out + s * K * C;
    int pos = 0;
    for (int64_t r = 0; r < K; ++r) {
        int differs = r == 0;
        if (r > 0)
            for (int64_t c = threadIdx.x; c < C; c += blockDim.x)
                if (chan_kind[c] == 1 && fs[(r - 1) * C + c] != fs[r * C + c]) differs = 1;   // NaN != NaN counts as a change, like `deltas != 0`
        if (__syncthreads_or(differs)) {
            const int64_t rs = r + 1 < K ? r + 1 : K - 1;
            for (int64_t c = threadIdx.x; c < C; c += blockDim.x) os[(int64_t)pos * C + c] = fs[(chan_kind[c] == 0 ? rs : r) * C + c];
            ++pos;
        }
    }
    __syncthreads();
    for (int64_t r = pos; r < K; ++r)
        for (int64_t c = threadIdx.x; c < C; c += blockDim.x) os[r * C + c] = os[(int64_t)(pos - 1) * C + c];
    if (threadIdx.x == 0) counts[s] = pos;
}

extern "C" int ncde_hybrid_compact(int dtype, const void* full, const int32_t* chan_kind, void* out, int32_t* counts,
                                   int64_t n_series, int64_t K, int64_t C, void* stream) {
    ncde::DeviceGuard device_guard(full);
    NCDE_REQUIRE(full && chan_kind && out && counts && K >= 1 && C >= 1, NCDE_ERR_INVALID, "hybrid_compact: bad arguments");
    NCDE_REQUIRE(n_series < (1ll << 31), NCDE_ERR_UNSUPPORTED, "hybrid_compact: too many series");
    if (n_series == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, (hybrid_compact_kernel<T><<<(unsigned)n_series, 128, 0, st>>>((const T*)full, chan_kind, (T*)out, counts, K, C)));
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// SmoothLinearInterpolation — src/ncde/interpolation.py:6-183.  The scalar factors (1/eps^2, 1/(3 eps^2), ...) are Python
// doubles in the reference and are rounded to the tensor dtype when they multiply it; they arrive here as doubles and are
// cast the same way.  Operation order follows the reference expression by expression (this file is compiled with -fmad=false).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void smooth_matching_kernel(const T* __restrict__ coeffs, T* __restrict__ out, int64_t n_series, int64_t K, int64_t C,
                                       double eps_d, int terms) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * (K - 2) * C) return;
    const int64_t c = tid % C, k = (tid / C) % (K - 2), s = tid / (C * (K - 2));
    const T* cs = coeffs + s * K * C + c;
    const T prev = cs[k * C], x = cs[(k + 1) * C], next = cs[(k + 2) * C];
    const T eps = (T)eps_d;
    const T x_eps = x + eps * (next - x);
    const T delta_prev = x - prev, delta_next = next - x;
    T* o = out + tid * terms;
    if (terms == 4) {
        const T Cc = delta_prev, D = x;
        const T B = (T)(1.0 / (eps_d * eps_d)) * ((T)3 * (x_eps - Cc * eps - D) - eps * (delta_next - Cc));
        const T A = (T)(1.0 / (3.0 * (eps_d * eps_d))) * (delta_next - Cc - (T)2 * B * eps);
        o[0] = A; o[1] = B; o[2] = Cc; o[3] = D;
    } else {
        const T D = (T)0, E = delta_prev, F = x;
        const T Cc = (T)(1.0 / (eps_d * eps_d * eps_d)) * ((T)10 * (x_eps - E * eps - F) - (T)4 * eps * (delta_next - E));
        const T B = (T)(1.0 / (2.0 * (eps_d * eps_d * eps_d))) * ((T)2 * (delta_next - E) - (T)3 * Cc * (T)(eps_d * eps_d));
        const T A = -(T)(1.0 / (10.0 * (eps_d * eps_d))) * ((T)6 * B * eps + (T)3 * Cc);
        o[0] = A; o[1] = B; o[2] = Cc; o[3] = D; o[4] = E; o[5] = F;
    }
}

// polynomial of the matching region, highest power first (interpolation.py:126-143): sum_i m[i] t^(terms-1-i), or its derivative
template <typename T>
__device__ __forceinline__ T smooth_poly(const T* __restrict__ m, int terms, T t, int deriv) {
    T acc = 0;
    if (deriv) {
        for (int i = 0; i < terms - 1; ++i) {
            const int pw = terms - 1 - i;          // term pw * t^(pw-1)
            T tp = 1;
            for (int q = 0; q < pw - 1; ++q) tp *= t;
            acc += m[i] * ((T)pw * tp);
        }
    } else {
        for (int i = 0; i < terms; ++i) {
            const int pw = terms - 1 - i;
            T tp = 1;
            for (int q = 0; q < pw; ++q) tp *= t;
            acc += m[i] * tp;
        }
    }
    return acc;
}

template <typename T>
__global__ void path_eval_smooth_kernel(const T* __restrict__ coeffs, const T* __restrict__ derivs, const T* __restrict__ knots,
                                        const T* __restrict__ match, int terms, T eps, int64_t n_series, int64_t K, int64_t C,
                                        const T* __restrict__ tq, int64_t n_t, int deriv, T* __restrict__ out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_series * n_t * C) return;
    const int64_t c = tid % C, q = (tid / C) % n_t, s = tid / (C * n_t);
    const T t = tq[q];
    const int idx = knot_index<T>(knots, (int)K, t);
    const T frac = t - knots[idx];
    T r;
    if (match && idx > 0 && frac < eps) {
        r = smooth_poly<T>(match + ((s * (K - 2) + idx - 1) * C + c) * terms, terms, frac, deriv);
    } else if (deriv) {
        r = derivs[(s * (K - 1) + idx) * C + c];
    } else {
        const T lo = coeffs[(s * K + idx) * C + c], hi = coeffs[(s * K + idx + 1) * C + c];
        r = lo + frac * (hi - lo) / (knots[idx + 1] - knots[idx]);
    }
    out[tid] = r;
}

extern "C" int ncde_smooth_matching_coeffs(int dtype, const void* coeffs, void* out, int64_t n_series, int64_t K, int64_t C,
                                           double eps, int terms, void* stream) {
    ncde::DeviceGuard device_guard(coeffs);
    NCDE_REQUIRE(coeffs && out && K >= 3 && C >= 1, NCDE_ERR_INVALID, "smooth_matching_coeffs: bad arguments");
    NCDE_REQUIRE(terms == 4 || terms == 6, NCDE_ERR_INVALID, "smooth_matching_coeffs: terms must be 4 (cubic) or 6 (quintic)");
    NCDE_REQUIRE(eps > 0 && eps <= 1, NCDE_ERR_INVALID, "gradient_matching_eps must be in (0, 1]");
    if (n_series == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, (smooth_matching_kernel<T><<<grid_for(n_series * (K - 2) * C, 256), 256, 0, st>>>(
                              (const T*)coeffs, (T*)out, n_series, K, C, eps, terms)));
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

extern "C" int ncde_path_eval_smooth(int dtype, const void* coeffs, const void* derivs, const void* knots, const void* match,
                                     int terms, double eps, int64_t n_series, int64_t K, int64_t C, const void* tq, int64_t n_t,
                                     int deriv, void* out, void* stream) {
    ncde::DeviceGuard device_guard(coeffs);
    NCDE_REQUIRE(coeffs && derivs && knots && tq && out && K >= 2 && C >= 1, NCDE_ERR_INVALID, "path_eval_smooth: bad arguments");
    NCDE_REQUIRE(!match || terms == 4 || terms == 6, NCDE_ERR_INVALID, "path_eval_smooth: terms must be 4 or 6");
    if (n_series == 0 || n_t == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, (path_eval_smooth_kernel<T><<<grid_for(n_series * n_t * C, 256), 256, 0, st>>>(
                              (const T*)coeffs, (const T*)derivs, (const T*)knots, (const T*)match, terms, (T)eps, n_series, K, C,
                              (const T*)tq, n_t, deriv, (T*)out)));
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}

extern "C" size_t ncde_cubic_scratch_bytes(int dtype, int64_t n_series, int64_t L, int64_t C) {
    size_t el = dtype == NCDE_F64 ? 8 : 4;
    size_t n = (size_t)n_series * (size_t)C * (size_t)L;
    return 3 * n * el + n * sizeof(int32_t) + 1024;
}

extern "C" int ncde_natural_cubic_coeffs(int dtype, const void* x, const void* t, void* out, int64_t n_series,
                                         int64_t L, int64_t C, int version, void* scratch, void* stream) {
    ncde::DeviceGuard device_guard(x);
    NCDE_REQUIRE(x && t && out && scratch, NCDE_ERR_INVALID, "natural_cubic_coeffs: null pointer");
    NCDE_REQUIRE(L >= 2 && C >= 1 && (version == 0 || version == 1), NCDE_ERR_INVALID,
                 "Must have a time dimension of size at least 2.");
    if (n_series == 0) return NCDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    size_t n = (size_t)n_series * (size_t)C * (size_t)L;
    DISPATCH_DTYPE(dtype, {
        T* sx = (T*)scratch;
        T* sd = sx + n;
        T* sr = sd + n;
        int32_t* si = (int32_t*)(sr + n);
        cubic_coeffs_kernel<T><<<grid_for(n_series * C, 128), 128, 0, st>>>((const T*)x, (const T*)t, (T*)out,
                                                                             n_series, L, C, version, sx, sd, sr, si);
    });
    NCDE_CUDA_OK(cudaGetLastError());
    return NCDE_OK;
}
