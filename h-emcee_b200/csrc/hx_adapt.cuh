// hx_adapt.cuh — device-resident warm-up adaptation.
//
// The reference re-reads (step, L) from its adapter before every half-move and updates the adapter after the move and
// before the accept (samplers/hamiltonian_ensemble.py:165-217).  Driving that from the host costs a round trip per
// half-move; here the adapter state lives in device memory, the move kernels read (step, L) from it, and three small
// kernels per half-move reproduce, formula by formula and in the run's float type,
//   DualAveragingAdapter.update   adaptation/dual_averaging.py:64-89 (harmonic / mean accept statistic :43-50)
//   ChEESAdapter.update           adaptation/chees.py:89-166 (inner product through cov(group2)^-1, :54-73)
//   CompositeAdapter.value        adaptation/adapter.py:64-70, ChEESAdapter._get_T chees.py:182-190, get_halton :193-219
// cov(group2) = n/(n-1) * M with the Gram matrix M of the centred complement the walk move already forms;
// solve(cov, p_proj^T) uses a Cholesky factor computed by one CTA (dim <= kAdaptMaxDim; larger problems keep the
// host-driven path).  All reductions run in a fixed order, so warm-up trajectories are bit-reproducible.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/hx.h"
#include "hx_common.cuh"

namespace hx {

constexpr int kAdaptMaxDim = 128;

template <typename T>
struct AdaptDev {
  int da_it;
  T da_ls, da_lsb, da_H;
  T ch_lT, ch_lTb, ch_m1, ch_m2, ch_it;
  float ch_halton;
  T eps;        // value(): step size of the next half-move
  int L;        // value(): leapfrog steps of the next half-move
};

__device__ __forceinline__ float halton2_f32(float index) {   // chees.py:193-219, base 2, accumulated in float32
  int i = (int)index;
  float f = 1.0f, r = 0.0f;
  while (i > 0) {
    f = f / 2.0f;
    r = r + f * (float)(i % 2);
    i /= 2;
  }
  return r;
}

template <typename T> __device__ __forceinline__ T exp_t(T x);
template <> __device__ __forceinline__ float exp_t<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double exp_t<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T logf_t(T x);
template <> __device__ __forceinline__ float logf_t<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double logf_t<double>(double x) { return log(x); }
template <typename T> __device__ __forceinline__ T pow_t(T x, T y);
template <> __device__ __forceinline__ float pow_t<float>(float x, float y) { return powf(x, y); }
template <> __device__ __forceinline__ double pow_t<double>(double x, double y) { return pow(x, y); }
template <typename T> __device__ __forceinline__ T sqrt_t2(T x);
template <> __device__ __forceinline__ float sqrt_t2<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double sqrt_t2<double>(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ T ceil_t(T x);
template <> __device__ __forceinline__ float ceil_t<float>(float x) { return ceilf(x); }
template <> __device__ __forceinline__ double ceil_t<double>(double x) { return ceil(x); }
template <typename T> __device__ __forceinline__ T log1p_t(T x);
template <> __device__ __forceinline__ float log1p_t<float>(float x) { return log1pf(x); }
template <> __device__ __forceinline__ double log1p_t<double>(double x) { return log1p(x); }

// value(): (step, L) of the next half-move from the adapter state
template <typename T>
__device__ __forceinline__ void adapt_value(AdaptDev<T>& st, const hx_adapt_params& p) {
  T step = (T)p.pass_step;
  int L = p.pass_L;
  if (p.use_da) step = exp_t<T>(st.da_ls);                                      // dual_averaging.py:91-92
  if (p.use_chees) {
    const T Tt = exp_t<T>(st.ch_lT);
    const T Tj = (T)(1.0 - p.ch_jitter) * Tt + ((T)p.ch_jitter * (T)st.ch_halton) * Tt;   // chees.py:182-190
    const T q = ceil_t<T>(Tj / step);                                           // adapter.py:64-70
    L = q > (T)1 ? (q < (T)2.0e9 ? (int)q : 2000000000) : 1;
  }
  st.eps = step;
  st.L = L;
}

template <typename T>
__global__ void k_adapt_value(AdaptDev<T>* st, hx_adapt_params p) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    AdaptDev<T> s = *st;
    adapt_value<T>(s, p);
    *st = s;
  }
}

// Cholesky factor (lower, row-major [d, ldl]) of cov = scale * M, M = [d, ldm] symmetric.  One CTA.  A non-positive pivot
// produces NaNs, which the ChEES gradient mask then zeroes (the reference's LU solve yields inf/NaN there as well).
template <typename T>
__global__ void __launch_bounds__(256) k_chol(const T* __restrict__ M, int ldm, int d, T scale, T* __restrict__ Lf, int ldl) {
  for (int idx = threadIdx.x; idx < d * d; idx += blockDim.x) {
    const int i = idx / d, j = idx - i * d;
    Lf[(size_t)i * ldl + j] = (j <= i) ? scale * M[(size_t)i * ldm + j] : T(0);
  }
  __syncthreads();
  for (int k = 0; k < d; ++k) {
    if (threadIdx.x == 0) Lf[(size_t)k * ldl + k] = sqrt_t2<T>(Lf[(size_t)k * ldl + k]);
    __syncthreads();
    const T piv = Lf[(size_t)k * ldl + k];
    for (int i = k + 1 + threadIdx.x; i < d; i += blockDim.x) Lf[(size_t)i * ldl + k] /= piv;
    __syncthreads();
    // trailing update of the lower triangle: A[i][j] -= L[i][k] * L[j][k], k < j <= i
    const int m = d - k - 1;
    for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
      const int i = k + 1 + idx / m, j = k + 1 + idx % m;
      if (j <= i) Lf[(size_t)i * ldl + j] -= Lf[(size_t)i * ldl + k] * Lf[(size_t)j * ldl + k];
    }
    __syncthreads();
  }
}

// Per-walker adapter statistics (thread per walker):
//   inv_a[i]  = 1 / clip(exp(la_i), 0, 1)  (harmonic) or clip(exp(la_i), 0, 1) (mean)          dual_averaging.py:43-50
//   acc[i]    = clip(exp(la_i), 0, 1);  accg[i] = acc_i * g_i with
//   g_i = t_n * (|cen_prop_i|^2 - |cen_prev_i|^2) * <cen_prop_i, cov^-1 pproj_i>, zeroed when acc_i <= 1e-4 or non-finite
//                                                                                                 chees.py:112-131
template <typename T>
__global__ void __launch_bounds__(128) k_adapt_rows(const T* __restrict__ la, const T* __restrict__ Xcur, const T* __restrict__ Xprop,
                                                    const T* __restrict__ pproj, const T* __restrict__ mean_prev,
                                                    const T* __restrict__ mean_prop, const T* __restrict__ Lf, int ldl, int rows, int d,
                                                    const AdaptDev<T>* st, hx_adapt_params p, T* __restrict__ stat_da,
                                                    T* __restrict__ stat_acc, T* __restrict__ stat_accg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  T a = exp_t<T>(la[i]);
  a = a < T(0) ? T(0) : (a > T(1) ? T(1) : a);
  if (p.use_da) stat_da[i] = p.agg_harmonic ? T(1) / a : a;
  if (!p.use_chees) return;
  const T* xc = Xcur + (size_t)i * d;
  const T* xp = Xprop + (size_t)i * d;
  const T* pp = pproj + (size_t)i * d;
  T y[kAdaptMaxDim];
  T sq_prop = T(0), sq_prev = T(0);
  for (int c = 0; c < d; ++c) {
    const T cp = xp[c] - mean_prop[c], cv = xc[c] - mean_prev[c];
    sq_prop += cp * cp;
    sq_prev += cv * cv;
  }
  // forward substitution L y' = pproj_i, then backward L^T y = y'
  for (int r = 0; r < d; ++r) {
    T sacc = pp[r];
    for (int c = 0; c < r; ++c) sacc -= Lf[(size_t)r * ldl + c] * y[c];
    y[r] = sacc / Lf[(size_t)r * ldl + r];
  }
  for (int r = d - 1; r >= 0; --r) {
    T sacc = y[r];
    for (int c = r + 1; c < d; ++c) sacc -= Lf[(size_t)c * ldl + r] * y[c];
    y[r] = sacc / Lf[(size_t)r * ldl + r];
  }
  T ip = T(0);
  for (int c = 0; c < d; ++c) ip += (xp[c] - mean_prop[c]) * y[c];
  const T Tt = exp_t<T>(st->ch_lT);
  const T t_n = (T)st->ch_halton * Tt;
  T g = t_n * (sq_prop - sq_prev) * ip;
  if (!(a > T(1e-4))) g = T(0);
  if (!isfinite(g)) g = T(0);
  stat_acc[i] = a;
  stat_accg[i] = a * g;
}

// Per-walker adapter statistics beyond kAdaptMaxDim (float runs on the tensor-core path): one warp per walker; the ChEES inner
// product arrives as column-tile partial sums ip_part[i * nt + t] of <pproj_i cov^-1, Xprop_i - mean_prop> (tc::chees_inner).
__global__ void __launch_bounds__(256) k_adapt_rows_big(const float* __restrict__ la, const float* __restrict__ Xcur,
                                                        const float* __restrict__ Xprop, const float* __restrict__ mean_prev,
                                                        const float* __restrict__ mean_prop, const float* __restrict__ ip_part, int nt,
                                                        int rows, int d, const AdaptDev<float>* st, hx_adapt_params p,
                                                        float* __restrict__ stat_da, float* __restrict__ stat_acc,
                                                        float* __restrict__ stat_accg) {
  const int lane = threadIdx.x & 31, i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= rows) return;
  float a = expf(la[i]);
  a = a < 0.f ? 0.f : (a > 1.f ? 1.f : a);
  if (p.use_da && lane == 0) stat_da[i] = p.agg_harmonic ? 1.f / a : a;
  if (!p.use_chees) return;
  const float* xc = Xcur + (size_t)i * d;
  const float* xp = Xprop + (size_t)i * d;
  float sq_prop = 0.f, sq_prev = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float cp = xp[c] - mean_prop[c], cv = xc[c] - mean_prev[c];
    sq_prop = fmaf(cp, cp, sq_prop);
    sq_prev = fmaf(cv, cv, sq_prev);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sq_prop += __shfl_xor_sync(0xffffffffu, sq_prop, o);
    sq_prev += __shfl_xor_sync(0xffffffffu, sq_prev, o);
  }
  if (lane != 0) return;
  float ip = 0.f;
  for (int t = 0; t < nt; ++t) ip += ip_part[(size_t)i * nt + t];
  const float Tt = expf(st->ch_lT);
  const float t_n = st->ch_halton * Tt;
  float g = t_n * (sq_prop - sq_prev) * ip;                                     // chees.py:112-131
  if (!(a > 1e-4f)) g = 0.f;
  if (!isfinite(g)) g = 0.f;
  stat_acc[i] = a;
  stat_accg[i] = a * g;
}

// fixed-order sum of v[0..n) by one CTA of 1024 threads: strided per-thread partials, then a shared-memory tree
template <typename T>
__device__ __forceinline__ T block_sum_fixed(const T* __restrict__ v, int n, T* sh) {
  T s = T(0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = blockDim.x >> 1; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  const T r = sh[0];
  __syncthreads();
  return r;
}

// DA + ChEES scalar updates, then value() for the next half-move.  One CTA of 1024 threads.
template <typename T>
__global__ void __launch_bounds__(1024) k_adapt_update(AdaptDev<T>* st, hx_adapt_params p, const T* __restrict__ stat_da,
                                                       const T* __restrict__ stat_acc, const T* __restrict__ stat_accg, int rows) {
  __shared__ T sh[1024];
  T s_da = T(0), s_acc = T(0), s_accg = T(0);
  if (p.use_da) s_da = block_sum_fixed<T>(stat_da, rows, sh);
  if (p.use_chees) {
    s_acc = block_sum_fixed<T>(stat_acc, rows, sh);
    s_accg = block_sum_fixed<T>(stat_accg, rows, sh);
  }
  if (threadIdx.x != 0) return;
  AdaptDev<T> s = *st;
  if (p.use_da) {                                                               // dual_averaging.py:64-89
    const int it = s.da_it + 1;
    const T rate = p.agg_harmonic ? (T)rows / s_da : s_da / (T)rows;
    T eta = T(1) / (T)((T)it + (T)p.da_t0);
    const T H_bar = (T(1) - eta) * s.da_H + eta * ((T)p.da_target_accept - rate);
    const T soft_t = (T)((T)it + (T)p.da_t0);
    const T log_eps = logf_t<T>((T)p.da_init_step) - (sqrt_t2<T>((T)it) / (soft_t * (T)p.da_gamma)) * H_bar;
    eta = pow_t<T>((T)it, (T)(-p.da_kappa));
    const T log_eps_bar = eta * log_eps + (T(1) - eta) * s.da_lsb;
    s.da_it = it; s.da_ls = log_eps; s.da_lsb = log_eps_bar; s.da_H = H_bar;
  }
  if (p.use_chees) {                                                            // chees.py:132-166
    const T g_hat = s_accg / (s_acc + (T)p.ch_reg);
    const T iteration = s.ch_it + T(1);
    const T m1 = (T)p.ch_beta1 * s.ch_m1 + (T)(1.0 - p.ch_beta1) * g_hat;
    const T m2 = (T)p.ch_beta2 * s.ch_m2 + (T)(1.0 - p.ch_beta2) * g_hat * g_hat;
    const T m1_hat = m1 / (T(1) - pow_t<T>((T)p.ch_beta1, iteration));
    const T m2_hat = m2 / (T(1) - pow_t<T>((T)p.ch_beta2, iteration));
    const T upd = (T)p.ch_lr * m1_hat / sqrt_t2<T>(m2_hat + (T)p.ch_reg);
    const T lo = logf_t<T>((T)p.ch_T_min), hi = logf_t<T>((T)p.ch_T_max);
    T log_T = s.ch_lT + upd;
    log_T = log_T < lo ? lo : (log_T > hi ? hi : log_T);
    const T a1 = logf_t<T>((T)p.ch_T_interp) + s.ch_lTb, a2 = logf_t<T>((T)(1.0 - p.ch_T_interp)) + log_T;
    const T mx = a1 > a2 ? a1 : a2, mn = a1 > a2 ? a2 : a1;
    T log_T_bar = mx + log1p_t<T>(exp_t<T>(mn - mx));                            // logaddexp
    log_T_bar = log_T_bar < lo ? lo : (log_T_bar > hi ? hi : log_T_bar);
    s.ch_lT = log_T; s.ch_lTb = log_T_bar; s.ch_m1 = m1; s.ch_m2 = m2; s.ch_it = iteration;
    s.ch_halton = halton2_f32((float)iteration);
  }
  adapt_value<T>(s, p);
  *st = s;
}

// Metropolis accept + chain emit + non-finite flag, key read from device memory (warm-up loop).  One warp per walker.
template <typename T>
__global__ void __launch_bounds__(256) k_accept_emit(T* cur, const T* __restrict__ prop, T* lp_cur, const T* __restrict__ lp_prop,
                                                     const T* __restrict__ logacc, const uint32_t* __restrict__ key_ptr, int legacy,
                                                     int n, int row0, int rows, int d, int32_t* accepts, T* chain_out, T* chain_lp,
                                                     int32_t* nonfinite) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= rows) return;
  const Key key{key_ptr[0], key_ptr[1]};
  const T lo = T(1e-10);
  const T span = T(1.0) - lo;
  const T u = uniform_rt<T>(key, (uint64_t)(row0 + w), (uint64_t)n, lo, span, legacy);
  const bool acc = log_t(u) < logacc[w];                      // sampler_utils.py:114-115
  const size_t off = (size_t)w * d;
  for (int c = lane; c < d; c += 32) {
    const T v = acc ? prop[off + c] : cur[off + c];
    cur[off + c] = v;
    if (chain_out) chain_out[off + c] = v;
  }
  if (lane == 0) {
    const T lps = acc ? lp_prop[w] : lp_cur[w];
    lp_cur[w] = lps;
    if (chain_lp) chain_lp[w] = lps;
    accepts[w] += acc ? 1 : 0;
    if (!isfinite(lps) && nonfinite) atomicOr(nonfinite, 1);
  }
}

}  // namespace hx
