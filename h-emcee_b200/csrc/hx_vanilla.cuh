// hx_vanilla.cuh — derivative-free affine-invariant ensemble moves:
//   stretch  reference src/hemcee/moves/vanilla/stretch.py:5-76   (Goodman-Weare, Alg. 1 of arXiv:2505.02987)
//   side     reference src/hemcee/moves/vanilla/side.py:5-56      (Alg. 2)
//   walk     reference src/hemcee/moves/vanilla/walk.py:5-42      (Alg. 4: one shared random combination of the complement)
// Same skeleton as the Hamiltonian moves: counter-based draws on global row indices (shard-invariant), per-walker complement
// choices through k_choice2 (the first entry of choice(.., (2,)) is the entry of choice(.., (1,)): same permutation),
// proposal build, batched log-density (k_log_prob), Metropolis accept with the MOVE's key
// (samplers/ensemble.py:103-104: accept_proposal(..., keys[0])).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "hx_common.cuh"
#include "hx_rng.cuh"

namespace hx {

enum : int { kVanillaStretch = 0, kVanillaSide = 1, kVanillaWalk = 2 };

// shift[c] = sum_b (X2[b][c] - mean[c]) * xi_b / sqrt(n), xi = normal(split(key, 2)[0], (n,))     walk.py:29-35
template <typename T>
__global__ void __launch_bounds__(256) k_vanilla_walk_shift(const T* __restrict__ X2, const T* __restrict__ mean, int n, int d, Key key,
                                                            const uint32_t* key_ptr, int legacy, T* __restrict__ shift) {
  if (key_ptr) { key.k0 = key_ptr[0]; key.k1 = key_ptr[1]; }
  const Key kn = split_key_rt(key, 0, 2, legacy);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  T s = T(0);
  for (int b = 0; b < n; ++b) s = fma_t(X2[(size_t)b * d + c] - mean[c], normal_rt<T>(kn, (uint64_t)b, (uint64_t)n, legacy), s);
  shift[c] = s / sqrt_t(T(n));
}

// proposals Q[rows, d] and the extra term of the log-accept ratio ((d-1) log z for the stretch move, 0 otherwise); warp per walker
template <typename T>
__global__ void __launch_bounds__(256) k_vanilla_propose(int kind, const T* __restrict__ X1, const T* __restrict__ X2,
                                                         const int32_t* __restrict__ idx2, Key key, const uint32_t* key_ptr, int legacy,
                                                         int n, int row0, int rows, int d, T stretch, const T* __restrict__ shift,
                                                         T* __restrict__ Q, T* __restrict__ extra) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  if (key_ptr) { key.k0 = key_ptr[0]; key.k1 = key_ptr[1]; }
  const size_t off = (size_t)r * d;
  if (kind == kVanillaStretch) {
    const Key kz = split_key_rt(key, (uint64_t)n + 1, (uint64_t)n + 2, legacy);     // keys[-1]   stretch.py:37
    const T u = uniform_rt<T>(kz, (uint64_t)(row0 + r), (uint64_t)n, T(0), T(1), legacy);
    const T sa = sqrt_t(stretch), isa = T(1) / sa;
    const T t = isa + u * (sa - isa);
    const T z = t * t;                                                                     // stretch.py:72-75
    const T* xj = X2 + (size_t)idx2[2 * r] * d;
    for (int c = lane; c < d; c += 32) Q[off + c] = xj[c] + z * (X1[off + c] - xj[c]);     // stretch.py:49
    if (lane == 0) extra[r] = T(d - 1) * log_t(z);
  } else if (kind == kVanillaSide) {
    const Key kz = split_key_rt(key, (uint64_t)n, (uint64_t)n + 2, legacy);           // keys[-2]   side.py:32
    const T z = normal_rt<T>(kz, (uint64_t)(row0 + r), (uint64_t)n, legacy);
    const T sz = stretch * z;
    const T* xj = X2 + (size_t)idx2[2 * r] * d;
    const T* xk = X2 + (size_t)idx2[2 * r + 1] * d;
    for (int c = lane; c < d; c += 32) Q[off + c] = X1[off + c] + sz * (xj[c] - xk[c]);    // side.py:49
    if (lane == 0) extra[r] = T(0);
  } else {
    for (int c = lane; c < d; c += 32) Q[off + c] = X1[off + c] + shift[c];                // walk.py:36
    if (lane == 0) extra[r] = T(0);
  }
}

// log_accept = extra + lp' - lp (stretch.py:53, side.py:52, walk.py:39); optional accept (key = the move's key) + chain emit
template <typename T>
__global__ void __launch_bounds__(256) k_vanilla_finish(T* X1w, const T* __restrict__ Q, T* lp1w, const T* __restrict__ lp_prop,
                                                        const T* __restrict__ extra, T* __restrict__ logacc, int do_accept, Key key,
                                                        const uint32_t* key_ptr, int legacy, int n, int row0, int rows, int d,
                                                        int32_t* accepts, T* chain_out, T* chain_lp, int32_t* nonfinite) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  if (key_ptr) { key.k0 = key_ptr[0]; key.k1 = key_ptr[1]; }
  const T lp1 = lp1w[r];
  const T la = (extra[r] + lp_prop[r]) - lp1;
  if (lane == 0) logacc[r] = la;
  if (!do_accept) return;
  const T lo = T(1e-10);
  const T span = T(1.0) - lo;
  const T u = uniform_rt<T>(key, (uint64_t)(row0 + r), (uint64_t)n, lo, span, legacy);
  const bool acc = log_t(u) < la;                                  // sampler_utils.py:114-115 (NaN compares false: reject)
  const size_t off = (size_t)r * d;
  for (int c = lane; c < d; c += 32) {
    const T v = acc ? Q[off + c] : X1w[off + c];
    X1w[off + c] = v;
    if (chain_out) chain_out[off + c] = v;
  }
  if (lane == 0) {
    const T lps = acc ? lp_prop[r] : lp1;
    lp1w[r] = lps;
    if (chain_lp) chain_lp[r] = lps;
    accepts[r] += acc ? 1 : 0;
    if (!isfinite(lps) && nonfinite) atomicOr(nonfinite, 1);
  }
}

}  // namespace hx
