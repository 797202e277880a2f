// hx_prep.cuh — complement-ensemble statistics for the walk move, plus small utility kernels.
//
//   mean  = mean_0(X2)                      (moves/hamiltonian/hmc_walk.py:35)
//   C     = (X2 - mean) / sqrt(n)           (moves/hamiltonian/hmc_walk.py:35), zero padded to ldd
//   M     = C^T C                           (projected form)
// All reductions are two-stage with a fixed summation order (no float atomics) so a chain is
// bitwise reproducible run to run.
#pragma once
#include "hx_common.cuh"
#include "hx_rng.cuh"

namespace hx {

constexpr int kColChunk = 64;  // rows per partial column sum

// partial[chunk][c] = sum over the chunk's rows of X[row][c]
template <typename T>
__global__ void k_colsum_partial(const T* __restrict__ X, int n, int d, T* __restrict__ partial) {
  const int chunk = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const int r_lo = chunk * kColChunk, r_hi = min(n, r_lo + kColChunk);
  T s = T(0);
  for (int r = r_lo; r < r_hi; ++r) s += X[(size_t)r * d + c];
  partial[(size_t)chunk * d + c] = s;
}

// 32 columns per CTA (lanes), 8 warps each summing every 8th partial in order; the 8 segment sums are added in a fixed order
template <typename T>
__global__ void __launch_bounds__(256) k_mean_final(const T* __restrict__ partial, int nchunks, int n, int d, T* __restrict__ mean) {
  __shared__ T seg[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  T s = T(0);
  if (c < d)
    for (int k = w; k < nchunks; k += 8) s += partial[(size_t)k * d + c];
  seg[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < d) {
    T t = T(0);
#pragma unroll
    for (int k = 0; k < 8; ++k) t += seg[k][lane];
    mean[c] = t / T(n);
  }
}

// C[r][c] = (X[r][c] - mean[c]) / sqrt(n) for c < d, 0 for d <= c < ldd
template <typename T>
__global__ void k_center(const T* __restrict__ X, const T* __restrict__ mean, int n, int d, int ldd,
                         T* __restrict__ C) {
  const T sc = sqrt_t(T(n));
  const size_t tot = (size_t)n * ldd;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ldd), c = (int)(i - (size_t)r * ldd);
    C[i] = (c < d) ? (X[(size_t)r * d + c] - mean[c]) / sc : T(0);
  }
}

// Gram partials: out[z][i][j] = sum_{r in slice z} C[r][i] * C[r][j]; 64x64 tile per CTA, 4x4 per thread.
template <typename T>
__global__ void __launch_bounds__(256) k_gram(const T* __restrict__ C, int n, int ldd, int rows_per_slice,
                                              T* __restrict__ out) {
  constexpr int TB = 64, KB = 16;
  __shared__ __align__(16) T sa[KB][TB + 4];
  __shared__ __align__(16) T sb[KB][TB + 4];
  const int ti = blockIdx.y * TB, tj = blockIdx.x * TB;
  const int z = blockIdx.z;
  const int r_lo = z * rows_per_slice, r_hi = min(n, r_lo + rows_per_slice);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int k0 = r_lo; k0 < r_hi; k0 += KB) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < KB * TB; idx += 256) {
      const int kk = idx / TB, c = idx - kk * TB;
      const int r = k0 + kk;
      T va = T(0), vb = T(0);
      if (r < r_hi) {
        if (ti + c < ldd) va = C[(size_t)r * ldd + ti + c];
        if (tj + c < ldd) vb = C[(size_t)r * ldd + tj + c];
      }
      sa[kk][c] = va;
      sb[kk][c] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      T a[4], b[4];
      ld4(&sa[kk][4 * ty], a);
      ld4(&sb[kk][4 * tx], b);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma_t(a[i], b[j], acc[i][j]);
    }
  }
  T* o = out + (size_t)z * ldd * ldd;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = ti + 4 * ty + i;
    if (gi >= ldd) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = tj + 4 * tx + j;
      if (gj < ldd) o[(size_t)gi * ldd + gj] = acc[i][j];
    }
  }
}

// Small ensembles (n * ldd <= kPrepSmall): mean, centred complement and Gram matrix by ONE CTA in one launch instead of
// 4-5 launch-latency-bound ones (the whole iteration of a 32-walker problem is ~10 launches).  Fixed summation order.
constexpr int kPrepSmall = 8192;
template <typename T>
__global__ void __launch_bounds__(256) k_prep_small(const T* __restrict__ X, int n, int d, int ldd, T* __restrict__ mean,
                                                    T* __restrict__ C, T* __restrict__ M) {
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    T s = T(0);
    for (int r = 0; r < n; ++r) s += X[(size_t)r * d + c];
    mean[c] = s / T(n);
  }
  __syncthreads();
  const T sc = sqrt_t(T(n));
  for (int i = threadIdx.x; i < n * ldd; i += blockDim.x) {
    const int r = i / ldd, c = i - r * ldd;
    C[i] = (c < d) ? (X[(size_t)r * d + c] - mean[c]) / sc : T(0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ldd * ldd; i += blockDim.x) {
    const int a = i / ldd, b = i - a * ldd;
    T s = T(0);
    for (int r = 0; r < n; ++r) s = fma_t(C[(size_t)r * ldd + a], C[(size_t)r * ldd + b], s);
    M[i] = s;
  }
}

template <typename T>
__global__ void k_sum_slices(const T* __restrict__ part, int nslices, size_t count, T* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    T s = T(0);
    for (int z = 0; z < nslices; ++z) s += part[(size_t)z * count + i];
    out[i] = s;
  }
}

// ---- RNG array kernels (KAT / parity entry points) -----------------------------------------------
template <bool LEGACY>
__global__ void k_split(Key key, uint64_t num, uint32_t* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num; i += (uint64_t)gridDim.x * blockDim.x) {
    const Key k = split_key<LEGACY>(key, i, num);
    out[2 * i] = k.k0;
    out[2 * i + 1] = k.k1;
  }
}
template <bool LEGACY>
__global__ void k_bits32(Key key, uint64_t m, uint32_t* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = bits32<LEGACY>(key, i, m);
}
template <bool LEGACY>
__global__ void k_bits64(Key key, uint64_t m, uint64_t* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = bits64<LEGACY>(key, i, m);
}
template <typename T, bool LEGACY>
__global__ void k_uniform(Key key, uint64_t m, T lo, T span, T* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = uniform_at<T, LEGACY>(key, i, m, lo, span);
}
template <typename T, bool LEGACY>
__global__ void k_normal(Key key, uint64_t m, T* __restrict__ out, bool exact) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = normal_at<T, LEGACY>(key, i, m, exact);
}

// ---- standalone Metropolis accept (samplers/sampler_utils.py:91-124) ------------------------------
template <typename T, bool LEGACY>
__global__ void k_accept(T* cur, const T* __restrict__ prop, T* lp_cur, const T* __restrict__ lp_prop,
                         const T* __restrict__ logacc, Key key, int n, int row0, int rows, int d,
                         int32_t* accepts) {
  // one warp per walker
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= rows) return;
  const T lo = T(1e-10);
  const T span = T(1.0) - lo;
  const T u = uniform_at<T, LEGACY>(key, (uint64_t)(row0 + w), (uint64_t)n, lo, span);
  const bool acc = log_t(u) < logacc[w];
  if (acc) {
    for (int c = lane; c < d; c += 32) cur[(size_t)w * d + c] = prop[(size_t)w * d + c];
  }
  if (lane == 0) {
    if (acc) lp_cur[w] = lp_prop[w];
    accepts[w] += acc ? 1 : 0;
  }
}

// ---- side-move complement choice (moves/hamiltonian/hmc_side.py:43-47) -----------------------------
// choice(key_i, arange(n), (2,), replace=False) = first two entries of jax.random._shuffle:
// `rounds` rounds of { key, sub = split(key); stable sort by random_bits(sub, 32, (n,)) }.
//   rounds == 1: the two positions with the smallest (sort key, index).
//   rounds == 2: the two smallest (key2, position) give ranks pa, pb; the answer is the element of
//                round-1 rank pa / pb, found by an MSB-first radix select over regenerated keys.
// One CTA per walker key.  keys come either from an explicit array or as split(parent, nsplit)[i].
template <bool LEGACY>
__device__ __forceinline__ void split2(Key k, Key& k_next, Key& sub) {
  k_next = split_key<LEGACY>(k, 0, 2);
  sub = split_key<LEGACY>(k, 1, 2);
}

struct MinPair { uint32_t k1, i1, k2, i2; };
__device__ __forceinline__ bool lex_less(uint32_t ka, uint32_t ia, uint32_t kb, uint32_t ib) {
  return (ka < kb) || (ka == kb && ia < ib);
}
__device__ __forceinline__ void minpair_push(MinPair& m, uint32_t k, uint32_t i) {
  if (lex_less(k, i, m.k1, m.i1)) { m.k2 = m.k1; m.i2 = m.i1; m.k1 = k; m.i1 = i; }
  else if (lex_less(k, i, m.k2, m.i2)) { m.k2 = k; m.i2 = i; }
}

template <bool LEGACY>
__global__ void __launch_bounds__(256) k_choice2(const uint32_t* __restrict__ keys, Key parent,
                                                 const uint32_t* parent_ptr, uint64_t nsplit, int64_t first,
                                                 int64_t nkeys, int n, int rounds, int32_t* __restrict__ out) {
  __shared__ MinPair wmin[8];
  __shared__ uint32_t hist[2][256];
  __shared__ uint32_t sel_prefix[2], sel_rank[2];
  __shared__ uint32_t ties[2][64];
  __shared__ uint32_t nties[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t w = blockIdx.x; w < nkeys; w += gridDim.x) {
    Key k;
    if (keys) { k.k0 = keys[2 * w]; k.k1 = keys[2 * w + 1]; }
    else {
      Key par = parent;
      if (parent_ptr) { par.k0 = parent_ptr[0]; par.k1 = parent_ptr[1]; }
      k = split_key<LEGACY>(par, (uint64_t)(first + w), nsplit);
    }
    if (rounds == 0) {  // n == 1: identity permutation (degenerate)
      if (tid == 0) { out[2 * w] = 0; out[2 * w + 1] = 0; }
      continue;
    }
    Key k1, sub1, k2, sub2;
    split2<LEGACY>(k, k1, sub1);
    Key sub_last = sub1;
    if (rounds == 2) { split2<LEGACY>(k1, k2, sub2); sub_last = sub2; }
    // two smallest (key, position) of the last round
    MinPair m{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    for (int j = tid; j < n; j += 256) minpair_push(m, bits32<LEGACY>(sub_last, (uint64_t)j, (uint64_t)n), (uint32_t)j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      MinPair t;
      t.k1 = __shfl_xor_sync(0xffffffffu, m.k1, o); t.i1 = __shfl_xor_sync(0xffffffffu, m.i1, o);
      t.k2 = __shfl_xor_sync(0xffffffffu, m.k2, o); t.i2 = __shfl_xor_sync(0xffffffffu, m.i2, o);
      minpair_push(m, t.k1, t.i1);
      minpair_push(m, t.k2, t.i2);
    }
    __syncthreads();
    if (lane == 0) wmin[warp] = m;
    __syncthreads();
    if (tid == 0) {
      MinPair a = wmin[0];
      for (int q = 1; q < 8; ++q) { minpair_push(a, wmin[q].k1, wmin[q].i1); minpair_push(a, wmin[q].k2, wmin[q].i2); }
      wmin[0] = a;
    }
    __syncthreads();
    const uint32_t pa = wmin[0].i1, pb = wmin[0].i2;
    if (rounds == 1) {
      if (tid == 0) { out[2 * w] = (int32_t)pa; out[2 * w + 1] = (int32_t)pb; }
      __syncthreads();
      continue;
    }
    // rounds == 2: select the elements of round-1 rank pa and pb (stable order: key, then index)
    if (tid < 2) { sel_prefix[tid] = 0; sel_rank[tid] = (tid == 0) ? pa : pb; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      hist[0][tid] = 0; hist[1][tid] = 0;
      __syncthreads();
      const uint32_t himask = (pass == 0) ? 0u : (0xFFFFFFFFu << (shift + 8));
      const uint32_t pf0 = sel_prefix[0], pf1 = sel_prefix[1];
      for (int j = tid; j < n; j += 256) {
        const uint32_t key1 = bits32<LEGACY>(sub1, (uint64_t)j, (uint64_t)n);
        const uint32_t b = (key1 >> shift) & 0xFFu;
        if ((key1 & himask) == pf0) atomicAdd(&hist[0][b], 1u);
        if ((key1 & himask) == pf1) atomicAdd(&hist[1][b], 1u);
      }
      __syncthreads();
      if (tid < 2) {
        uint32_t rank = sel_rank[tid], cum = 0;
        int b = 0;
        for (; b < 256; ++b) {
          const uint32_t c = hist[tid][b];
          if (rank < cum + c) break;
          cum += c;
        }
        sel_rank[tid] = rank - cum;
        sel_prefix[tid] |= ((uint32_t)b << shift);
      }
      __syncthreads();
    }
    // sel_prefix = exact key value; sel_rank = rank among equal keys in index order
    if (tid < 2) nties[tid] = 0;
    __syncthreads();
    {
      const uint32_t v0 = sel_prefix[0], v1 = sel_prefix[1];
      for (int j = tid; j < n; j += 256) {
        const uint32_t key1 = bits32<LEGACY>(sub1, (uint64_t)j, (uint64_t)n);
        if (key1 == v0) { const uint32_t s = atomicAdd(&nties[0], 1u); if (s < 64) ties[0][s] = (uint32_t)j; }
        if (key1 == v1) { const uint32_t s = atomicAdd(&nties[1], 1u); if (s < 64) ties[1][s] = (uint32_t)j; }
      }
    }
    __syncthreads();
    if (tid < 2) {
      // the sel_rank-th smallest index among the ties (tiny list; selection by counting)
      const uint32_t cnt = min(nties[tid], 64u), want = sel_rank[tid];
      uint32_t ans = 0;
      for (uint32_t a = 0; a < cnt; ++a) {
        uint32_t less = 0;
        for (uint32_t b = 0; b < cnt; ++b) less += (ties[tid][b] < ties[tid][a]) ? 1u : 0u;
        if (less == want) ans = ties[tid][a];
      }
      out[2 * w + tid] = (int32_t)ans;
    }
    __syncthreads();
  }
}

}  // namespace hx
