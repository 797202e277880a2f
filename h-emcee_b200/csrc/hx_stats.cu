// hx_stats.cu — on-device chain statistics: running moments over all walkers and the lagged-product
// accumulators from which the reference's integrated autocorrelation time is recovered WITHOUT storing the chain.
//
// The reference keeps every iteration on the host (backend/backend.py:174-211, 262 MB per iteration at c5) and estimates
// tau from the stored chain (autocorr.py:15-34 function_1d, :44-118 integrated_time): per walker and parameter the series
// mean is removed, the autocovariance sum_t (x_t - m)(x_{t+k} - m) is formed by FFT, normalised by lag 0, averaged over
// walkers, and summed up to the first window M >= c tau(M).  For lags k < K that autocovariance is an exact function of
//     A_k   = sum_{t >= k} x_t x_{t-k}            (lagsum)
//     total = sum_t x_t,  the first K values and the last K values of the series
// because  sum_{t<T-k} (x_t - m)(x_{t+k} - m) = A_k - m (head_k + tail_k) + (T - k) m^2  with head_k = total - (last k values),
// tail_k = total - (first k values).  One feed per iteration updates them in O(K) per tracked series; tau(M) for M < K is then
// the reference's number up to f64 rounding (hemcee_b200/stats.py finishes the estimate on the device with torch).
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/hx.h"
#include "hx_ctx.h"
#include "hx_err.h"

namespace hx {

constexpr int kStatsRowBlocks = 64;     // row blocks of the column-sum partials (fixed: the sums are bit-reproducible)

// partial[0][rb][c] = sum over rows rb, rb + RB, ... of X[r][c];  partial[1][rb][c] = same for X^2
template <typename T>
__global__ void __launch_bounds__(256) k_stats_colsum(const T* __restrict__ X, int W, int d, double* __restrict__ partial) {
  __shared__ double s1[8][33], s2[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  double a = 0.0, b = 0.0;
  if (c < d)
    for (int r = blockIdx.y * 8 + rl; r < W; r += 8 * gridDim.y) {
      const double x = (double)X[(size_t)r * d + c];
      a += x; b += x * x;
    }
  s1[rl][threadIdx.x & 31] = a; s2[rl][threadIdx.x & 31] = b;
  __syncthreads();
  if (rl == 0 && c < d) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { a += s1[k][threadIdx.x]; b += s2[k][threadIdx.x]; }
    partial[(size_t)blockIdx.y * d + c] = a;
    partial[((size_t)gridDim.y + blockIdx.y) * d + c] = b;
  }
}

// adds the partials in fixed order; thread 0 of block 0 counts the iteration (LAST kernel of a feed: the others read the old count)
__global__ void k_stats_moments_final(const double* __restrict__ partial, int d, int RB, double* __restrict__ sum,
                                      double* __restrict__ sumsq, long long* __restrict__ count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < d) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < RB; ++k) { a += partial[(size_t)k * d + c]; b += partial[((size_t)RB + k) * d + c]; }
    sum[c] += a; sumsq[c] += b;
  }
  if (c == 0) *count += 1;
}

// series s = tracked walker j (row j * stride) x tracked dim index i (s = j * nd + i); blockIdx.y = chunk of 8 lags
template <typename T>
__global__ void __launch_bounds__(256) k_stats_lags(const T* __restrict__ X, int d, int stride, int nd, const int32_t* __restrict__ dims,
                                                    int S, int K, const long long* __restrict__ count, double* __restrict__ lagsum,
                                                    T* __restrict__ ring, T* __restrict__ first, double* __restrict__ total) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const long long t = *count;
  const int j = s / nd, i = s - j * nd;
  const T x = X[(size_t)j * stride * d + (dims ? dims[i] : i)];
  const int k0 = blockIdx.y * 8;
  const int slot = (int)(t % K);
  if (k0 == 0) {
    total[s] += (double)x;
    if (t < K) first[(size_t)t * S + s] = x;
    lagsum[s] += (double)x * (double)x;
  }
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const int k = k0 + kk;
    if (k >= 1 && k < K && k <= t) {
      int sl = slot - k; if (sl < 0) sl += K;
      lagsum[(size_t)k * S + s] += (double)x * (double)ring[(size_t)sl * S + s];
    }
  }
  // slot t % K holds x_{t-K}, which no lag < K reads: no ordering needed against the other lag chunks
  if (k0 == 0) ring[(size_t)slot * S + s] = x;
}

static int check_stats(const hx_chain_stats* st, const char* who) {
  if (!st) return hx_fail(HX_E_NULL, "%s: NULL stats", who);
  if (st->nwalkers < 1 || st->dim < 1) return hx_fail(HX_E_SHAPE, "%s: nwalkers and dim must be >= 1", who);
  if (!st->count_dev || !st->sum_dev || !st->sumsq_dev || !st->partial_dev) return hx_fail(HX_E_NULL, "%s: NULL moment buffer", who);
  if (st->n_track < 0 || st->n_dims < 0 || st->max_lag < 0) return hx_fail(HX_E_SHAPE, "%s: negative size", who);
  if (st->n_track > 0 && st->n_dims > 0 && st->max_lag > 0) {
    if (st->walker_stride < 1 || (int64_t)(st->n_track - 1) * st->walker_stride >= st->nwalkers)
      return hx_fail(HX_E_SHAPE, "%s: tracked walkers 0, %d, ... (%d of them) exceed nwalkers %d", who, st->walker_stride, st->n_track, st->nwalkers);
    if (st->n_dims > st->dim) return hx_fail(HX_E_SHAPE, "%s: n_dims %d > dim %d", who, st->n_dims, st->dim);
    if (!st->lagsum_dev || !st->ring_dev || !st->first_dev || !st->total_dev) return hx_fail(HX_E_NULL, "%s: NULL lag buffer", who);
    if ((int64_t)st->n_track * st->n_dims > 0x7FFFFFFFll) return hx_fail(HX_E_SHAPE, "%s: too many tracked series", who);
  }
  return 0;
}

template <typename T>
static int stats_feed_t(hx_ctx* ctx, const hx_chain_stats* st, const T* X, cudaStream_t s) {
  const int S = st->n_track * st->n_dims;
  if (S > 0 && st->max_lag > 0) {
    dim3 g((S + 255) / 256, (st->max_lag + 7) / 8);
    k_stats_lags<T><<<g, 256, 0, s>>>(X, st->dim, st->walker_stride, st->n_dims, st->dims_dev, S, st->max_lag,
                                      (const long long*)st->count_dev, st->lagsum_dev, (T*)st->ring_dev, (T*)st->first_dev, st->total_dev);
    ctx->launches += 1;
  }
  dim3 g((st->dim + 31) / 32, kStatsRowBlocks);
  k_stats_colsum<T><<<g, 256, 0, s>>>(X, st->nwalkers, st->dim, st->partial_dev);
  k_stats_moments_final<<<(st->dim + 255) / 256, 256, 0, s>>>(st->partial_dev, st->dim, kStatsRowBlocks, st->sum_dev, st->sumsq_dev,
                                                             (long long*)st->count_dev);
  ctx->launches += 2;
  HX_LAUNCH_CHECK("k_stats");
  return 0;
}

int stats_feed(hx_ctx* ctx, int dtype, const hx_chain_stats* st, const void* X, cudaStream_t s) {
  if (dtype == HX_F32) return stats_feed_t<float>(ctx, st, (const float*)X, s);
  return stats_feed_t<double>(ctx, st, (const double*)X, s);
}

}  // namespace hx

using namespace hx;

extern "C" HX_API int32_t hx_stats_row_blocks(void) { return kStatsRowBlocks; }

extern "C" HX_API int hx_stats_feed(hx_ctx* ctx, int dtype, const hx_chain_stats* stats, const void* X_dev, void* stream) {
  if (!ctx || !X_dev) return hx_fail(HX_E_NULL, "hx_stats_feed: NULL argument");
  if (dtype != HX_F32 && dtype != HX_F64) return hx_fail(HX_E_DTYPE, "hx_stats_feed: unknown dtype %d", dtype);
  if (int e = check_stats(stats, "hx_stats_feed")) return e;
  HX_CUDA(cudaSetDevice(ctx->device));
  return stats_feed(ctx, dtype, stats, X_dev, (cudaStream_t)stream);
}

extern "C" HX_API int hx_ctx_set_chain_stats(hx_ctx* ctx, const hx_chain_stats* stats) {
  if (!ctx) return hx_fail(HX_E_NULL, "hx_ctx_set_chain_stats: NULL ctx");
  if (!stats) { ctx->stats_on = 0; return 0; }
  if (int e = check_stats(stats, "hx_ctx_set_chain_stats")) return e;
  ctx->stats = *stats;
  ctx->stats_on = 1;
  return 0;
}
