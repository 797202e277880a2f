// hx_spd.cuh — inverse of a symmetric positive definite matrix of a few hundred to ~1000 rows in fp64 (blocked Cholesky,
// blocked inverse of the factor, X^T X).  Used by the device-resident ChEES adapter beyond dim 128, whose criterion needs
// solve(cov(group2), p_proj^T) for every walker (reference adaptation/chees.py:57-73): with cov^-1 formed once per half-move
// the solve for all walkers becomes ONE tensor-core GEMM with a row-dot epilogue (hx_tc_walk.cu chees_inner), instead of the
// reference's LU solve with n right-hand sides.
//
// All kernels work on a row-major fp64 workspace A[d, ld]; 64 x 64 blocks, ~3 launches per block column (d = 1000: 49
// launches, ~0.5 ms -- the warm-up half-move around it takes 8 ms at that size).  A non-positive pivot yields NaNs, which
// the ChEES gradient mask then zeroes, like the reference's LU solve on a singular covariance.
#pragma once
#include <cuda_runtime.h>
#include "hx_ctx.h"
#include "hx_err.h"

namespace hx {
namespace spd {

constexpr int NB = 64;      // block size
constexpr int KC = 32;      // columns of a block staged in shared memory at a time

// A <- scale * M (lower triangle, fp32 -> fp64), zero above the diagonal
__global__ void k_spd_load(const float* __restrict__ M, int ldm, int d, double scale, double* __restrict__ A, int ld) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j < d) A[(size_t)i * ld + j] = (j <= i) ? scale * (double)M[(size_t)i * ldm + j] : 0.0;
}

// Block column k0: every CTA factors the diagonal block D = A[k0:k0+nb, k0:k0+nb] in shared memory (redundantly; CTA 0 writes
// it back), then solves its 64 rows below the block against it: A[i, k0:k0+nb] <- A[i, k0:k0+nb] L11^-T
__global__ void __launch_bounds__(NB) k_chol_panel(double* __restrict__ A, int ld, int d, int k0) {
  __shared__ double D[NB][NB + 1];
  const int nb = min(NB, d - k0), t = threadIdx.x;
  for (int r = 0; r < nb; ++r) D[r][t] = (t < nb && t <= r) ? A[(size_t)(k0 + r) * ld + k0 + t] : 0.0;
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    if (t == 0) D[k][k] = sqrt(D[k][k]);
    __syncthreads();
    if (t > k && t < nb) D[t][k] /= D[k][k];
    __syncthreads();
    if (t > k && t < nb)
      for (int j = k + 1; j <= t; ++j) D[t][j] -= D[t][k] * D[j][k];
    __syncthreads();
  }
  if (blockIdx.x == 0)
    for (int r = 0; r < nb; ++r)
      if (t < nb && t <= r) A[(size_t)(k0 + r) * ld + k0 + t] = D[r][t];
  const int i = k0 + nb + blockIdx.x * NB + t;
  if (i < d) {
    double x[NB];
#pragma unroll 8
    for (int c = 0; c < NB; ++c) x[c] = c < nb ? A[(size_t)i * ld + k0 + c] : 0.0;
    for (int c = 0; c < nb; ++c) {
      double s = x[c];
      for (int q = 0; q < c; ++q) s -= x[q] * D[c][q];
      x[c] = s / D[c][c];
    }
    for (int c = 0; c < nb; ++c) A[(size_t)i * ld + k0 + c] = x[c];
  }
}

// trailing update of the lower triangle: A[i][j] -= sum_c A[i][k0+c] A[j][k0+c] for i, j >= k1 = k0 + nb; 64 x 64 tiles,
// the panel staged 32 columns at a time (static shared memory stays below 48 KB)
__global__ void __launch_bounds__(256) k_chol_syrk(double* __restrict__ A, int ld, int d, int k0, int nb) {
  if (blockIdx.x > blockIdx.y) return;                       // tile column <= tile row
  __shared__ double Pi[NB][KC + 1], Pj[NB][KC + 1];
  const int k1 = k0 + nb, i0 = k1 + blockIdx.y * NB, j0 = k1 + blockIdx.x * NB;
  const int tr = (threadIdx.x >> 4) * 4, tc = (threadIdx.x & 15) * 4;
  double acc[4][4] = {};
  for (int c0 = 0; c0 < nb; c0 += KC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < NB * KC; idx += 256) {
      const int r = idx / KC, c = idx % KC;
      Pi[r][c] = (i0 + r < d && c0 + c < nb) ? A[(size_t)(i0 + r) * ld + k0 + c0 + c] : 0.0;
      Pj[r][c] = (j0 + r < d && c0 + c < nb) ? A[(size_t)(j0 + r) * ld + k0 + c0 + c] : 0.0;
    }
    __syncthreads();
    for (int c = 0; c < KC; ++c) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = Pi[tr + u][c]; b[u] = Pj[tc + u][c]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * b[v];
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + tr + u, j = j0 + tc + v;
      if (i < d && j <= i) A[(size_t)i * ld + j] -= acc[u][v];
    }
}

// X[b-th diagonal block] = inverse of the lower-triangular diagonal block of L; one CTA per block, thread = column
__global__ void __launch_bounds__(NB) k_trinv_diag(const double* __restrict__ Lm, int ld, int d, double* __restrict__ X) {
  __shared__ double D[NB][NB + 1];
  const int k0 = blockIdx.x * NB, nb = min(NB, d - k0), t = threadIdx.x;
  for (int r = 0; r < nb; ++r) D[r][t] = (t < nb && t <= r) ? Lm[(size_t)(k0 + r) * ld + k0 + t] : 0.0;
  __syncthreads();
  if (t < nb) {
    double x[NB];
    for (int r = 0; r < nb; ++r) {                           // column t of the inverse: L x = e_t
      double s = (r == t) ? 1.0 : 0.0;
      for (int q = t; q < r; ++q) s -= D[r][q] * x[q];
      x[r] = (r < t) ? 0.0 : s / D[r][r];
    }
    for (int r = 0; r < nb; ++r) X[(size_t)(k0 + r) * ld + k0 + t] = x[r];
  }
}

// block row ib of X = L^-1, block column jb = blockIdx.x < ib:  X_ij = -X_ii (sum_{k = j}^{i-1} L_ik X_kj)
__global__ void __launch_bounds__(256) k_trinv_row(const double* __restrict__ Lm, int ld, int d, int ib, double* __restrict__ X) {
  __shared__ double Sa[NB][KC + 1], Sb[KC][NB + 1];
  const int jb = blockIdx.x, i0 = ib * NB, j0 = jb * NB;
  const int tr = (threadIdx.x >> 4) * 4, tc = (threadIdx.x & 15) * 4;
  double acc[4][4] = {};
  for (int k0 = j0; k0 < i0; k0 += KC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < NB * KC; idx += 256) {
      const int r = idx / KC, c = idx % KC;
      Sa[r][c] = (i0 + r < d) ? Lm[(size_t)(i0 + r) * ld + k0 + c] : 0.0;          // L_ik   (k0 + c < i0 <= d)
      const int r2 = idx / NB, c2 = idx % NB;
      Sb[r2][c2] = X[(size_t)(k0 + r2) * ld + j0 + c2];                            // X_kj   (rows < i0: complete)
    }
    __syncthreads();
    for (int c = 0; c < KC; ++c) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = Sa[tr + u][c]; b[u] = Sb[c][tc + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * b[v];
    }
  }
  // X_ij = -X_ii W with W = the sum above (held in registers): 32 rows of W / 32 columns of X_ii at a time
  double o[4][4] = {};
  for (int c0 = 0; c0 < NB; c0 += KC) {
    __syncthreads();
    if (tr >= c0 && tr < c0 + KC) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) Sb[tr - c0 + u][tc + v] = acc[u][v];
    }
    for (int idx = threadIdx.x; idx < NB * KC; idx += 256) {
      const int r = idx / KC, c = idx % KC;
      Sa[r][c] = (i0 + r < d && i0 + c0 + c < d) ? X[(size_t)(i0 + r) * ld + i0 + c0 + c] : 0.0;
    }
    __syncthreads();
    for (int c = 0; c < KC; ++c) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = Sa[tr + u][c]; b[u] = Sb[c][tc + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) o[u][v] += a[u] * b[v];
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (i0 + tr + u < d) X[(size_t)(i0 + tr + u) * ld + j0 + tc + v] = -o[u][v];
}

// out[a][b] = sum_r X[r][a] X[r][b]  (X lower triangular: r >= max(a, b)), fp32 result [d, ldo]; 64 x 64 tiles
__global__ void __launch_bounds__(256) k_xtx(const double* __restrict__ X, int ld, int d, float* __restrict__ out, int ldo) {
  __shared__ double Sa[KC][NB + 1], Sb[KC][NB + 1];
  const int a0 = blockIdx.y * NB, b0 = blockIdx.x * NB;
  const int tr = (threadIdx.x >> 4) * 4, tc = (threadIdx.x & 15) * 4;
  double acc[4][4] = {};
  for (int r0 = max(a0, b0); r0 < d; r0 += KC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < KC * NB; idx += 256) {
      const int r = idx / NB, c = idx % NB;
      Sa[r][c] = (r0 + r < d && a0 + c < d) ? X[(size_t)(r0 + r) * ld + a0 + c] : 0.0;
      Sb[r][c] = (r0 + r < d && b0 + c < d) ? X[(size_t)(r0 + r) * ld + b0 + c] : 0.0;
    }
    __syncthreads();
    for (int r = 0; r < KC; ++r) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = Sa[r][tr + u]; b[u] = Sb[r][tc + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * b[v];
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (a0 + tr + u < d && b0 + tc + v < d) out[(size_t)(a0 + tr + u) * ldo + b0 + tc + v] = (float)acc[u][v];
}

// inv_out[d, ldo] (fp32) = (scale * M)^-1 for the symmetric positive definite M[d, ldm] (fp32; its lower triangle is read).
// work: 2 * d * ld doubles with ld = d rounded up to a multiple of NB.
inline int spd_inverse(hx_ctx* ctx, const float* M, int ldm, int d, double scale, double* work, float* inv_out, int ldo, cudaStream_t s) {
  const int nblk = (d + NB - 1) / NB, ld = nblk * NB;
  double* A = work;
  double* X = work + (size_t)ld * ld;
  HX_CUDA(cudaMemsetAsync(work, 0, (size_t)2 * ld * ld * sizeof(double), s));
  k_spd_load<<<dim3((d + 127) / 128, d), 128, 0, s>>>(M, ldm, d, scale, A, ld);
  for (int b = 0; b < nblk; ++b) {
    const int k0 = b * NB, nb = d - k0 < NB ? d - k0 : NB, below = d - k0 - nb;
    k_chol_panel<<<below > 0 ? (below + NB - 1) / NB : 1, NB, 0, s>>>(A, ld, d, k0);
    if (below > 0) {
      const int tiles = (below + NB - 1) / NB;
      k_chol_syrk<<<dim3(tiles, tiles), 256, 0, s>>>(A, ld, d, k0, nb);
    }
    ctx->launches += below > 0 ? 2 : 1;
  }
  k_trinv_diag<<<nblk, NB, 0, s>>>(A, ld, d, X);
  for (int ib = 1; ib < nblk; ++ib) k_trinv_row<<<ib, 256, 0, s>>>(A, ld, d, ib, X);
  k_xtx<<<dim3(nblk, nblk), 256, 0, s>>>(X, ld, d, inv_out, ldo);
  ctx->launches += 2 + nblk;
  HX_LAUNCH_CHECK("spd_inverse");
  return 0;
}

}  // namespace spd
}  // namespace hx
