// hx_tc_walk.cu — tensor-core path of the walk half-move for full-covariance Gaussian targets at large
// dim / walker counts (fp32 state, split 16-bit / 8-bit operands on tcgen05, fp32 accumulation in TMEM).
//
// Same projected-form algorithm as the exact SIMT kernel (hx_move.cuh; reference
// moves/hamiltonian/hmc_walk.py:6-107), expressed as GEMMs, with the two per-kick products folded into
// one ("B-form"): for a Gaussian target grad U = (q - mu) P, so  g @ M = (q - mu) @ (P M).
//   C^T planes    <- ((X2 - mean)/sqrt(n))^T                      (hmc_walk.py:35)   bf16 hi/lo (Gram) and fp16 + 2 x e4m3 (projection)
//   M             <- C^T C                                         k_tc_gemm_pair(CT, CT), split-K          bf16 x3
//   Bop           <- M P  (fp64 tensor cores, d^3 flops)           k_mp_dmma
//   P0 planes     <- normal(key, (n, n))[row shard]                (hmc_walk.py:36)   k_rng_planes, one half-move ahead on a side stream
//   S0            <- P0 C                                          k_tc_gemm(P0, CT)  fp16 main + 2 x e4m3 correction products
//   stepwise form, per kick:  S <- S - a (q - mu) Bop^T ; Qs += a (q - mu) ; q += eps S      k_tc_gemm<EPI_KICKB>   fp16 hi/lo
//   propagator form: 2 dim basis trajectories with the same kicks (bf16 x3), then [q_L - mu | S_L] = z0 [Tq | Ts] in ONE GEMM
//   energies:  dK = -1/2 <Qs P, S0 + S_L> ; lp' = -1/2 <(q_L - mu) P, q_L - mu>            k_tc_gemm<EPI_DOT>     fp16 hi/lo
//   finish:    dH = (lp1 - lp') + dK ; log_acc ; optional Metropolis accept + chain emit (+ row push to the peers)
// Operand kinds and what each costs in accuracy are in hx_tc_gemm.cuh and DESIGN.md sections 2 and 4: with bf16 hi/lo (2^-16
// per product) the energy error of an ill-conditioned target (kappa 1e4) is ~1e-2 and visibly lowers the acceptance rate, so
// the kicks of the stepwise form and all energy products use fp16 hi/lo operands with power-of-two scales (2^-22).
#include <cstring>
#include <cstdio>
#include <cstdlib>

#include "../../include/hx.h"
#include "hx_ctx.h"
#include "hx_rng.cuh"
#include "hx_common.cuh"
#include "hx_tc_gemm.cuh"
#include "hx_tc_fused.cuh"
#include "hx_tc_walk.h"
#include <cuda_fp16.h>
#include <cuda_fp8.h>

namespace hx {
namespace tc {

constexpr int BN = 256;

// ---- tensor maps -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode(hx_ctx* ctx, EncodeFn* fn) {
  if (!ctx->encode_fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    HX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess) return hx_fail(HX_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
    ctx->encode_fn = p;
  }
  *fn = (EncodeFn)ctx->encode_fn;
  return 0;
}

// planes array [rows_total, Kp], K-major; box = [one 128-byte swizzle row of K, box_rows], 128B swizzle
static int make_tmap(hx_ctx* ctx, CUtensorMap* out, const void* base, uint64_t rows_total, uint64_t Kp, uint32_t box_rows,
                     int kind) {
  EncodeFn enc;
  if (int e = get_encode(ctx, &enc)) return e;
  const uint64_t esz = kind == KIND_TF32 ? 4 : 2;
  cuuint64_t dims[2] = {Kp, rows_total};
  cuuint64_t strides[1] = {Kp * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = kind == KIND_TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                               : kind == KIND_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
  CUresult r = enc(out, dt, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return hx_fail(HX_E_UNSUPPORTED, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// ---- fp16 + e4m3 operand form (KIND_F16F8, hx_tc_gemm.cuh) ------------------------------------------------------------------
// x ~= fp16(x) + residual; the tensor core gets fp16(x) for the main product and e4m3(x), e4m3(residual * 2^11) for the two
// correction products.  8 consecutive K elements at a time: 16 bytes of the fp16 plane, 8 + 8 bytes of the e4m3 plane whose
// 128-byte rows hold [64 values | 64 residuals] per 64-element K-block.
// fmt8 = 2 (B operand of the fused projection, hx_tc_fused.cuh FMT 1): residuals at TRUE scale in e5m2 and the halves swapped,
// rows = [64 x e5m2(residual) | 64 x e4m3(value)], so that they pair with the A rows [e4m3(value) | e5m2(residual)].
__device__ __forceinline__ void store_f16f8_x8(uint16_t* __restrict__ plane16, uint8_t* __restrict__ plane8_row, size_t o16, int j0,
                                               const float (&x)[8], int fmt8) {
  __align__(16) __half h[8];
  __align__(8) __nv_fp8x2_storage_t v8[4], r8[4];
#pragma unroll
  for (int t = 0; t < 8; ++t) h[t] = __float2half_rn(x[t]);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float rs = fmt8 == 2 ? 1.0f : kF8ResidualScale;
    const float r0 = (x[2 * t] - __half2float(h[2 * t])) * rs;
    const float r1 = (x[2 * t + 1] - __half2float(h[2 * t + 1])) * rs;
    v8[t] = __nv_cvt_float2_to_fp8x2(make_float2(x[2 * t], x[2 * t + 1]), __NV_SATFINITE, __NV_E4M3);
    r8[t] = fmt8 == 2 ? __nv_cvt_float2_to_fp8x2(make_float2(r0, r1), __NV_SATFINITE, __NV_E5M2)
                      : __nv_cvt_float2_to_fp8x2(make_float2(r0, r1), __NV_SATFINITE, __NV_E4M3);
  }
  *reinterpret_cast<uint4*>(plane16 + o16) = *reinterpret_cast<const uint4*>(h);
  uint8_t* p8 = plane8_row + (size_t)(j0 >> 6) * 128 + (j0 & 63);
  *reinterpret_cast<uint2*>(p8 + (fmt8 == 2 ? 64 : 0)) = *reinterpret_cast<const uint2*>(v8);
  *reinterpret_cast<uint2*>(p8 + (fmt8 == 2 ? 0 : 64)) = *reinterpret_cast<const uint2*>(r8);
}

// power-of-two scale of the centred complement C = (X2 - mean)/sqrt(n): max |C| * sc lies in [32, 64), inside the fp16 and
// e4m3 ranges with room for the accumulating sums.  amax_bits = bits of max |X2 - mean| (k_absmax_centred).
__device__ __forceinline__ float f8_scale_from_amax(uint32_t amax_bits, int n) {
  const float amax = __uint_as_float(amax_bits) / sqrtf((float)n);
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);                       // amax = m * 2^e, m in [0.5, 1)
  return ldexpf(1.f, min(6 - e, 100));    // amax * sc = m * 64 in [32, 64); (nearly) all-zero complement: finite scale
}
__global__ void __launch_bounds__(256) k_absmax_centred(const float* __restrict__ X2, const float* __restrict__ mean, int n, int d,
                                                        uint32_t* __restrict__ amax_bits) {
  // threads over columns (coalesced rows, mean[c] read once), blockIdx.y over chunks of 64 rows; 4 independent loads in flight
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r_lo = blockIdx.y * 64, r_hi = min(n, r_lo + 64);
  float m = 0.f;
  if (c < d) {
    const float mc = mean ? mean[c] : 0.f;
    int r = r_lo;
    for (; r + 4 <= r_hi; r += 4) {
      const float a0 = X2[(size_t)r * d + c], a1 = X2[(size_t)(r + 1) * d + c];
      const float a2 = X2[(size_t)(r + 2) * d + c], a3 = X2[(size_t)(r + 3) * d + c];
      m = fmaxf(fmaxf(m, fmaxf(fabsf(a0 - mc), fabsf(a1 - mc))), fmaxf(fabsf(a2 - mc), fabsf(a3 - mc)));
    }
    for (; r < r_hi; ++r) m = fmaxf(m, fabsf(X2[(size_t)r * d + c] - mc));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, wm[w]);
    if (m > 0.f) atomicMax(amax_bits, __float_as_uint(m));       // non-negative floats order like their bit patterns
  }
}
// max |src[r * ld + c]| over r < rows, c < cols
__global__ void __launch_bounds__(256) k_absmax_ld(const float* __restrict__ src, int rows, int cols, int ld, uint32_t* __restrict__ amax_bits) {
  float m = 0.f;
  const size_t tot = (size_t)rows * cols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (size_t)cols), c = (int)(i - (size_t)r * cols);
    m = fmaxf(m, fabsf(src[(size_t)r * ld + c]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}
// scale_out = {sc, 1/sc}
__global__ void k_f8_scale(const uint32_t* __restrict__ amax_bits, int n, float* __restrict__ scale_out) {
  const float sc = f8_scale_from_amax(*amax_bits, n);
  scale_out[0] = sc;
  scale_out[1] = 1.0f / sc;
}

// ---- elementwise / layout kernels -----------------------------------------------------------------------------
// CT planes [2, dB_pad, Kp]: CT[c][j] = (X2[j][c] - mean[c]) / sqrt(n); zero outside (c < d, j < n)
// planes32 (nullable): the same matrix as tf32 hi/lo planes [2, dB_pad, Kp] for the Gram product
// One CTA transposes a 256 (walkers j) x 32 (columns c) block through shared memory: coalesced 128-byte reads of X2 rows, and
// 16-byte stores (8 consecutive j) of the K-major planes.
__global__ void __launch_bounds__(256) k_center_planes_T(const float* __restrict__ X2, const float* __restrict__ mean, int n, int d,
                                                         int dB_pad, int Kp, __nv_bfloat16* __restrict__ planes,
                                                         float* __restrict__ planes32, uint16_t* __restrict__ planes8,
                                                         const float* __restrict__ scale8, int fmt8) {
  constexpr int TJ = 256, TC = 32;
  __shared__ float tile[TJ][TC + 1];
  const int j0 = blockIdx.x * TJ, c0 = blockIdx.y * TC;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float sc = sqrtf((float)n);
  const int c = c0 + tx;
  const float m = (c < d) ? mean[c] : 0.f;
  // 8 independent row loads in flight per thread (the rows are 4 d bytes apart: every load is a separate DRAM burst)
#pragma unroll 1
  for (int g0 = 0; g0 < TJ; g0 += 64) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + g0 + ty + 8 * u;
      v[u] = (j < n && c < d) ? X2[(size_t)j * d + c] : m;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) tile[g0 + ty + 8 * u][tx] = (v[u] - m) / sc;
  }
  __syncthreads();
  const size_t plane = (size_t)dB_pad * Kp;
  // lane = 4 consecutive 8-walker chunks x 8 columns: the shared-memory reads (row pitch 33 words, rows 8 apart per chunk)
  // hit 32 distinct banks, and the four lanes of one column store 64 contiguous bytes
  const int jcl = threadIdx.x & 3, ccl = (threadIdx.x >> 2) & 7, wq = threadIdx.x >> 5;
  for (int it = 0; it < TC / 8; ++it) {
    const int cc = it * 8 + ccl, jc = wq * 4 + jcl;
    const int crow = c0 + cc, jb = j0 + jc * 8;
    if (crow >= dB_pad || jb >= Kp) continue;                       // Kp is a multiple of 64: chunks are all-in or all-out
    __align__(16) __nv_bfloat16 hi[8], lo[8];
    float h32[8], l32[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float v = tile[jc * 8 + t][cc];
      split_bf16(v, hi[t], lo[t]);
      if (planes32) split_tf32(v, h32[t], l32[t]);
    }
    const size_t o = (size_t)crow * Kp + jb;
    if (planes8) {      // scaled copy in fp16 + e4m3 form for the momentum projection (KIND_F16F8)
      const float sc = scale8[0];
      float xs[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) xs[t] = tile[jc * 8 + t][cc] * sc;
      store_f16f8_x8(planes8, reinterpret_cast<uint8_t*>(planes8 + plane + (size_t)crow * Kp), o, jb, xs, fmt8);
    }
    *reinterpret_cast<uint4*>(planes + o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(planes + plane + o) = *reinterpret_cast<const uint4*>(lo);
    if (planes32) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        reinterpret_cast<float4*>(planes32 + o)[q] = make_float4(h32[4 * q], h32[4 * q + 1], h32[4 * q + 2], h32[4 * q + 3]);
        reinterpret_cast<float4*>(planes32 + plane + o)[q] = make_float4(l32[4 * q], l32[4 * q + 1], l32[4 * q + 2], l32[4 * q + 3]);
      }
    }
  }
}

// planes [2, rows_pad, Kp] <- split(src[r*ld + k]) for r < dim_r, k < dim_k, zero elsewhere
template <int KIND>
__global__ void k_split_planes(const float* __restrict__ src, int dim_r, int dim_k, int ld, int rows_pad, int Kp,
                               typename Kind<KIND>::elem* __restrict__ planes, const uint32_t* __restrict__ amax = nullptr,
                               const float* __restrict__ row_scale = nullptr) {
  const size_t tot = (size_t)rows_pad * Kp, plane = tot;
  const float sc16 = (KIND == KIND_F16 && amax) ? f16_scale(*amax, kF16TopB) : 1.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / Kp), k = (int)(i - (size_t)r * Kp);
    const float v = (r < dim_r && k < dim_k) ? src[(size_t)r * ld + k] : 0.f;
    if (KIND == KIND_F16) {
      __half hi, lo;
      split_f16(v * ((row_scale && r < dim_r) ? row_scale[r] : sc16), hi, lo);
      reinterpret_cast<__half*>(planes)[i] = hi;
      reinterpret_cast<__half*>(planes)[plane + i] = lo;
    } else if (KIND == KIND_BF16) {
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      reinterpret_cast<__nv_bfloat16*>(planes)[i] = hi;
      reinterpret_cast<__nv_bfloat16*>(planes)[plane + i] = lo;
    } else {
      float hi, lo;
      split_tf32(v, hi, lo);
      reinterpret_cast<float*>(planes)[i] = hi;
      reinterpret_cast<float*>(planes)[plane + i] = lo;
    }
  }
}

// B = M P: the product of the Gram matrix (eigenvalues up to ~1e3 at kappa = 1e4) and the precision (up to ~10) is O(1) by
// cancellation, and accumulating it in fp32 costs ~1e-4 of the ensemble scale in the proposed positions (measured against an f64
// integration; oracle-side check in tests/test_oracle_moves.py), so it is formed with fp64 products and accumulation.
// B = M P on the fp64 tensor-core path: DMMA m8n8k4 (fp64 operands and accumulators), fp32 inputs widened once when a
// K-block is staged in shared memory, register-prefetched double buffering.  64 x 64 output tile per CTA, four warps of
// 32 x 32 (4 x 4 DMMA tiles).  Row strides of the staged tiles are = 4 mod 16 doubles: the 8-byte fragment loads of a
// half-warp (4 rows x 4 k) then cover the 32 banks exactly once.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(128) k_mp_dmma(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                 int dim, float* __restrict__ out, int ldo) {
  constexpr int TB = 64, KB = 16, LDA = KB + 4, LDB = TB + 4;
  __shared__ double sa[2][TB * LDA];      // [m][k]
  __shared__ double sb[2][KB * LDB];      // [k][n]
  const int ti = blockIdx.y * TB, tj = blockIdx.x * TB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  float ra[8], rb[8];
  auto gload = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int idx = threadIdx.x + r * 128;
      const int m = idx >> 4, k = idx & 15;
      ra[r] = (ti + m < dim && k0 + k < dim) ? A[(size_t)(ti + m) * lda + k0 + k] : 0.f;
      const int k2 = idx >> 6, n = idx & 63;
      rb[r] = (k0 + k2 < dim && tj + n < dim) ? B[(size_t)(k0 + k2) * ldb + tj + n] : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int idx = threadIdx.x + r * 128;
      sa[buf][(idx >> 4) * LDA + (idx & 15)] = (double)ra[r];
      sb[buf][(idx >> 6) * LDB + (idx & 63)] = (double)rb[r];
    }
  };
  const int nkb = (dim + KB - 1) / KB;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkb) gload((kb + 1) * KB);
#pragma unroll
    for (int kk = 0; kk < KB; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sa[buf][(wm + 8 * i + g) * LDA + kk + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sb[buf][(kk + t) * LDB + wn + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (kb + 1 < nkb) sstore(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = ti + wm + 8 * i + g;
    if (gi >= dim) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = tj + wn + 8 * j + 2 * t;
      if (gj < dim) out[(size_t)gi * ldo + gj] = (float)acc[i][j][0];
      if (gj + 1 < dim) out[(size_t)gi * ldo + gj + 1] = (float)acc[i][j][1];
    }
  }
}

// Rows [0, rows_chunk) of the P0 planes (hi at `planes`, lo at `planes + plane`): P0[r][j] = normal(key, (n, n))[row0 + r, j],
// zero for r >= rows or j >= n; 8 consecutive j per thread
template <bool LEGACY, int FMT>      // FMT 0: bf16 hi/lo planes, 1: fp16 + e4m3 planes (KIND_F16F8)
__global__ void __launch_bounds__(256) k_rng_planes(Key key, const uint32_t* key_ptr, int n, int row0, int rows,
                                                    int rows_chunk, int Kp, size_t plane, __nv_bfloat16* __restrict__ planes,
                                                    unsigned int* __restrict__ work_ctr) {
  if (key_ptr) { key.k0 = key_ptr[0]; key.k1 = key_ptr[1]; }
  const size_t nvec = (size_t)rows_chunk * Kp / 8;
  const uint64_t nn = (uint64_t)n * (uint64_t)n;
  // Work items of 256 x 2 vectors are handed out through an atomic counter instead of a static grid-stride split: the kernel
  // shares the SMs with GEMM CTAs of another stream, so its resident CTAs progress at very different rates (and may be placed
  // unevenly), and a static split makes the whole draw wait for the slowest one (measured: a bimodal 18.3 / 22 ms per iteration)
  __shared__ unsigned int s_item;
  const unsigned int nitems = (unsigned int)((nvec + 511) / 512);      // short items: the kernel's tail is at most one item long
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(work_ctr, 1u);
    __syncthreads();
    const unsigned int item = s_item;
    __syncthreads();
    if (item >= nitems) break;
#pragma unroll 1
  for (int u4 = 0; u4 < 2; ++u4) {
    const size_t v = (size_t)item * 512 + (size_t)u4 * 256 + threadIdx.x;
    if (v >= nvec) break;
    const size_t i = v * 8;
    const int r = (int)(i / Kp), j0 = (int)(i - (size_t)r * Kp);
    float x[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int j = j0 + t;
      x[t] = 0.f;
      if (r < rows && j < n) x[t] = normal_at<float, LEGACY>(key, (uint64_t)(row0 + r) * (uint64_t)n + (uint64_t)j, nn);
    }
    if (FMT == 2) {        // diagnostics (hx_ctx_set_tuning(10, 1)): the draw's arithmetic without its stores
      float t = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) t += x[u];
      if (t == 12345.678f) planes[i] = __float2bfloat16_rn(t);
    } else if (FMT == 1) {
      uint16_t* p16 = reinterpret_cast<uint16_t*>(planes);
      store_f16f8_x8(p16, reinterpret_cast<uint8_t*>(p16 + plane + (size_t)r * Kp), i, j0, x, 1);
    } else {
      __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) split_bf16(x[t], hi[t], lo[t]);
      *reinterpret_cast<uint4*>(planes + i) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(planes + plane + i) = *reinterpret_cast<const uint4*>(lo);
    }
  }
  }
}

// Q <- X1 ; QC planes (hi/lo of the leapfrog operand kind) <- split(X1 - mu), zero padded
template <int KIND>
__global__ void k_tc_init(const float* __restrict__ X1, const float* __restrict__ mu, int rows, int d, int rows_pad, int Kp,
                          float* __restrict__ Q, typename Kind<KIND>::elem* __restrict__ planes) {
  const size_t tot = (size_t)rows_pad * Kp, plane = tot;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / Kp), k = (int)(i - (size_t)r * Kp);
    float v = 0.f;
    if (r < rows && k < d) {
      const float x = X1[(size_t)r * d + k];
      Q[(size_t)r * d + k] = x;
      v = x - (mu ? mu[k] : 0.f);
    }
    if (KIND == KIND_BF16) {
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      reinterpret_cast<__nv_bfloat16*>(planes)[i] = hi;
      reinterpret_cast<__nv_bfloat16*>(planes)[plane + i] = lo;
    } else {
      float hi, lo;
      split_tf32(v, hi, lo);
      reinterpret_cast<float*>(planes)[i] = hi;
      reinterpret_cast<float*>(planes)[plane + i] = lo;
    }
  }
}

// basis of the propagator run: bX = [I; 0], bS = [0; I]  (each [2d, d])
__global__ void k_basis_init(float* __restrict__ bX, float* __restrict__ bS, int d) {
  const size_t tot = (size_t)2 * d * d;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / d), c = (int)(i - (size_t)r * d);
    bX[i] = (r == c) ? 1.f : 0.f;
    bS[i] = (r == c + d) ? 1.f : 0.f;
  }
}

// z0 = [X1 - mu | S0] as A-operand planes with K2 = 2*dK columns (x part at [0, dK), S part at [dK, 2 dK)), zero padded:
// zB = bf16 hi/lo [2, rA, K2] (nullable), zT = tf32 hi/lo [2, rA, K2] (nullable)
__global__ void k_z0_planes(const float* __restrict__ X1, const float* __restrict__ mu, const float* __restrict__ S0, int rows,
                            int d, int rA, int dK, __nv_bfloat16* __restrict__ zB, float* __restrict__ zT) {
  const int K2 = 2 * dK;
  const size_t tot4 = (size_t)rA * K2 / 4, plane = (size_t)rA * K2;
  for (size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < tot4; i4 += (size_t)gridDim.x * blockDim.x) {
    const size_t i = i4 * 4;
    const int r = (int)(i / K2), kk = (int)(i - (size_t)r * K2);
    const int part = kk >= dK, c0 = kk - part * dK;
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = c0 + t;
      float x = 0.f;
      if (r < rows && c < d) x = part ? S0[(size_t)r * d + c] : X1[(size_t)r * d + c] - (mu ? mu[c] : 0.f);
      v[t] = x;
    }
    if (zB) {
      __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) split_bf16(v[t], hi[t], lo[t]);
      *reinterpret_cast<uint2*>(zB + i) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(zB + plane + i) = *reinterpret_cast<const uint2*>(lo);
    }
    if (zT) {
      float hi[4], lo[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) split_tf32(v[t], hi[t], lo[t]);
      *reinterpret_cast<float4*>(zT + i) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(zT + plane + i) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// The same z0 with the energy operand in fp16 hi/lo form: one warp per row (walker), which first finds the row's max-abs and
// scales the row by a power of two (row_scale[r], read back by the epilogues).  zB = bf16 hi/lo [2, rA, K2] (unscaled, apply
// GEMM), zH = fp16 hi/lo [2, rA, K2].  Rows >= `rows` and columns >= d are zero.  dK <= 32 * 4 * kZ0Chunks.
constexpr int kZ0Chunks = 8;      // float4 chunks per lane and half: dK <= 1024
__global__ void __launch_bounds__(256) k_z0_planes_f16(const float* __restrict__ X1, const float* __restrict__ mu,
                                                       const float* __restrict__ S0, int rows, int d, int rA, int dK,
                                                       __nv_bfloat16* __restrict__ zB, __half* __restrict__ zH,
                                                       float* __restrict__ row_scale) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rA) return;
  const int K2 = 2 * dK;
  const size_t plane = (size_t)rA * K2;
  float v[2][kZ0Chunks][4];
  float m = 0.f;
#pragma unroll
  for (int part = 0; part < 2; ++part)
#pragma unroll
    for (int t = 0; t < kZ0Chunks; ++t) {
      const int c0 = 4 * lane + 128 * t;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u;
        float x = 0.f;
        if (r < rows && c < d) x = part ? S0[(size_t)r * d + c] : X1[(size_t)r * d + c] - (mu ? mu[c] : 0.f);
        v[part][t][u] = x;
        m = fmaxf(m, fabsf(x));
      }
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float sc = f16_scale(__float_as_uint(m), kF16TopZ);
  if (lane == 0 && r < rows) row_scale[r] = sc;
#pragma unroll
  for (int part = 0; part < 2; ++part)
#pragma unroll
    for (int t = 0; t < kZ0Chunks; ++t) {
      const int c0 = 4 * lane + 128 * t;
      if (c0 >= dK) continue;
      const size_t i = (size_t)r * K2 + (size_t)part * dK + c0;
      __align__(8) __nv_bfloat16 bh[4], bl[4];
      __align__(8) __half hh[4], hl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        split_bf16(v[part][t][u], bh[u], bl[u]);
        split_f16(v[part][t][u] * sc, hh[u], hl[u]);
      }
      if (zB) {
        *reinterpret_cast<uint2*>(zB + i) = *reinterpret_cast<const uint2*>(bh);
        *reinterpret_cast<uint2*>(zB + plane + i) = *reinterpret_cast<const uint2*>(bl);
      }
      *reinterpret_cast<uint2*>(zH + i) = *reinterpret_cast<const uint2*>(hh);
      *reinterpret_cast<uint2*>(zH + plane + i) = *reinterpret_cast<const uint2*>(hl);
    }
}

// Start of a kick sequence with fp16 hi/lo operands: Q <- X1, planes <- fp16 hi/lo of (X1 - mu) * row_scale[r], where the
// power-of-two row scale comes from max(|X1 - mu|, trajectory length * |S0|) of the row (what q - mu can grow to along the
// trajectory, with 16x head-room on top).  One warp per row; dK <= 128 * kZ0Chunks.
__global__ void __launch_bounds__(256) k_tc_init_f16(const float* __restrict__ X1, const float* __restrict__ mu,
                                                     const float* __restrict__ S0, float traj, int rows, int d, int rows_pad, int Kp,
                                                     float* __restrict__ Q, __half* __restrict__ planes, float* __restrict__ row_scale) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows_pad) return;
  const size_t plane = (size_t)rows_pad * Kp;
  float v[kZ0Chunks][4];
  float m = 0.f;
#pragma unroll
  for (int t = 0; t < kZ0Chunks; ++t) {
    const int c0 = 4 * lane + 128 * t;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u;
      float x = 0.f;
      if (r < rows && c < d) {
        const float xr = X1[(size_t)r * d + c];
        Q[(size_t)r * d + c] = xr;
        x = xr - (mu ? mu[c] : 0.f);
        m = fmaxf(m, fmaxf(fabsf(x), traj * fabsf(S0[(size_t)r * d + c])));
      }
      v[t][u] = x;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float sc = f16_scale(__float_as_uint(m), kF16TopZ);
  if (lane == 0 && r < rows) row_scale[r] = sc;
#pragma unroll
  for (int t = 0; t < kZ0Chunks; ++t) {
    const int c0 = 4 * lane + 128 * t;
    if (c0 >= Kp) continue;
    const size_t i = (size_t)r * Kp + c0;
    __align__(8) __half hh[4], hl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) split_f16(v[t][u] * sc, hh[u], hl[u]);
    *reinterpret_cast<uint2*>(planes + i) = *reinterpret_cast<const uint2*>(hh);
    *reinterpret_cast<uint2*>(planes + plane + i) = *reinterpret_cast<const uint2*>(hl);
  }
}

// B-operand planes of a propagator block: planes[row_off + c][kk] = split(src[k(kk)][c]) for c < d, where src is
// [2d, d] (basis index k major) and kk -> k maps the z0 column layout (kk < dK: k = kk; else k = d + kk - dK);
// zero outside.  Covers rows [row_off, row_off + rows_blk) of both planes ([2, rows_pad, K2]).
template <int KIND>
__global__ void k_transpose_planes(const float* __restrict__ src, int d, int dK, int rows_blk, int rows_pad, int row_off,
                                   typename Kind<KIND>::elem* __restrict__ planes, const uint32_t* __restrict__ amax = nullptr) {
  __shared__ float tile[32][33];
  const float sc16 = (KIND == KIND_F16) ? f16_scale(*amax, kF16TopB) : 1.f;
  const int K2 = 2 * dK;
  const int kk0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {       // i: kk index, threadIdx.x: c index (coalesced over c)
    const int kk = kk0 + i, c = c0 + threadIdx.x;
    const int part = kk >= dK, kc = kk - part * dK;
    tile[i][threadIdx.x] = (kc < d && c < d) ? src[(size_t)(part * d + kc) * d + c] : 0.f;
  }
  __syncthreads();
  const size_t plane = (size_t)rows_pad * K2;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {       // i: c index, threadIdx.x: kk index (coalesced over kk)
    const int c = c0 + i, kk = kk0 + threadIdx.x;
    if (c < rows_blk && kk < K2) {
      const size_t o = (size_t)(row_off + c) * K2 + kk;
      const float v = tile[threadIdx.x][i];
      if (KIND == KIND_F16) {
        __half hi, lo;
        split_f16(v * sc16, hi, lo);
        reinterpret_cast<__half*>(planes)[o] = hi;
        reinterpret_cast<__half*>(planes)[plane + o] = lo;
      } else if (KIND == KIND_BF16) {
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        reinterpret_cast<__nv_bfloat16*>(planes)[o] = hi;
        reinterpret_cast<__nv_bfloat16*>(planes)[plane + o] = lo;
      } else {
        float hi, lo;
        split_tf32(v, hi, lo);
        reinterpret_cast<float*>(planes)[o] = hi;
        reinterpret_cast<float*>(planes)[plane + o] = lo;
      }
    }
  }
}

// out[i] = sum_z part[z][i] in fixed order (split-K slices)
__global__ void k_sum_slices_f(const float* __restrict__ part, int nslices, size_t count, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < nslices; ++z) s += part[(size_t)z * count + i];
    out[i] = s;
  }
}

struct FinishArgs {
  const float* lpp; const float* dkp; int NT;
  const float* lp1; const float* Q;
  float* logacc; float* lp_out;
  int rows, d, n, row0, has_lower0; float lower0;
  int fuse, legacy; Key akey; const uint32_t* akey_ptr;
  float* X1w; float* lp1w; int32_t* accepts; float* chain_out; float* chain_lp; int32_t* nonfinite;
  int push; DistPush dp;
};

// one warp per walker: energies, log-accept and (optionally) the Metropolis accept + chain emit.  In row-sharded runs the
// same stores also go into every peer's copy of the ensemble over NVLink (the exchange step of the half-move, fused), and
// the last CTA publishes the half-move's epoch (hx_dist.cuh).
__global__ void __launch_bounds__(256) k_tc_finish(FinishArgs a) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r < a.rows) {
    float lp = 0.f, dk = 0.f;
    for (int t = 0; t < a.NT; ++t) { lp += a.lpp[(size_t)r * a.NT + t]; dk += a.dkp[(size_t)r * a.NT + t]; }
    lp *= -0.5f;
    dk *= -0.5f;
    if (a.has_lower0 && a.Q[(size_t)r * a.d] < a.lower0) lp = -CUDART_INF_F;
    const float lp1 = a.lp1[r];
    float dH = (lp1 - lp) + dk;                              // hmc_walk.py:54
    if (isnan(dH)) dH = CUDART_INF_F;                        // hmc_walk.py:55
    const float la = fminf(0.f, -dH);                        // hmc_walk.py:56
    if (lane == 0) { a.logacc[r] = la; a.lp_out[r] = lp; }
    if (a.fuse) {
      Key akey = a.akey;
      if (a.akey_ptr) { akey.k0 = a.akey_ptr[0]; akey.k1 = a.akey_ptr[1]; }
      const float lo = 1e-10f, span = 1.0f - lo;
      const float u = uniform_rt<float>(akey, (uint64_t)(a.row0 + r), (uint64_t)a.n, lo, span, a.legacy);
      const bool acc = logf(u) < la;                           // sampler_utils.py:114-115
      const float lps = acc ? lp : lp1;
      if (lane == 0) {
        a.lp1w[r] = lps;
        if (a.chain_lp) a.chain_lp[r] = lps;
        a.accepts[r] += acc ? 1 : 0;
        if (!isfinite(lps) && a.nonfinite) atomicOr(a.nonfinite, 1);
        if (a.push)
          for (int p = 0; p < a.dp.world; ++p)
            if (p != a.dp.rank) reinterpret_cast<float*>(static_cast<char*>(a.dp.pt.base[p]) + a.dp.lp_off)[r] = lps;
      }
      const size_t off = (size_t)r * a.d;
      if (a.push && (a.d & 3) == 0) {
        // 16-byte stores: one float4 per lane and step, replicated to every peer
        const int d4 = a.d >> 2;
        const float4* q4 = reinterpret_cast<const float4*>(a.Q + off);
        float4* x4 = reinterpret_cast<float4*>(a.X1w + off);
        float4* c4 = a.chain_out ? reinterpret_cast<float4*>(a.chain_out + off) : nullptr;
        const bool cvec = a.chain_out && ((reinterpret_cast<uintptr_t>(a.chain_out + off) & 15) == 0);
        for (int c = lane; c < d4; c += 32) {
          const float4 v = acc ? q4[c] : x4[c];
          if (acc) x4[c] = v;
          if (cvec) c4[c] = v;
          else if (a.chain_out) { float* co = a.chain_out + off + 4 * c; co[0] = v.x; co[1] = v.y; co[2] = v.z; co[3] = v.w; }
#pragma unroll 1
          for (int p = 0; p < a.dp.world; ++p)
            if (p != a.dp.rank) reinterpret_cast<float4*>(static_cast<char*>(a.dp.pt.base[p]) + a.dp.x_off)[(off >> 2) + c] = v;
        }
      } else {
        for (int c = lane; c < a.d; c += 32) {
          const float v = acc ? a.Q[off + c] : a.X1w[off + c];
          a.X1w[off + c] = v;
          if (a.chain_out) a.chain_out[off + c] = v;
          if (a.push)
            for (int p = 0; p < a.dp.world; ++p)
              if (p != a.dp.rank) reinterpret_cast<float*>(static_cast<char*>(a.dp.pt.base[p]) + a.dp.x_off)[off + c] = v;
        }
      }
    }
  }
  if (a.push) dist_publish(a.dp);
}

// ---- GEMM launcher ---------------------------------------------------------------------------------------------------
// rowsA_pad / rowsB_pad = plane stride in rows; mapA_rows = rows covered by the A tensor map (0: both full planes);
// n_grid = padded N the grid covers (0: g.N)
template <int EPI, int KIND, int BNT = BN>
static int launch_gemm(hx_ctx* ctx, const void* A, int rowsA_pad, const void* B, int rowsB_pad, int Kp, GemmArgs g,
                       cudaStream_t s, uint64_t mapA_rows = 0, int n_grid = 0, int slices = 1) {
  CUtensorMap tmA, tmB;
  if (int e = make_tmap(ctx, &tmA, A, mapA_rows ? mapA_rows : 2ull * rowsA_pad, Kp, kBM, KIND)) return e;
  if (int e = make_tmap(ctx, &tmB, B, 2ull * rowsB_pad, Kp, BNT, KIND)) return e;
  g.Kp = Kp; g.rowsA_pad = rowsA_pad; g.rowsB_pad = rowsB_pad;
  g.row_epilogue = ctx->row_epilogue | (g.epi_force_row ? (1 << EPI) : 0);
  if (int e = ensure_smem_attr(ctx, (const void*)k_tc_gemm<BNT, EPI, KIND>, SmemLayout<BNT>::BYTES)) return e;
  dim3 grid(((n_grid ? n_grid : g.N) + BNT - 1) / BNT, (g.M + kBM - 1) / kBM, slices);
  k_tc_gemm<BNT, EPI, KIND><<<grid, kThreads, SmemLayout<BNT>::BYTES, s>>>(tmA, tmB, g);
  HX_LAUNCH_CHECK("k_tc_gemm");
  ctx->launches += 1;
  return 0;
}

// CTA-pair variant (k_tc_gemm_pair): 256-column tiles, grid.y rounded up to whole pairs (the odd tail tile reads zero rows
// through the tensor map's out-of-bounds fill and stores nothing)
template <int EPI, int KIND>
static int launch_gemm_pair(hx_ctx* ctx, const void* A, int rowsA_pad, const void* B, int rowsB_pad, int Kp, GemmArgs g,
                            cudaStream_t s, uint64_t mapA_rows = 0, int slices = 1, int n_grid = 0) {
  CUtensorMap tmA, tmB;
  if (int e = make_tmap(ctx, &tmA, A, mapA_rows ? mapA_rows : 2ull * rowsA_pad, Kp, kBM, KIND)) return e;
  if (int e = make_tmap(ctx, &tmB, B, 2ull * rowsB_pad, Kp, 128, KIND)) return e;
  g.Kp = Kp; g.rowsA_pad = rowsA_pad; g.rowsB_pad = rowsB_pad; g.n_cover = n_grid; g.row_epilogue = ctx->row_epilogue;
  if (int e = ensure_smem_attr(ctx, (const void*)k_tc_gemm_pair<EPI, KIND>, SmemLayoutPair::BYTES)) return e;
  const int mt = (g.M + kBM - 1) / kBM;
  dim3 grid((((n_grid ? n_grid : g.N) + 255) / 256) * ((mt + 1) & ~1), 1, slices);      // 1-D: see the tile mapping in the kernel
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = SmemLayoutPair::BYTES; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, k_tc_gemm_pair<EPI, KIND>, tmA, tmB, g);
  if (le != cudaSuccess) {
    int nclusters = -1;
    cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, k_tc_gemm_pair<EPI, KIND>, &cfg);
    (void)cudaGetLastError();
    return hx_fail((int)le, "launch k_tc_gemm_pair grid (%u,%u,%u) smem %d failed: %s (max active clusters %d, query %s)", grid.x, grid.y,
                   grid.z, (int)SmemLayoutPair::BYTES, cudaGetErrorString(le), nclusters, cudaGetErrorString(oe));
  }
  ctx->launches += 1;
  return 0;
}


static inline int pad_to(int x, int m) { return ((x + m - 1) / m) * m; }
static inline int grid_for(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return (int)g;
}


// ---- fused draw + projection (hx_tc_fused.cuh) --------------------------------------------------------------------------
template <int NW, bool SHARE, int FMT>
static int launch_fused_t(hx_ctx* ctx, const CUtensorMap& tmB, const FusedArgs& f, int tiles, int slices, cudaStream_t s) {
  auto kern = k_tc_proj_fused<NW, SHARE, FMT>;
  if (int e = ensure_smem_attr(ctx, (const void*)kern, FusedSmem::BYTES)) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(SHARE ? 2 * tiles : tiles, slices); cfg.blockDim = dim3(64 + 32 * NW);
  cfg.dynamicSmemBytes = FusedSmem::BYTES; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = SHARE ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tmB, f);
  if (le != cudaSuccess) {
    (void)cudaGetLastError();
    return hx_fail((int)le, "launch k_tc_proj_fused<%d,%d> grid (%u,%u) smem %d failed: %s", NW, (int)SHARE, cfg.gridDim.x, cfg.gridDim.y,
                   (int)FusedSmem::BYTES, cudaGetErrorString(le));
  }
  ctx->launches += 1;
  return 0;
}
static bool fused_eligible(const hx_ctx* ctx, int impl, int nprod, int d) {
  return ctx->proj_fused && impl == HX_RNG_PARTITIONABLE && nprod == 3 && ctx->precision != HX_PREC_F16F8 && d <= 1024 &&
         ctx->proj_slices <= 1;
}
// split-K of the fused projection: a function of n ONLY (never of the row shard), so that any row sharding reproduces the
// unsharded result bit for bit.  Two slices halve the wave-quantisation loss of one-CTA-per-SM tiles (c5: 256 pairs on 74
// pair slots = 3.46 waves -> 6.92 half-length waves) and keep each slice's share of C^T (67 MB) resident in L2.
static int fused_slices(const hx_ctx* ctx, int nkb) {
  int s = ctx->fused_slices > 0 ? ctx->fused_slices : (nkb >= 64 ? 2 : 1);
  if (s > nkb / 16) s = nkb / 16 > 0 ? nkb / 16 : 1;
  return s;
}
// operand form of the fused projection: 1 = bf16 hi/lo (three kind::f16 products), 2 = fp16 main + two fp8 products at true scale
static int fused_format(const hx_ctx* ctx) {
  if (ctx->fused_fmt == 1 || ctx->fused_fmt == 2) return ctx->fused_fmt;
  return ctx->precision == HX_PREC_AUTO ? 2 : 1;
}
// S0[rows, d] = normal(key, (n, n))[row0 : row0 + rows] @ C, C^T given as bf16 hi/lo planes [2, dAB, nK] (scale8 == NULL) or as
// fp16 + fp8 planes scaled by scale8[0] (k_center_planes_T fmt8 = 2)
static int launch_proj_fused(hx_ctx* ctx, const void* CT, int dAB, int nK, int n, int row0, int rows, int d, Key key,
                             const uint32_t* key_ptr, float* S0, cudaStream_t s, const float* scale8 = nullptr) {
  CUtensorMap tmB;
  if (int e = make_tmap(ctx, &tmB, CT, 2ull * dAB, nK, 256, KIND_BF16)) return e;
  FusedArgs f;
  memset(&f, 0, sizeof(f));
  f.key = key; f.key_ptr = key_ptr; f.n = n; f.row0 = row0; f.rows = rows; f.d = d; f.num_kb = nK / 64;
  f.rowsB_pad = dAB; f.out = S0; f.ldo = d; f.diag = ctx->fused_diag; f.relaxed_wait = ctx->fused_relaxed;
  f.scale_ptr = scale8; f.stagger_ns = ctx->fused_stagger_ns;
  const bool share = d > 512;
  f.units = share ? 2 : (d + 255) / 256;
  const int tiles = (rows + kBM - 1) / kBM;
  int slices = fused_slices(ctx, f.num_kb);
  const int kbs = (f.num_kb + slices - 1) / slices;
  slices = (f.num_kb + kbs - 1) / kbs;
  void* pS0P = nullptr;
  const bool atomic2 = slices == 2 && !ctx->fused_no_atomic;
  if (atomic2) {
    HX_CUDA(cudaMemsetAsync(S0, 0, (size_t)rows * d * sizeof(float), s));
    f.kb_per_slice = kbs; f.slice_stride = 0; f.atomic_out = 1;
  } else if (slices > 1) {
    if (int e = ws_get(ctx, WS_TC_S0P, (size_t)slices * rows * d * sizeof(float), &pS0P)) return e;
    f.out = (float*)pS0P; f.kb_per_slice = kbs; f.slice_stride = (size_t)rows * d;
  }
  // RNG warps per CTA.  A CTA pair deals 8 items per K-block: with 24 warps every warp keeps its row group and steps 3 K-blocks per
  // item (constant index increments, loop-invariant store offsets): 4.53 ms alone at c5 against 4.67 with 28 warps (round 2)
  const int nw = ctx->fused_nw > 0 ? ctx->fused_nw : (share ? 24 : 28);
  int e;
#define HX_FUSED_NW(NWV) \
  case NWV: e = scale8 ? (share ? launch_fused_t<NWV, true, 1>(ctx, tmB, f, tiles, slices, s) : launch_fused_t<NWV, false, 1>(ctx, tmB, f, tiles, slices, s)) \
                       : (share ? launch_fused_t<NWV, true, 0>(ctx, tmB, f, tiles, slices, s) : launch_fused_t<NWV, false, 0>(ctx, tmB, f, tiles, slices, s)); break;
  switch (nw) {
    HX_FUSED_NW(8) HX_FUSED_NW(16) HX_FUSED_NW(20) HX_FUSED_NW(24) HX_FUSED_NW(28)
    default: return hx_fail(HX_E_UNSUPPORTED, "fused projection: %d RNG warps per CTA is not instantiated (8, 16, 20, 24, 28)", nw);
  }
#undef HX_FUSED_NW
  if (e) return e;
  if (slices > 1 && !atomic2) {
    k_sum_slices_f<<<grid_for((size_t)rows * d, 256), 256, 0, s>>>((const float*)pS0P, slices, (size_t)rows * d, S0);
    HX_LAUNCH_CHECK("k_sum_slices_f");
    ctx->launches += 1;
  }
  return 0;
}

struct LeapArgs {
  const float* X1; const float* mu; int rows, d, rA, dK, dB, NT, L; double eps;
  float* Q; float* S; float* S0; float* Qs; const float* Bo; const float* prec; int ld_prec;
  void* pBoP; void* pPP; void* pQC; void* pQC2; void* pQsP; float* lpp; float* dkp;
  int energies;     // 0: integrate only (propagator basis run)
  // KIND_F16: max-abs words ([0] P, [2] Bop), per-row scales of the q - mu operand (written by k_tc_init_f16)
  uint32_t* am; float* rs;
};

// L+1 kicks in B-form (every kick but the last followed by a drift; hmc_walk.py:85-102), then the two energy products.
template <int KIND>
static int leapfrog_bform(hx_ctx* ctx, const LeapArgs& a, cudaStream_t s) {
  typedef typename Kind<KIND>::elem E;
  const int rows = a.rows, d = a.d, rA = a.rA, dK = a.dK, dB = a.dB;
  if (KIND == KIND_F16) {
    // fp16 hi/lo operands: Bop (and P for the energies) scaled by one power of two each from their max-abs, q - mu per row
    k_absmax_ld<<<grid_for((size_t)d * d / 4, 256), 256, 0, s>>>(a.Bo, d, d, d, a.am + 2);
    HX_LAUNCH_CHECK("k_absmax_ld(Bop)");
    k_split_planes<KIND_F16><<<grid_for((size_t)dB * dK, 256), 256, 0, s>>>(a.Bo, d, d, d, dB, dK, (__half*)a.pBoP, a.am + 2);
    HX_LAUNCH_CHECK("k_split_planes(Bop, f16)");
    if (a.energies) {
      k_absmax_ld<<<grid_for((size_t)d * d / 4, 256), 256, 0, s>>>(a.prec, d, d, a.ld_prec, a.am);
      HX_LAUNCH_CHECK("k_absmax_ld(P)");
      k_split_planes<KIND_F16><<<grid_for((size_t)dB * dK, 256), 256, 0, s>>>(a.prec, d, d, a.ld_prec, dB, dK, (__half*)a.pPP, a.am);
      HX_LAUNCH_CHECK("k_split_planes(P, f16)");
      ctx->launches += 2;
    }
    k_tc_init_f16<<<(rA + 7) / 8, 256, 0, s>>>(a.X1, a.mu, a.S0, (float)(a.eps * a.L), rows, d, rA, dK, a.Q, (__half*)a.pQC, a.rs);
    HX_LAUNCH_CHECK("k_tc_init_f16");
    ctx->launches += 3;
  } else {
    k_split_planes<KIND><<<grid_for((size_t)dB * dK, 256), 256, 0, s>>>(a.Bo, d, d, d, dB, dK, (E*)a.pBoP);
    HX_LAUNCH_CHECK("k_split_planes(Bop)");
    k_split_planes<KIND><<<grid_for((size_t)dB * dK, 256), 256, 0, s>>>(a.prec, d, d, a.ld_prec, dB, dK, (E*)a.pPP);
    HX_LAUNCH_CHECK("k_split_planes(P)");
    k_tc_init<KIND><<<grid_for((size_t)rA * dK, 256), 256, 0, s>>>(a.X1, a.mu, rows, d, rA, dK, a.Q, (E*)a.pQC);
    HX_LAUNCH_CHECK("k_tc_init");
    ctx->launches += 3;
  }
  GemmArgs gg;
  memset(&gg, 0, sizeof(gg));
  gg.nprod = 3; gg.M = rows; gg.N = d; gg.mu = a.mu; gg.NT = a.NT; gg.eps = (float)a.eps;
  if (KIND == KIND_F16) { gg.amax = a.am; gg.row_scale = a.rs; gg.unscale_idx = 2; }
  // The q - mu operand planes ping-pong between two buffers: a kick's epilogue writes the NEXT operand while sibling CTAs of the
  // same row tile (other column tiles) may still be reading the current one (updating in place is a race).
  void* planes[2] = {a.pQC, a.pQC2};
  const bool narrow = ((rows + kBM - 1) / kBM) * ((d + BN - 1) / BN) * 2 <= ctx->sm_count;
  const bool narrow64 = ((rows + kBM - 1) / kBM) * ((d + 127) / 128) * 2 <= ctx->sm_count;      // still under half the SMs at 128
  for (int k = 0; k <= a.L; ++k) {
    GemmArgs x = gg;
    x.S_in = (k == 0) ? a.S0 : a.S; x.S = a.S; x.Q = a.Q; x.Qs = a.Qs; x.first = (k == 0) ? 1 : 0;
    x.planes_out = planes[(k + 1) & 1]; x.rows_out_pad = rA; x.Kout_p = dK;
    x.a = (float)((k == 0 || k == a.L) ? 0.5 * (float)a.eps : (float)a.eps);
    x.drift = k < a.L ? 1 : 0;
    // many rows (stepwise kicks over the walkers): the epilogue is bound by its HBM traffic and the coalesced chunk path only adds
    // shared-memory round trips (c5, 32768 rows: 5.9 -> 7.0 ms per half-move); few rows (latency-bound): coalesced (-6 %)
    x.epi_force_row = !(narrow || narrow64);
    // few rows (the propagator's 2d basis trajectories): 128-wide tiles double the CTAs and halve both the MMA chain and
    // the per-thread epilogue of these latency-bound launches
    if (narrow64) {
      if (int e = launch_gemm<EPI_KICKB, KIND, 64>(ctx, planes[k & 1], rA, a.pBoP, dB, dK, x, s)) return e;
    } else if (narrow) {
      if (int e = launch_gemm<EPI_KICKB, KIND, 128>(ctx, planes[k & 1], rA, a.pBoP, dB, dK, x, s)) return e;
    } else {
      if (int e = launch_gemm<EPI_KICKB, KIND>(ctx, planes[k & 1], rA, a.pBoP, dB, dK, x, s)) return e;
    }
  }
  void* planes_L = planes[a.L & 1];      // q_L - mu (written by the last drift, kick L-1)
  if (!a.energies) return 0;
  // energies: dK = -1/2 <Qs P, S0 + S_L> ; lp' = -1/2 <(q_L - mu) P, q_L - mu>
  if (KIND == KIND_F16)
    k_split_planes<KIND_F16><<<grid_for((size_t)rA * dK, 256), 256, 0, s>>>(a.Qs, rows, d, d, rA, dK, (__half*)a.pQsP, nullptr, a.rs);
  else
    k_split_planes<KIND><<<grid_for((size_t)rA * dK, 256), 256, 0, s>>>(a.Qs, rows, d, d, rA, dK, (E*)a.pQsP);
  HX_LAUNCH_CHECK("k_split_planes(Qs)");
  ctx->launches += 1;
  if (KIND == KIND_F16) gg.dot_pair = 2;      // both energy products are (row-scaled operand) x P
  GemmArgs x = gg;
  x.mu = nullptr; x.X = a.S0; x.Y = a.S; x.part = a.dkp;
  if (int e = launch_gemm<EPI_DOT, KIND>(ctx, a.pQsP, rA, a.pPP, dB, dK, x, s)) return e;
  GemmArgs y = gg;
  y.X = a.Q; y.Y = nullptr; y.part = a.lpp;
  return launch_gemm<EPI_DOT, KIND>(ctx, planes_L, rA, a.pPP, dB, dK, y, s);
}

bool walk_eligible(const hx_ctx* ctx, int dtype, const hx_target* tgt, int n, int rows, int flags) {
  if (ctx->precision == HX_PREC_EXACT) return false;
  if (dtype != HX_F32 || tgt->kind != HX_TARGET_GAUSS_FULL || flags != 0) return false;
  // measured on B200 (bench.py --workload m1..m6): the tensor-core path overtakes the SIMT kernels from about dim 128 x 2048
  // walkers per group (its ~30 launches per half-move cost ~0.35 ms whatever the size)
  if (ctx->precision == HX_PREC_AUTO) return tgt->dim >= 128 && n >= 2048 && rows >= 128;
  return tgt->dim >= 16 && n >= 64;      // forced tensor-core modes (tests)
}

// ---- momentum-draw jobs ------------------------------------------------------------------------------------------------
// P0 = normal(key, (n, n))[row0 : row0 + rows] does not depend on the ensemble, only on the key.  A job draws it as bf16
// hi/lo planes [2, rA, nK] into one of two buffers on the low-priority side stream, chunk by chunk (one event per
// chunk), so that (a) the projection GEMM of chunk c overlaps the draw of chunk c+1 and (b) the draw for the NEXT
// half-move (key known in advance) runs under the current half-move's tensor-core work.
static int side_init(hx_ctx* ctx) {
  if (ctx->side_ready) return 0;
  int prio_lo = 0, prio_hi = 0;
  HX_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  HX_CUDA(cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_lo));
  HX_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  for (int b = 0; b < 2; ++b) {
    ctx->job[b].valid = 0;
    ctx->ev_chunk[b] = new std::vector<cudaEvent_t>();
  }
  // same L1/shared carve-out as the GEMM kernel, otherwise the two kernels cannot share an SM
  HX_CUDA(cudaFuncSetAttribute(k_rng_planes<false, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  HX_CUDA(cudaFuncSetAttribute(k_rng_planes<true, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  HX_CUDA(cudaFuncSetAttribute(k_rng_planes<false, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  HX_CUDA(cudaFuncSetAttribute(k_rng_planes<true, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  HX_CUDA(cudaFuncSetAttribute(k_rng_planes<false, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  ctx->side_ready = 1;
  return 0;
}

static bool job_matches(const P0Job& j, int fmt, int impl, int n, int row0, int rows, Key key, const uint32_t* key_ptr) {
  if (!j.valid || j.fmt != fmt || j.impl != impl || j.n != n || j.row0 != row0 || j.rows != rows) return false;
  if (key_ptr || j.key_ptr) return key_ptr == j.key_ptr;
  return key.k0 == j.k0 && key.k1 == j.k1;
}

static int chunk_rows(const hx_ctx* ctx, int rA, int NT) {
  int tiles = ctx->sm_count / NT;             // projection GEMM of one chunk = one wave of 128 x BN tiles
  if (ctx->proj_pair && tiles > 1) tiles &= ~1;      // CTA pairs own two consecutive row tiles
  int chunk = tiles * kBM;
  if (chunk < kBM) chunk = kBM;
  if (chunk > rA) chunk = rA;
  return chunk;
}

// enqueue the draw of (key, row0, rows) into buffer b on the side stream
static int job_launch(hx_ctx* ctx, int b, __nv_bfloat16* buf, int fmt, int impl, int n, int row0, int rows, int rA, int nK, int chunk,
                      Key key, const uint32_t* key_ptr, cudaStream_t s) {
  const int nchunks = (rows + chunk - 1) / chunk;
  std::vector<cudaEvent_t>& evs = *ctx->ev_chunk[b];
  while ((int)evs.size() < nchunks) {
    cudaEvent_t e;
    HX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    evs.push_back(e);
  }
  // ordered after everything already enqueued on s (the buffer's previous consumer in particular)
  HX_CUDA(cudaEventRecord(ctx->ev_fork, s));
  HX_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
  const size_t plane = (size_t)rA * nK;
  void* pCtr;
  if (int e = ws_get(ctx, WS_TC_CTR, 2 * 512 * sizeof(unsigned int), &pCtr)) return e;
  if (nchunks > 512) return hx_fail(HX_E_SHAPE, "momentum draw: %d row chunks (max 512)", nchunks);
  unsigned int* ctr = reinterpret_cast<unsigned int*>(pCtr) + (size_t)b * 512;
  HX_CUDA(cudaMemsetAsync(ctr, 0, (size_t)nchunks * sizeof(unsigned int), ctx->side));
  for (int c = 0; c < nchunks; ++c) {
    const int r0 = c * chunk;
    const int rc = (rows - r0) < chunk ? (rows - r0) : chunk;
    const int rcp = pad_to(rc, kBM);
    // 4 CTAs (32 warps) per SM: measured (scripts/overlap_bench.py) to overlap fully with a co-resident GEMM CTA, whereas
    // 1-2 CTAs/SM serialise (too few warps to cover the ALU latency once the GEMM's warps take issue slots)
    int rng_grid = grid_for((size_t)rcp * nK / 8, 256);
    const int cap = ctx->sm_count * (ctx->rng_ctas_per_sm > 0 ? ctx->rng_ctas_per_sm : 4);
    if (rng_grid > cap) rng_grid = cap;
    __nv_bfloat16* dst = buf + (size_t)r0 * nK;
    if (ctx->draw_nostore) {
      k_rng_planes<false, 2><<<rng_grid, 256, 0, ctx->side>>>(key, key_ptr, n, row0 + r0, rc, rcp, nK, plane, dst, ctr + c);
    } else if (impl == HX_RNG_LEGACY) {
      if (fmt) k_rng_planes<true, 1><<<rng_grid, 256, 0, ctx->side>>>(key, key_ptr, n, row0 + r0, rc, rcp, nK, plane, dst, ctr + c);
      else k_rng_planes<true, 0><<<rng_grid, 256, 0, ctx->side>>>(key, key_ptr, n, row0 + r0, rc, rcp, nK, plane, dst, ctr + c);
    } else {
      if (fmt) k_rng_planes<false, 1><<<rng_grid, 256, 0, ctx->side>>>(key, key_ptr, n, row0 + r0, rc, rcp, nK, plane, dst, ctr + c);
      else k_rng_planes<false, 0><<<rng_grid, 256, 0, ctx->side>>>(key, key_ptr, n, row0 + r0, rc, rcp, nK, plane, dst, ctr + c);
    }
    HX_LAUNCH_CHECK("k_rng_planes");
    HX_CUDA(cudaEventRecord(evs[c], ctx->side));
    ctx->launches += 1;
  }
  P0Job& j = ctx->job[b];
  j.valid = 1; j.fmt = fmt; j.impl = impl; j.n = n; j.row0 = row0; j.rows = rows; j.k0 = key.k0; j.k1 = key.k1; j.key_ptr = key_ptr;
  return 0;
}

int chees_inner(hx_ctx* ctx, const float* pproj, int rows, int d, const float* covinv, const float* Xprop, const float* mean_prop,
                float** part_out, int* nt_out, cudaStream_t s) {
  typedef __nv_bfloat16 bf;
  const int dB = pad_to(d, BN), dK = pad_to(d, 64), rA = pad_to(rows, kBM), NT = (d + BN - 1) / BN;
  void *pA, *pB, *pPart;
  if (int e = ws_get(ctx, WS_TC_QC, (size_t)2 * rA * dK * sizeof(float), &pA)) return e;
  if (int e = ws_get(ctx, WS_TC_MP, (size_t)2 * dB * dK * sizeof(float), &pB)) return e;
  if (int e = ws_get(ctx, WS_TC_DKP, (size_t)rows * NT * sizeof(float), &pPart)) return e;
  k_split_planes<KIND_BF16><<<grid_for((size_t)rA * dK, 256), 256, 0, s>>>(pproj, rows, d, d, rA, dK, (bf*)pA);
  k_split_planes<KIND_BF16><<<grid_for((size_t)dB * dK, 256), 256, 0, s>>>(covinv, d, d, d, dB, dK, (bf*)pB);
  HX_LAUNCH_CHECK("k_split_planes");
  ctx->launches += 2;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.nprod = 3; g.M = rows; g.N = d; g.NT = NT; g.X = Xprop; g.mu = mean_prop; g.part = (float*)pPart;
  if (int e = launch_gemm<EPI_DOT, KIND_BF16>(ctx, pA, rA, pB, dB, dK, g, s)) return e;
  *part_out = (float*)pPart;
  *nt_out = NT;
  return 0;
}

int walk_half(hx_ctx* ctx, int impl, const hx_target* tgt, const float* mean_dev, const float* X1, const float* X2, int n,
              int row0, int rows, double eps, Key key, const uint32_t* key_ptr, int L, const float* lp1, float* Q,
              float* logacc, float* S, float* lp_out, const FuseF32* fuse, const NextDraw* next, cudaStream_t s) {
  const int d = tgt->dim;
  const int nprod = (ctx->precision == HX_PREC_BF16) ? 1 : 3;   // momentum projection / Gram only
  // momentum projection from fp16 + e4m3 operands: explicit mode, and what AUTO picks (measured 0.25 ms per half-move faster at
  // c5 than bf16 hi/lo next to the concurrent draw, S0 within 1.7e-5 instead of 1.3e-5 of an f64 product)
  // P0 drawn inside the projection GEMM (bf16 hi/lo operands, one accumulator); otherwise the staged draw + GEMM below
  const bool fused = fused_eligible(ctx, impl, nprod, d);
  const int f8 = (!fused && (ctx->precision == HX_PREC_F16F8 || ctx->precision == HX_PREC_AUTO)) ? 1 : 0;
  const int fmt8 = f8 ? 1 : ((fused && fused_format(ctx) == 2) ? 2 : 0);      // form of the fp16 + fp8 copy of C^T (0: none)
  const int dB = pad_to(d, BN);            // rows of B-operand planes (N side)
  const int dK = pad_to(d, 64);            // K padding when dim is the reduction axis
  const int nK = pad_to(n, 64);            // K padding when the walker index is the reduction axis
  const int rA = pad_to(rows, kBM);        // rows of A-operand planes over walkers
  const int dA = pad_to(d, kBM);
  const int dAB = dB > dA ? dB : dA;       // CT planes serve as both A (128-row boxes) and B (BN-row boxes)
  const int NT = (d + BN - 1) / BN;
  const int K2 = 2 * dK;                   // K of the propagator GEMMs: z0 = [q0 - mu | S0]
  const int bR = 2 * d, bRA = pad_to(bR, kBM);   // propagator basis rows
  typedef __nv_bfloat16 bf;
  // leapfrog form: step-by-step kicks on every walker, or (Gaussian target => linear dynamics) the L-step propagator
  // built once per half-move from 2d basis trajectories and applied to all walkers in one GEMM
  // AUTO: cost model fitted to B200 timings (scripts / DESIGN.md §7).  A kick over `rows` walkers is bound by its epilogue's
  // HBM traffic (28 bytes per element at ~2.6 TB/s effective) with a ~25 us latency floor; the basis run is (L+1)
  // latency-bound launches on 2d rows; the apply + energy products cost about four kicks' worth of traffic.
  bool prop = ctx->leap_form == HX_LEAP_PROPAGATOR;
  if (ctx->leap_form == HX_LEAP_AUTO) {
    const double bw = 2.6e12, elems = (double)rows * d * 28.0;
    // (+ the kick GEMM's own main loop, ~1 us per 64-wide K-block: without it the model picked the stepwise form for 8 ranks at
    //  c5, 4096 rows each, where the propagator form measures 0.75 ms against 0.82-1.0 ms per half-move -- and keeps the 8-GPU
    //  chain bit-equal to the 1/2/4-GPU chains)
    const double t_kick = (elems / bw > 25e-6 ? elems / bw : 25e-6) + 16e-6 * (d / 1000.0);
    const double t_bk = 27e-6 + 15e-6 * (d / 1000.0) * (d / 1000.0);
    const double t_apply = 4.0 * elems / bw > 80e-6 ? 4.0 * elems / bw : 80e-6;
    prop = (L + 3) * t_kick > (L + 1) * t_bk + 60e-6 + t_apply;
  }
  const bool g1_tf32 = ctx->precision == HX_PREC_TF32X3;
  // propagator form: operand kind of the energy products -- tf32 hi/lo (0), bf16 hi/lo (1), fp16 hi/lo with row / matrix scales (2)
  // kicks / apply: bf16 hi/lo, or fp16 hi/lo with row / matrix scales (same MMA rate, 2^-22 instead of 2^-16 per product).
  // hx_ctx_set_tuning(14, v): 0 = auto: fp16 for the stepwise form, whose log-accept error against f64 drops from 1.1e-2 to
  // 7.7e-4 rms at c5 scale; bf16 for the propagator form, whose error (2.9e-3, set by the apply's long accumulation) does not
  // change while the extra max-abs launches cost 0.2 ms; 1 = fp16 for both; 2 = bf16 for both.  fp16 kicks imply fp16 energies.
  const bool kf16 = !g1_tf32 && dK <= 128 * kZ0Chunks && (ctx->kick_f16 == 1 || (ctx->kick_f16 == 0 && !prop));
  const int ekind = g1_tf32 ? 0 : (kf16 ? 2 : ((ctx->energy_bf16 == 2 && dK > 128 * kZ0Chunks) ? 0 : ctx->energy_bf16));
  const bool e_bf16 = ekind == 1, e_f16 = ekind == 2;
  const int rQ = prop ? (rA > bRA ? rA : bRA) : rA;            // rows of the q - mu operand planes (basis run reuses them)

  void *pCT, *pM, *pBo, *pBoP, *pPP, *pP0, *pQC, *pQC2, *pQsP, *pS0, *pQs, *pLPP, *pDKP;
  if (int e = ws_get(ctx, WS_TC_CT, (size_t)2 * dAB * nK * sizeof(bf), &pCT)) return e;
  if (int e = ws_get(ctx, WS_M, (size_t)d * d * sizeof(float), &pM)) return e;
  if (int e = ws_get(ctx, WS_TC_BO, (size_t)d * d * sizeof(float), &pBo)) return e;
  if (int e = ws_get(ctx, WS_TC_MP, (size_t)2 * dB * dK * sizeof(float), &pBoP)) return e;
  if (int e = ws_get(ctx, WS_TC_PP, (size_t)2 * dB * dK * sizeof(float), &pPP)) return e;
  const size_t buf_elems = (size_t)2 * rA * nK;               // one draw buffer = [2 planes, rA, nK]
  pP0 = nullptr;
  if (!fused)
    if (int e = ws_get(ctx, WS_TC_P0, 2 * buf_elems * sizeof(bf), &pP0)) return e;
  if (int e = ws_get(ctx, WS_TC_QC, (size_t)2 * rQ * dK * sizeof(float), &pQC)) return e;
  if (int e = ws_get(ctx, WS_TC_QC2, (size_t)2 * (prop ? bRA : rA) * dK * sizeof(float), &pQC2)) return e;
  if (int e = ws_get(ctx, WS_TC_GP, (size_t)2 * (prop ? bRA : rA) * dK * sizeof(float), &pQsP)) return e;
  if (int e = ws_get(ctx, WS_TC_S0, (size_t)rows * d * sizeof(float), &pS0)) return e;
  if (int e = ws_get(ctx, WS_G, (size_t)rows * d * sizeof(float), &pQs)) return e;
  if (int e = ws_get(ctx, WS_TC_LPP, (size_t)rows * NT * sizeof(float), &pLPP)) return e;
  if (int e = ws_get(ctx, WS_TC_DKP, (size_t)rows * NT * sizeof(float), &pDKP)) return e;
  float* S0 = (float*)pS0;
  float* Qs = (float*)pQs;
  const float* mu = (const float*)tgt->mean;
  if (int e = side_init(ctx)) return e;

  // (0) momentum draw (hmc_walk.py:36): use the prefetched job if it is this half-move's, otherwise start it now; then
  //     queue the next half-move's draw behind it
  const int chunk = chunk_rows(ctx, rA, NT);
  const int nchunks = (rows + chunk - 1) / chunk;
  bf* Pbuf[2] = {(bf*)pP0, (bf*)pP0 + buf_elems};
  if (!fused && (ctx->p0_base != pP0 || ctx->p0_elems != buf_elems)) {    // workspace moved or reshaped: prefetched draws are gone
    ctx->job[0].valid = ctx->job[1].valid = 0;
    ctx->p0_base = pP0; ctx->p0_elems = buf_elems;
  }
  int cur = -1;
  if (!fused)
    for (int b = 0; b < 2; ++b)
      if (job_matches(ctx->job[b], f8, impl, n, row0, rows, key, key_ptr)) cur = b;
  const bool prefetched = cur >= 0;
  if (fused) {
    cur = 0;
    next = nullptr;               // nothing to prefetch: the draw lives inside the projection kernel
  } else {
    if (cur < 0) {
      cur = ctx->job[0].valid ? (ctx->job[1].valid ? 0 : 1) : 0;
      if (int e = job_launch(ctx, cur, Pbuf[cur], f8, impl, n, row0, rows, rA, nK, chunk, key, key_ptr, s)) return e;
    }
    ctx->job[cur].valid = 0;       // consumed by this half-move
  }

  // hx_ctx_set_tuning(6, -2 / -1): queue the next half-move's draw already at the start of the complement statistics / of the
  // propagator basis run (a head start over the projection GEMM, at the price of slowing those latency-bound launches)
  const bool draw_next_early = next && next->rows == rows && ctx->draw_fork < 0;
  if (draw_next_early && ctx->draw_fork <= -2) {
    if (int e = job_launch(ctx, 1 - cur, Pbuf[1 - cur], f8, impl, n, next->row0, next->rows, rA, nK, chunk, next->key, next->key_ptr, s))
      return e;
  }
  // (1) complement statistics: mean_dev already holds mean_0(X2) (hx_api.cu); C^T planes, M = C^T C, Bop = M P
  if (int e = timing_mark(ctx, s, HX_PH_PREP)) return e;
  // The Gram matrix enters every kick through B = P M, where the ill-conditioning of the target amplifies its error.
  // hx_ctx_set_tuning(2, 1) forms it from tf32 hi/lo operands (2^-21 per product) instead of bf16 hi/lo; on B200 that did
  // not improve the proposal against an f64 integration (scripts/slices_sweep.py), so bf16 hi/lo is the default.
  const bool gram32 = ctx->gram_tf32 && nprod == 3;
  void* pCT32 = nullptr;
  if (gram32)
    if (int e = ws_get(ctx, WS_TC_CT32, (size_t)2 * dAB * nK * sizeof(float), &pCT32)) return e;
  void *pCT8 = nullptr, *pSc8 = nullptr;
  if (fmt8) {
    // power-of-two scale of C from max |X2 - mean| (order-independent reduction: bit-reproducible)
    if (int e = ws_get(ctx, WS_TC_CT8, (size_t)2 * dAB * nK * sizeof(uint16_t), &pCT8)) return e;
    if (int e = ws_get(ctx, WS_TC_SC8, 16, &pSc8)) return e;
    uint32_t* amax = reinterpret_cast<uint32_t*>(pSc8) + 2;
    HX_CUDA(cudaMemsetAsync(amax, 0, sizeof(uint32_t), s));
    k_absmax_centred<<<dim3((d + 255) / 256, (n + 63) / 64), 256, 0, s>>>(X2, mean_dev, n, d, amax);
    HX_LAUNCH_CHECK("k_absmax_centred");
    k_f8_scale<<<1, 1, 0, s>>>(amax, n, (float*)pSc8);
    HX_LAUNCH_CHECK("k_f8_scale");
    ctx->launches += 2;
  }
  {
    dim3 g((nK + 255) / 256, (dAB + 31) / 32);
    k_center_planes_T<<<g, 256, 0, s>>>(X2, mean_dev, n, d, dAB, nK, (bf*)pCT, (float*)pCT32, (uint16_t*)pCT8, (const float*)pSc8, fmt8);
    HX_LAUNCH_CHECK("k_center_planes_T");
  }
  GemmArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.nprod = nprod;
  {
    // Gram M = C^T C: only (d/128) x (d/256) output tiles with K = n, so split K over enough slices to fill the SMs (the
    // fp32 slice sum also shortens the tensor core's truncating accumulation chains)
    const int tiles = ((d + BN - 1) / BN) * ((d + kBM - 1) / kBM), nkb = nK / (gram32 ? 32 : 64);
    int slices = ctx->sm_count / tiles;
    if (slices > 16) slices = 16;
    if (slices > nkb) slices = nkb;
    if (slices < 1) slices = 1;
    const int kbs = (nkb + slices - 1) / slices;
    slices = (nkb + kbs - 1) / kbs;
    void* pMs = pM;
    if (slices > 1)
      if (int e = ws_get(ctx, WS_GRAMP, (size_t)slices * d * d * sizeof(float), &pMs)) return e;
    ga.M = d; ga.N = d; ga.out = (float*)pMs; ga.ldo = d;
    ga.kb_per_slice = slices > 1 ? kbs : 0; ga.slice_stride = (size_t)d * d;
    if (gram32) {
      if (int e = launch_gemm<EPI_STORE, KIND_TF32>(ctx, pCT32, dAB, pCT32, dAB, nK, ga, s, 0, 0, slices)) return e;
    } else if (ctx->aux_pair && d > 128) {
      if (int e = launch_gemm_pair<EPI_STORE, KIND_BF16>(ctx, pCT, dAB, pCT, dAB, nK, ga, s, 0, slices)) return e;
    } else {
      if (int e = launch_gemm<EPI_STORE, KIND_BF16>(ctx, pCT, dAB, pCT, dAB, nK, ga, s, 0, 0, slices)) return e;
    }
    if (slices > 1) {
      k_sum_slices_f<<<grid_for((size_t)d * d, 256), 256, 0, s>>>((const float*)pMs, slices, (size_t)d * d, (float*)pM);
      HX_LAUNCH_CHECK("k_sum_slices_f");
      ctx->launches += 1;
    }
    ga.kb_per_slice = 0; ga.slice_stride = 0;
  }
  {
    {
      dim3 g((d + 63) / 64, (d + 63) / 64);
      k_mp_dmma<<<g, 128, 0, s>>>((const float*)pM, d, (const float*)tgt->precision, tgt->ld, d, (float*)pBo, d);
      HX_LAUNCH_CHECK("k_mp_dmma");
    }
  }
  ctx->launches += 2;

  void *pAM = nullptr, *pRS = nullptr, *pRSB = nullptr;
  if (kf16 || (prop && e_f16)) {
    // max-abs words [0] P, [1] W P, [2] Bop, [3] [Tq ; Ts] and the per-row scales of the fp16 operands
    if (int e = ws_get(ctx, WS_TC_AM, 16, &pAM)) return e;
    if (int e = ws_get(ctx, WS_TC_RSC, (size_t)rA * sizeof(float), &pRS)) return e;
    if (int e = ws_get(ctx, WS_TC_RSB, (size_t)bRA * sizeof(float), &pRSB)) return e;
    HX_CUDA(cudaMemsetAsync(pAM, 0, 16, s));
  }

  // (2) propagator (Gaussian target: the leapfrog is a linear map of z0 = [q0 - mu | S0]).  Integrate the 2d basis
  //     trajectories with the same kick/drift recurrence (tf32 hi/lo operands), giving q_L - mu = z0 Tq, S_L = z0 Ts
  //     and the impulse-weighted position sum Qs = z0 W; fold the precision into W for the kinetic-energy product.
  void *pBX = nullptr, *pBS = nullptr, *pTQ = nullptr, *pTS = nullptr, *pTW = nullptr, *pWP = nullptr, *pB1 = nullptr,
       *pB2 = nullptr, *pZB = nullptr, *pZT = nullptr, *pPP32 = nullptr;
  if (draw_next_early && ctx->draw_fork == -1) {
    if (int e = job_launch(ctx, 1 - cur, Pbuf[1 - cur], f8, impl, n, next->row0, next->rows, rA, nK, chunk, next->key, next->key_ptr, s))
      return e;
  }
  if (prop) {
    if (int e = timing_mark(ctx, s, HX_PH_BASIS)) return e;
    const size_t bsz = (size_t)bR * d * sizeof(float);
    if (int e = ws_get(ctx, WS_TC_BX, bsz, &pBX)) return e;
    if (int e = ws_get(ctx, WS_TC_BS, bsz, &pBS)) return e;
    if (int e = ws_get(ctx, WS_TC_TQ, bsz, &pTQ)) return e;
    if (int e = ws_get(ctx, WS_TC_TS, bsz, &pTS)) return e;
    if (int e = ws_get(ctx, WS_TC_TW, bsz, &pTW)) return e;
    if (int e = ws_get(ctx, WS_TC_WP, bsz, &pWP)) return e;
    if (int e = ws_get(ctx, WS_TC_B1, (size_t)2 * (2 * dB) * K2 * sizeof(float), &pB1)) return e;
    if (int e = ws_get(ctx, WS_TC_B2, (size_t)2 * dB * K2 * sizeof(float), &pB2)) return e;
    if (int e = ws_get(ctx, WS_TC_ZB, g1_tf32 ? 16 : (size_t)2 * rA * K2 * sizeof(bf), &pZB)) return e;
    if (int e = ws_get(ctx, WS_TC_ZT, e_bf16 ? 16 : (size_t)2 * rA * K2 * (e_f16 ? sizeof(__half) : sizeof(float)), &pZT)) return e;

    k_basis_init<<<grid_for((size_t)bR * d, 256), 256, 0, s>>>((float*)pBX, (float*)pBS, d);
    HX_LAUNCH_CHECK("k_basis_init");
    LeapArgs la;
    la.X1 = (const float*)pBX; la.mu = nullptr; la.rows = bR; la.d = d; la.rA = bRA; la.dK = dK; la.dB = dB; la.NT = NT; la.L = L;
    la.eps = eps; la.Q = (float*)pTQ; la.S = (float*)pTS; la.S0 = (float*)pBS; la.Qs = (float*)pTW; la.Bo = (const float*)pBo;
    la.prec = (const float*)tgt->precision; la.ld_prec = tgt->ld; la.pBoP = pBoP; la.pPP = pPP; la.pQC = pQC; la.pQC2 = pQC2; la.pQsP = pQsP;
    la.lpp = nullptr; la.dkp = nullptr; la.energies = 0; la.am = (uint32_t*)pAM; la.rs = (float*)pRSB;
    // the basis run uses the operand kind of the apply GEMM: its rounding (2^-16 per bf16x3 product) is of the same order
    // as the apply's, and bf16 halves both the MMA time and the number of TMA round trips of these latency-bound launches
    if (int e = g1_tf32 ? leapfrog_bform<KIND_TF32>(ctx, la, s)
                        : (kf16 ? leapfrog_bform<KIND_F16>(ctx, la, s) : leapfrog_bform<KIND_BF16>(ctx, la, s)))
      return e;
    // tf32 hi/lo planes of the precision matrix for the energy products
    if (int e = ws_get(ctx, WS_TC_PP32, (size_t)2 * dB * dK * sizeof(float), &pPP32)) return e;
    k_split_planes<KIND_TF32><<<grid_for((size_t)dB * dK, 256), 256, 0, s>>>((const float*)tgt->precision, d, d, tgt->ld, dB, dK, (float*)pPP32);
    HX_LAUNCH_CHECK("k_split_planes(P, tf32)");
    // WP = W P  (rows: basis index): A = planes of W, B = planes of P (symmetric)
    k_split_planes<KIND_TF32><<<grid_for((size_t)bRA * dK, 256), 256, 0, s>>>((const float*)pTW, bR, d, d, bRA, dK, (float*)pQsP);
    HX_LAUNCH_CHECK("k_split_planes(W)");
    GemmArgs gw;
    memset(&gw, 0, sizeof(gw));
    gw.nprod = 3; gw.M = bR; gw.N = d; gw.out = (float*)pWP; gw.ldo = d;
    if (e_f16) gw.amax_out = (uint32_t*)pAM + 1;              // max |W P| for the scale of its fp16 planes
    if (int e = launch_gemm<EPI_STORE, KIND_TF32>(ctx, pQsP, bRA, pPP32, dB, dK, gw, s)) return e;
    if (e_f16) {
      // fp16 hi/lo planes of P (over the bf16 planes the basis run no longer needs), scaled from max |P|
      k_absmax_ld<<<grid_for((size_t)d * d / 4, 256), 256, 0, s>>>((const float*)tgt->precision, d, d, tgt->ld, (uint32_t*)pAM);
      HX_LAUNCH_CHECK("k_absmax_ld(P)");
      k_split_planes<KIND_F16><<<grid_for((size_t)dB * dK, 256), 256, 0, s>>>((const float*)tgt->precision, d, d, tgt->ld, dB, dK,
                                                                             (__half*)pPP, (const uint32_t*)pAM);
      HX_LAUNCH_CHECK("k_split_planes(P, f16)");
      ctx->launches += 2;
    }
    // B operands: [Tq ; Ts]^T for the apply GEMM, (W P)^T for the kinetic-energy product
    dim3 tg((K2 + 31) / 32, (dB + 31) / 32), tb(32, 8);
    if (kf16) {
      k_absmax_ld<<<grid_for((size_t)bR * d / 4, 256), 256, 0, s>>>((const float*)pTQ, bR, d, d, (uint32_t*)pAM + 3);
      k_absmax_ld<<<grid_for((size_t)bR * d / 4, 256), 256, 0, s>>>((const float*)pTS, bR, d, d, (uint32_t*)pAM + 3);
      HX_LAUNCH_CHECK("k_absmax_ld(Tq, Ts)");
      k_transpose_planes<KIND_F16><<<tg, tb, 0, s>>>((const float*)pTQ, d, dK, dB, 2 * dB, 0, (__half*)pB1, (const uint32_t*)pAM + 3);
      k_transpose_planes<KIND_F16><<<tg, tb, 0, s>>>((const float*)pTS, d, dK, dB, 2 * dB, dB, (__half*)pB1, (const uint32_t*)pAM + 3);
      ctx->launches += 2;
    } else if (g1_tf32) {
      k_transpose_planes<KIND_TF32><<<tg, tb, 0, s>>>((const float*)pTQ, d, dK, dB, 2 * dB, 0, (float*)pB1);
      k_transpose_planes<KIND_TF32><<<tg, tb, 0, s>>>((const float*)pTS, d, dK, dB, 2 * dB, dB, (float*)pB1);
    } else {
      k_transpose_planes<KIND_BF16><<<tg, tb, 0, s>>>((const float*)pTQ, d, dK, dB, 2 * dB, 0, (bf*)pB1);
      k_transpose_planes<KIND_BF16><<<tg, tb, 0, s>>>((const float*)pTS, d, dK, dB, 2 * dB, dB, (bf*)pB1);
    }
    if (e_f16) k_transpose_planes<KIND_F16><<<tg, tb, 0, s>>>((const float*)pWP, d, dK, dB, dB, 0, (__half*)pB2, (const uint32_t*)pAM + 1);
    else if (e_bf16) k_transpose_planes<KIND_BF16><<<tg, tb, 0, s>>>((const float*)pWP, d, dK, dB, dB, 0, (bf*)pB2);
    else k_transpose_planes<KIND_TF32><<<tg, tb, 0, s>>>((const float*)pWP, d, dK, dB, dB, 0, (float*)pB2);
    HX_LAUNCH_CHECK("k_transpose_planes");
    ctx->launches += 5;
  }

  // (3) projection S0 = P0 C, chunk by chunk as the draw completes.  The NEXT half-move's draw is queued here: it then runs
  //     under the throughput-bound projection / apply GEMMs (measured to overlap fully, scripts/overlap_bench.py) and stays
  //     out of the way of the small latency-bound launches above, which it would slow 3-4x.
  const bool draw_next = next && next->rows == rows && !draw_next_early;
  const int fork_at = ctx->draw_fork < 0 ? 0 : (ctx->draw_fork > nchunks ? nchunks : ctx->draw_fork);   // after that many projection chunks
  if (draw_next && fork_at == 0) {
    if (int e = job_launch(ctx, 1 - cur, Pbuf[1 - cur], f8, impl, n, next->row0, next->rows, rA, nK, chunk, next->key, next->key_ptr, s))
      return e;
  }
  if (fused) {
    if (int e = timing_begin(ctx, s)) return e;
    if (int e = fmt8 == 2 ? launch_proj_fused(ctx, pCT8, dAB, nK, n, row0, rows, d, key, key_ptr, S0, s, (const float*)pSc8)
                          : launch_proj_fused(ctx, pCT, dAB, nK, n, row0, rows, d, key, key_ptr, S0, s)) return e;
    if (int e = timing_mark(ctx, s, HX_PH_LEAP)) return e;
  } else {
    std::vector<cudaEvent_t>& evs = *ctx->ev_chunk[cur];
    if (int e = timing_begin(ctx, s)) return e;
    // optional split-K (hx_ctx_set_tuning(1, slices)): slices of K summed in fp32 in a fixed order.  Measured on B200
    // (scripts/slices_sweep.py): no accuracy gain for the projection, +20 % time, so the default is one slice.
    const int nkb = nK / 64;
    int slices = ctx->proj_slices > 0 ? ctx->proj_slices : 1;
    if (slices > nkb / 16) slices = nkb / 16 > 0 ? nkb / 16 : 1;      // at least 16 K-blocks (1024 walkers) per slice
    const int kbs = (nkb + slices - 1) / slices;
    slices = (nkb + kbs - 1) / kbs;
    void* pS0P = nullptr;
    if (slices > 1)
      if (int e = ws_get(ctx, WS_TC_S0P, (size_t)slices * chunk * d * sizeof(float), &pS0P)) return e;
    // CTA-pair kernel (hx_ctx_set_tuning(8, 1)): when the whole draw was queued ahead of time (the steady state of a batch
    // loop) the projection is ONE launch over all rows -- back-to-back waves without a tail per chunk; otherwise chunk by
    // chunk behind the draw like the single-CTA kernel
    const bool pair = ctx->proj_pair && slices == 1 && d > 128;
    const int group = (prefetched && (pair ? !ctx->pair_chunked : ctx->proj_one_launch != 0)) ? nchunks : 1;   // chunks per launch
    for (int c = 0; c < nchunks; c += group) {
      const int r0 = c * chunk;
      const int cl = (c + group < nchunks ? c + group : nchunks) - 1;      // last chunk of this launch
      const int rend = (cl + 1) * chunk < rows ? (cl + 1) * chunk : rows;
      const int rc = rend - r0;
      HX_CUDA(cudaStreamWaitEvent(s, evs[cl], 0));
      ga.M = rc; ga.N = d; ga.ldo = d;
      ga.out = slices > 1 ? (float*)pS0P : S0 + (size_t)r0 * d;
      ga.kb_per_slice = slices > 1 ? kbs : 0; ga.slice_stride = (size_t)rc * d;
      ga.one_acc = (ctx->one_acc && !f8 && !pair && nprod == 3) ? 1 : 0;
      const uint64_t map_rows = (uint64_t)rA + pad_to(rc, kBM);
      const bf* Ac = Pbuf[cur] + (size_t)r0 * nK;
      if (f8) ga.scale_ptr = (const float*)pSc8;
      int e;
      if (pair) e = f8 ? launch_gemm_pair<EPI_STORE, KIND_F16F8>(ctx, Ac, rA, pCT8, dAB, nK, ga, s, map_rows, slices)
                       : launch_gemm_pair<EPI_STORE, KIND_BF16>(ctx, Ac, rA, pCT, dAB, nK, ga, s, map_rows, slices);
      else e = f8 ? launch_gemm<EPI_STORE, KIND_F16F8>(ctx, Ac, rA, pCT8, dAB, nK, ga, s, map_rows, 0, slices)
                  : launch_gemm<EPI_STORE, KIND_BF16>(ctx, Ac, rA, pCT, dAB, nK, ga, s, map_rows, 0, slices);
      if (e) return e;
      if (slices > 1) {
        k_sum_slices_f<<<grid_for((size_t)rc * d, 256), 256, 0, s>>>((const float*)pS0P, slices, (size_t)rc * d, S0 + (size_t)r0 * d);
        HX_LAUNCH_CHECK("k_sum_slices_f");
        ctx->launches += 1;
      }
      if (draw_next && fork_at > c && fork_at <= cl + 1 && fork_at > 0) {
        if (int e2 = job_launch(ctx, 1 - cur, Pbuf[1 - cur], f8, impl, n, next->row0, next->rows, rA, nK, chunk, next->key, next->key_ptr, s))
          return e2;
      }
    }
    ga.kb_per_slice = 0; ga.slice_stride = 0; ga.one_acc = 0;
    if (int e = timing_mark(ctx, s, HX_PH_LEAP)) return e;
  }

  // (4) leapfrog + energy products
  if (prop) {
    if (e_f16)
      k_z0_planes_f16<<<(rA + 7) / 8, 256, 0, s>>>(X1, mu, S0, rows, d, rA, dK, kf16 ? nullptr : (bf*)pZB, (__half*)pZT, (float*)pRS);
    else
      k_z0_planes<<<grid_for((size_t)rA * K2 / 4, 256), 256, 0, s>>>(X1, mu, S0, rows, d, rA, dK, g1_tf32 ? nullptr : (bf*)pZB,
                                                                     e_bf16 ? nullptr : (float*)pZT);
    HX_LAUNCH_CHECK("k_z0_planes");
    ctx->launches += 1;
    GemmArgs g1;
    memset(&g1, 0, sizeof(g1));
    g1.nprod = 3; g1.M = rows; g1.N = d; g1.mu = mu; g1.Q = Q; g1.S = S; g1.blk_cols = dB;
    g1.planes_out = pQC; g1.rows_out_pad = rQ; g1.Kout_p = dK; g1.planes_bf16 = ekind;
    g1.row_scale = (const float*)pRS;
    g1.amax = (const uint32_t*)pAM; g1.unscale_idx = 3;
    if (kf16) {
      if (int e = launch_gemm_pair<EPI_PROP, KIND_F16>(ctx, pZT, rA, pB1, 2 * dB, K2, g1, s, 0, 1, 2 * dB)) return e;
    } else if (g1_tf32) {
      if (int e = launch_gemm<EPI_PROP, KIND_TF32>(ctx, pZT, rA, pB1, 2 * dB, K2, g1, s, 0, 2 * dB)) return e;
    } else if (ctx->aux_pair) {
      if (int e = launch_gemm_pair<EPI_PROP, KIND_BF16>(ctx, pZB, rA, pB1, 2 * dB, K2, g1, s, 0, 1, 2 * dB)) return e;
    } else {
      if (int e = launch_gemm<EPI_PROP, KIND_BF16>(ctx, pZB, rA, pB1, 2 * dB, K2, g1, s, 0, 2 * dB)) return e;
    }
    GemmArgs g2;
    memset(&g2, 0, sizeof(g2));
    g2.nprod = 3; g2.M = rows; g2.N = d; g2.NT = NT; g2.X = S0; g2.Y = S; g2.part = (float*)pDKP;
    if (e_f16) {
      g2.amax = (const uint32_t*)pAM; g2.row_scale = (const float*)pRS; g2.dot_pair = 1;
      if (ctx->aux_pair == 2) {
        if (int e = launch_gemm_pair<EPI_DOT, KIND_F16>(ctx, pZT, rA, pB2, dB, K2, g2, s)) return e;
      } else if (int e = launch_gemm<EPI_DOT, KIND_F16>(ctx, pZT, rA, pB2, dB, K2, g2, s)) return e;
    } else if (e_bf16) {
      if (int e = launch_gemm<EPI_DOT, KIND_BF16>(ctx, pZB, rA, pB2, dB, K2, g2, s)) return e;
    } else {
      if (int e = launch_gemm<EPI_DOT, KIND_TF32>(ctx, pZT, rA, pB2, dB, K2, g2, s)) return e;
    }
    GemmArgs g3;
    memset(&g3, 0, sizeof(g3));
    g3.nprod = 3; g3.M = rows; g3.N = d; g3.NT = NT; g3.mu = mu; g3.X = Q; g3.part = (float*)pLPP;
    if (e_f16) {
      g3.amax = (const uint32_t*)pAM; g3.row_scale = (const float*)pRS; g3.dot_pair = 2;
      if (ctx->aux_pair == 2) {
        if (int e = launch_gemm_pair<EPI_DOT, KIND_F16>(ctx, pQC, rQ, pPP, dB, dK, g3, s)) return e;
      } else if (int e = launch_gemm<EPI_DOT, KIND_F16>(ctx, pQC, rQ, pPP, dB, dK, g3, s)) return e;       // pPP: fp16 planes of P
    } else if (e_bf16) {
      if (int e = launch_gemm<EPI_DOT, KIND_BF16>(ctx, pQC, rQ, pPP, dB, dK, g3, s)) return e;      // pPP: bf16 planes from the basis run
    } else {
      if (int e = launch_gemm<EPI_DOT, KIND_TF32>(ctx, pQC, rQ, pPP32, dB, dK, g3, s)) return e;
    }
  } else {
    LeapArgs la;
    la.X1 = X1; la.mu = mu; la.rows = rows; la.d = d; la.rA = rA; la.dK = dK; la.dB = dB; la.NT = NT; la.L = L; la.eps = eps;
    la.Q = Q; la.S = S; la.S0 = S0; la.Qs = Qs; la.Bo = (const float*)pBo; la.prec = (const float*)tgt->precision;
    la.ld_prec = tgt->ld; la.pBoP = pBoP; la.pPP = pPP; la.pQC = pQC; la.pQC2 = pQC2; la.pQsP = pQsP; la.lpp = (float*)pLPP; la.dkp = (float*)pDKP;
    la.energies = 1; la.am = (uint32_t*)pAM; la.rs = (float*)pRS;
    const int e = g1_tf32 ? leapfrog_bform<KIND_TF32>(ctx, la, s)
                          : (kf16 ? leapfrog_bform<KIND_F16>(ctx, la, s) : leapfrog_bform<KIND_BF16>(ctx, la, s));
    if (e) return e;
  }

  // (5) energies, log-accept, optional fused accept + chain emit
  if (int e = timing_mark(ctx, s, HX_PH_FINISH)) return e;
  FinishArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.lpp = (const float*)pLPP; fa.dkp = (const float*)pDKP; fa.NT = NT; fa.lp1 = lp1; fa.Q = Q; fa.logacc = logacc;
  fa.lp_out = lp_out; fa.rows = rows; fa.d = d; fa.n = n; fa.row0 = row0; fa.has_lower0 = tgt->has_lower0;
  fa.lower0 = (float)tgt->lower0; fa.legacy = impl == HX_RNG_LEGACY ? 1 : 0;   // tensor-core path: fast normals by definition
  if (fuse) {
    fa.fuse = 1; fa.akey_ptr = fuse->akey_ptr; fa.X1w = const_cast<float*>(X1); fa.lp1w = fuse->lp1w;
    fa.accepts = fuse->accepts; fa.chain_out = fuse->chain_out; fa.chain_lp = fuse->chain_lp; fa.nonfinite = fuse->nonfinite;
    if (fuse->push) { fa.push = 1; fa.dp = *fuse->push; }
  }
  k_tc_finish<<<(rows + 7) / 8, 256, 0, s>>>(fa);
  HX_LAUNCH_CHECK("k_tc_finish");
  ctx->launches += 2;
  return timing_mark(ctx, s, -1);
}

// ---- diagnostics: how well do the momentum draw (ALU) and the projection GEMM (tensor pipe) overlap? --------------------------
// mode 0: `reps` draws alone; 1: `reps` projection GEMMs alone; 2: both at once (draw on the side stream into one buffer, GEMM on
// `s` from the other).  Returns the wall time of the whole batch in ms (CUDA events spanning both streams).
int microbench(hx_ctx* ctx, int mode, int n, int rows, int d, int reps, double* ms_out, cudaStream_t s) {
  const int dB = pad_to(d, BN), nK = pad_to(n, 64), rA = pad_to(rows, kBM), dA = pad_to(d, kBM);
  const int dAB = dB > dA ? dB : dA, NT = (d + BN - 1) / BN;
  typedef __nv_bfloat16 bf;
  void *pCT, *pP0, *pS0, *pSc8;
  const int f8 = (ctx->precision == HX_PREC_F16F8 || ctx->precision == HX_PREC_AUTO) ? 1 : 0;
  const size_t buf_elems = (size_t)2 * rA * nK;
  if (int e = ws_get(ctx, WS_TC_CT, (size_t)2 * dAB * nK * sizeof(bf), &pCT)) return e;
  if (int e = ws_get(ctx, WS_TC_P0, 2 * buf_elems * sizeof(bf), &pP0)) return e;
  if (int e = ws_get(ctx, WS_TC_S0, (size_t)rows * d * sizeof(float), &pS0)) return e;
  if (int e = ws_get(ctx, WS_TC_SC8, 16, &pSc8)) return e;
  if (int e = side_init(ctx)) return e;
  ctx->job[0].valid = ctx->job[1].valid = 0;
  ctx->p0_base = nullptr;
  HX_CUDA(cudaMemsetAsync(pCT, 0x3c, (size_t)2 * dAB * nK * sizeof(bf), s));     // finite bf16 / fp16 / e4m3 pattern
  HX_CUDA(cudaMemsetAsync(pSc8, 0x3c, 16, s));
  if (mode == 3) {
    Key key{1u, 2u};
    cudaEvent_t e0, e1;
    HX_CUDA(cudaEventCreate(&e0)); HX_CUDA(cudaEventCreate(&e1));
    HX_CUDA(cudaEventRecord(e0, s));
    for (int r = 0; r < reps; ++r)
      if (int e = launch_proj_fused(ctx, pCT, dAB, nK, n, 0, rows, d, key, nullptr, (float*)pS0, s,
                                    fused_format(ctx) == 2 ? (const float*)pSc8 : nullptr)) return e;
    HX_CUDA(cudaEventRecord(e1, s));
    HX_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    HX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    ms_out[0] = ms_out[1] = ms_out[2] = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
  }
  bf* Pbuf[2] = {(bf*)pP0, (bf*)pP0 + buf_elems};
  const int chunk = chunk_rows(ctx, rA, NT);
  const int gchunk = ((ctx->proj_pair && !ctx->pair_chunked) || (!ctx->proj_pair && ctx->proj_one_launch)) ? rA : chunk;        // pair kernel: one launch over all rows (steady state of a batch loop)
  Key key{1u, 2u};
  if (int e = job_launch(ctx, 0, Pbuf[0], f8, 0, n, 0, rows, rA, nK, chunk, key, nullptr, s)) return e;   // valid operand for the GEMM
  HX_CUDA(cudaStreamSynchronize(ctx->side));
  cudaEvent_t e0, e1, eside, eg;
  HX_CUDA(cudaEventCreate(&e0)); HX_CUDA(cudaEventCreate(&e1)); HX_CUDA(cudaEventCreate(&eside)); HX_CUDA(cudaEventCreate(&eg));
  HX_CUDA(cudaEventRecord(e0, s));
  for (int r = 0; r < reps; ++r) {
    if (mode == 0 || mode == 2)
      if (int e = job_launch(ctx, 1, Pbuf[1], f8, 0, n, 0, rows, rA, nK, chunk, key, nullptr, s)) return e;
    if (mode == 1 || mode == 2) {
      GemmArgs ga;
      memset(&ga, 0, sizeof(ga));
      ga.nprod = 3;
      const int nchunks = (rows + gchunk - 1) / gchunk;
      for (int c = 0; c < nchunks; ++c) {
        const int r0 = c * gchunk, rc = (rows - r0) < gchunk ? (rows - r0) : gchunk;
        ga.M = rc; ga.N = d; ga.out = (float*)pS0 + (size_t)r0 * d; ga.ldo = d;
        ga.scale_ptr = (const float*)pSc8;
          const uint64_t mr = (uint64_t)rA + pad_to(rc, kBM);
        const bf* Ac = Pbuf[0] + (size_t)r0 * nK;
        int e;
        if (ctx->proj_pair) e = f8 ? launch_gemm_pair<EPI_STORE, KIND_F16F8>(ctx, Ac, rA, pCT, dAB, nK, ga, s, mr)
                                   : launch_gemm_pair<EPI_STORE, KIND_BF16>(ctx, Ac, rA, pCT, dAB, nK, ga, s, mr);
        else e = f8 ? launch_gemm<EPI_STORE, KIND_F16F8>(ctx, Ac, rA, pCT, dAB, nK, ga, s, mr)
                    : launch_gemm<EPI_STORE, KIND_BF16>(ctx, Ac, rA, pCT, dAB, nK, ga, s, mr);
        if (e) return e;
      }
    }
  }
  HX_CUDA(cudaEventRecord(eg, s));
  HX_CUDA(cudaEventRecord(eside, ctx->side));
  HX_CUDA(cudaStreamWaitEvent(s, eside, 0));
  HX_CUDA(cudaEventRecord(e1, s));
  HX_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f, msg = 0.f, msr = 0.f;
  HX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  HX_CUDA(cudaEventElapsedTime(&msg, e0, eg));
  HX_CUDA(cudaEventElapsedTime(&msr, e0, eside));
  ms_out[0] = ms; ms_out[1] = msg; ms_out[2] = msr;      // total, GEMM stream done, draw stream done
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(eside); cudaEventDestroy(eg);
  ctx->job[0].valid = ctx->job[1].valid = 0;
  return 0;
}

}  // namespace tc
}  // namespace hx
