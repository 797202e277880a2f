// hx_tc_walk.cu — tensor-core path of the walk half-move for full-covariance Gaussian targets at large
// dim / walker counts (fp32 state, split-bf16 operands on tcgen05, fp32 accumulation in TMEM).
//
// Same projected-form algorithm as the exact SIMT kernel (hx_move.cuh; reference
// moves/hamiltonian/hmc_walk.py:6-107), expressed as GEMMs:
//   C^T planes    <- ((X2 - mean)/sqrt(n))^T                    (hmc_walk.py:35)
//   M             <- C^T C                                       tc_gemm(CT, CT)
//   P0 planes     <- normal(key, (n, n))[row shard]              (hmc_walk.py:36), counter-based, bf16 hi/lo
//   S0            <- P0 C                                        tc_gemm(P0, CT)
//   per kick:  g  <- (q - mu) Prec                               tc_gemm(QC, PrecPlanes)  epilogue GRAD
//              S  <- S - a g M ; dK += a<g, S- + S+> ; q += eps S   tc_gemm(G, MPlanes)   epilogue KICK
//   finish:    dH = (lp1 - lp') - dK/2 ; log_acc ; optional Metropolis accept + chain emit
#include <cstring>
#include <cstdio>
#include <cstdlib>

#include "../../include/hx.h"
#include "hx_ctx.h"
#include "hx_rng.cuh"
#include "hx_common.cuh"
#include "hx_tc_gemm.cuh"
#include "hx_tc_walk.h"

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

// bf16 planes array [rows_total, Kp], K-major; box = [64 k-elements (128 B), box_rows], 128B swizzle
static int make_tmap(hx_ctx* ctx, CUtensorMap* out, const void* base, uint64_t rows_total, uint64_t Kp, uint32_t box_rows) {
  EncodeFn enc;
  if (int e = get_encode(ctx, &enc)) return e;
  cuuint64_t dims[2] = {Kp, rows_total};
  cuuint64_t strides[1] = {Kp * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return hx_fail(HX_E_UNSUPPORTED, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// ---- elementwise / layout kernels -----------------------------------------------------------------------------
// CT planes [2, dB_pad, Kp]: CT[c][j] = (X2[j][c] - mean[c]) / sqrt(n); zero outside (c < d, j < n)
__global__ void k_center_planes_T(const float* __restrict__ X2, const float* __restrict__ mean, int n, int d,
                                  int dB_pad, int Kp, __nv_bfloat16* __restrict__ planes) {
  __shared__ float tile[32][33];
  const int j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float sc = sqrtf((float)n);
  for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
    const int j = j0 + jj, c = c0 + threadIdx.x;
    tile[jj][threadIdx.x] = (j < n && c < d) ? (X2[(size_t)j * d + c] - mean[c]) / sc : 0.f;
  }
  __syncthreads();
  const size_t plane = (size_t)dB_pad * Kp;
  for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
    const int c = c0 + cc, j = j0 + threadIdx.x;
    if (c < dB_pad && j < Kp) {
      __nv_bfloat16 hi, lo;
      split_bf16(tile[threadIdx.x][cc], hi, lo);
      planes[(size_t)c * Kp + j] = hi;
      planes[plane + (size_t)c * Kp + j] = lo;
    }
  }
}

// planes [2, rows_pad, Kp] <- split(src[r*ld + k]) for r, k < dim, zero elsewhere
__global__ void k_split_planes(const float* __restrict__ src, int dim_r, int dim_k, int ld, int rows_pad, int Kp,
                               __nv_bfloat16* __restrict__ planes) {
  const size_t tot = (size_t)rows_pad * Kp, plane = tot;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / Kp), k = (int)(i - (size_t)r * Kp);
    const float v = (r < dim_r && k < dim_k) ? src[(size_t)r * ld + k] : 0.f;
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    planes[i] = hi;
    planes[plane + i] = lo;
  }
}

// P0 planes [2, rows_pad, Kp]: P0[r][j] = normal(key, (n, n))[row0 + r, j]; 8 consecutive j per thread
__global__ void __launch_bounds__(256) k_rng_planes(Key key, const uint32_t* key_ptr, int legacy, int n, int row0, int rows,
                                                    int rows_pad, int Kp, __nv_bfloat16* __restrict__ planes) {
  if (key_ptr) { key.k0 = key_ptr[0]; key.k1 = key_ptr[1]; }
  const size_t plane = (size_t)rows_pad * Kp;
  const size_t nvec = plane / 8;
  const uint64_t nn = (uint64_t)n * (uint64_t)n;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
    const size_t i = v * 8;
    const int r = (int)(i / Kp), j0 = (int)(i - (size_t)r * Kp);
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int j = j0 + t;
      float x = 0.f;
      if (r < rows && j < n) x = normal_rt<float>(key, (uint64_t)(row0 + r) * (uint64_t)n + (uint64_t)j, nn, legacy != 0);
      split_bf16(x, hi[t], lo[t]);
    }
    *reinterpret_cast<uint4*>(planes + i) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(planes + plane + i) = *reinterpret_cast<const uint4*>(lo);
  }
}

// Q <- X1 ; QC planes <- split(X1 - mu) (zero padded) ; dK partials <- 0
__global__ void k_tc_init(const float* __restrict__ X1, const float* __restrict__ mu, int rows, int d, int rows_pad, int Kp,
                          float* __restrict__ Q, __nv_bfloat16* __restrict__ planes, float* __restrict__ dkp, int NT) {
  const size_t tot = (size_t)rows_pad * Kp, plane = tot;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / Kp), k = (int)(i - (size_t)r * Kp);
    float v = 0.f;
    if (r < rows && k < d) {
      const float x = X1[(size_t)r * d + k];
      Q[(size_t)r * d + k] = x;
      v = x - (mu ? mu[k] : 0.f);
    }
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    planes[i] = hi;
    planes[plane + i] = lo;
    if (k < NT && r < rows) dkp[(size_t)r * NT + k] = 0.f;
  }
}

struct FinishArgs {
  const float* lpp; const float* dkp; int NT;
  const float* lp1; const float* Q;
  float* logacc; float* lp_out;
  int rows, d, n, row0, has_lower0; float lower0;
  int fuse, legacy; Key akey; const uint32_t* akey_ptr;
  float* X1w; float* lp1w; int32_t* accepts; float* chain_out; float* chain_lp; int32_t* nonfinite;
};

// one warp per walker: energies, log-accept and (optionally) the Metropolis accept + chain emit
__global__ void __launch_bounds__(256) k_tc_finish(FinishArgs a) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= a.rows) return;
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
  if (!a.fuse) return;
  Key akey = a.akey;
  if (a.akey_ptr) { akey.k0 = a.akey_ptr[0]; akey.k1 = a.akey_ptr[1]; }
  const float lo = 1e-10f, span = 1.0f - lo;
  const float u = uniform_rt<float>(akey, (uint64_t)(a.row0 + r), (uint64_t)a.n, lo, span, a.legacy != 0);
  const bool acc = logf(u) < la;                           // sampler_utils.py:114-115
  const float lps = acc ? lp : lp1;
  if (lane == 0) {
    a.lp1w[r] = lps;
    if (a.chain_lp) a.chain_lp[r] = lps;
    a.accepts[r] += acc ? 1 : 0;
    if (!isfinite(lps) && a.nonfinite) atomicOr(a.nonfinite, 1);
  }
  const size_t off = (size_t)r * a.d;
  for (int c = lane; c < a.d; c += 32) {
    const float v = acc ? a.Q[off + c] : a.X1w[off + c];
    a.X1w[off + c] = v;
    if (a.chain_out) a.chain_out[off + c] = v;
  }
}

// ---- GEMM launcher ---------------------------------------------------------------------------------------------------
template <int EPI>
static int launch_gemm(hx_ctx* ctx, const __nv_bfloat16* A, int rowsA_pad, const __nv_bfloat16* B, int rowsB_pad, int Kp,
                       GemmArgs g, cudaStream_t s) {
  CUtensorMap tmA, tmB;
  if (int e = make_tmap(ctx, &tmA, A, 2ull * rowsA_pad, Kp, kBM)) return e;
  if (int e = make_tmap(ctx, &tmB, B, 2ull * rowsB_pad, Kp, BN)) return e;
  g.Kp = Kp; g.rowsA_pad = rowsA_pad; g.rowsB_pad = rowsB_pad;
  static bool attr_done[3] = {false, false, false};
  if (!attr_done[EPI]) {
    HX_CUDA(cudaFuncSetAttribute(k_tc_gemm<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<BN>::BYTES));
    attr_done[EPI] = true;
  }
  dim3 grid((g.N + BN - 1) / BN, (g.M + kBM - 1) / kBM);
  k_tc_gemm<BN, EPI><<<grid, kThreads, SmemLayout<BN>::BYTES, s>>>(tmA, tmB, g);
  HX_LAUNCH_CHECK("k_tc_gemm");
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

bool walk_eligible(const hx_ctx* ctx, int dtype, const hx_target* tgt, int n, int rows, int flags) {
  if (ctx->precision == HX_PREC_EXACT) return false;
  if (dtype != HX_F32 || tgt->kind != HX_TARGET_GAUSS_FULL || flags != 0) return false;
  if (ctx->precision == HX_PREC_AUTO) return tgt->dim >= 256 && n >= 2048 && rows >= 128;
  return tgt->dim >= 16 && n >= 64;      // forced tensor-core modes (tests)
}

int walk_half(hx_ctx* ctx, int impl, const hx_target* tgt, const float* mean_dev, const float* X1, const float* X2, int n,
              int row0, int rows, double eps, Key key, const uint32_t* key_ptr, int L, const float* lp1, float* Q,
              float* logacc, float* S, float* lp_out, const FuseF32* fuse, cudaStream_t s) {
  const int d = tgt->dim;
  const int nprod = (ctx->precision == HX_PREC_BF16) ? 1 : 3;
  const int dB = pad_to(d, BN);            // rows of B-operand planes (N side)
  const int dK = pad_to(d, 64);            // K padding when dim is the reduction axis
  const int nK = pad_to(n, 64);            // K padding when the walker index is the reduction axis
  const int rA = pad_to(rows, kBM);        // rows of A-operand planes over walkers
  const int dA = pad_to(d, kBM);           // rows of A-operand planes over dim (Gram)
  const int dAB = dB > dA ? dB : dA;       // CT planes serve as both A (128-row boxes) and B (BN-row boxes)
  const int NT = (d + BN - 1) / BN;
  typedef __nv_bfloat16 bf;
  void *pCT, *pM, *pMP, *pPP, *pP0, *pQC, *pGP, *pG, *pLPP, *pDKP;
  if (int e = ws_get(ctx, WS_TC_CT, (size_t)2 * dAB * nK * sizeof(bf), &pCT)) return e;
  if (int e = ws_get(ctx, WS_M, (size_t)d * d * sizeof(float), &pM)) return e;
  if (int e = ws_get(ctx, WS_TC_MP, (size_t)2 * dB * dK * sizeof(bf), &pMP)) return e;
  if (int e = ws_get(ctx, WS_TC_PP, (size_t)2 * dB * dK * sizeof(bf), &pPP)) return e;
  // momentum-draw pipeline: chunks of >= 1024 rows (multiple of 128), ~8 per shard, two buffers in flight
  int chunk = pad_to((rows + 7) / 8, kBM);
  if (chunk < 1024) chunk = pad_to(rows < 1024 ? rows : 1024, kBM);
  const int nchunks = (rows + chunk - 1) / chunk;
  const size_t buf_elems = (size_t)2 * chunk * nK;            // one buffer = [2 planes, chunk, nK]
  if (int e = ws_get(ctx, WS_TC_P0, (nchunks > 1 ? 2 : 1) * buf_elems * sizeof(bf), &pP0)) return e;
  if (int e = ws_get(ctx, WS_TC_QC, (size_t)2 * rA * dK * sizeof(bf), &pQC)) return e;
  if (int e = ws_get(ctx, WS_TC_GP, (size_t)2 * rA * dK * sizeof(bf), &pGP)) return e;
  if (int e = ws_get(ctx, WS_G, (size_t)rows * d * sizeof(float), &pG)) return e;
  if (int e = ws_get(ctx, WS_TC_LPP, (size_t)rows * NT * sizeof(float), &pLPP)) return e;
  if (int e = ws_get(ctx, WS_TC_DKP, (size_t)rows * NT * sizeof(float), &pDKP)) return e;

  // (1) complement statistics: mean_dev already holds mean_0(X2) (hx_api.cu); C^T planes, M = C^T C
  {
    dim3 g((nK + 31) / 32, (dAB + 31) / 32), b(32, 8);
    k_center_planes_T<<<g, b, 0, s>>>(X2, mean_dev, n, d, dAB, nK, (bf*)pCT);
    HX_LAUNCH_CHECK("k_center_planes_T");
  }
  GemmArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.nprod = nprod;
  ga.M = d; ga.N = d; ga.out = (float*)pM; ga.ldo = d;
  if (int e = launch_gemm<EPI_STORE>(ctx, (const bf*)pCT, dAB, (const bf*)pCT, dAB, nK, ga, s)) return e;
  k_split_planes<<<grid_for((size_t)dB * dK, 256), 256, 0, s>>>((const float*)pM, d, d, d, dB, dK, (bf*)pMP);
  HX_LAUNCH_CHECK("k_split_planes(M)");
  k_split_planes<<<grid_for((size_t)dB * dK, 256), 256, 0, s>>>((const float*)tgt->precision, d, d, tgt->ld, dB, dK, (bf*)pPP);
  HX_LAUNCH_CHECK("k_split_planes(P)");

  // (2) momentum draw (hmc_walk.py:36) and projection S0 = P0 C, pipelined over row chunks: the
  //     counter-based draw of chunk c+1 (ALU pipes, side stream) overlaps the tcgen05 GEMM of chunk c.
  {
    if (!ctx->side_ready) {
      int prio_lo = 0, prio_hi = 0;
      HX_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      HX_CUDA(cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_lo));
      HX_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
      for (int i = 0; i < 2; ++i) {
        HX_CUDA(cudaEventCreateWithFlags(&ctx->ev_rng[i], cudaEventDisableTiming));
        HX_CUDA(cudaEventCreateWithFlags(&ctx->ev_gemm[i], cudaEventDisableTiming));
      }
      // same L1/shared carve-out as the GEMM kernel, otherwise the two kernels cannot share an SM
      HX_CUDA(cudaFuncSetAttribute(k_rng_planes, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      ctx->side_ready = 1;
    }
    bf* Pbuf[2] = {(bf*)pP0, (bf*)pP0 + (nchunks > 1 ? buf_elems : 0)};
    const bool trace = getenv("HX_TRACE") != nullptr;
    cudaEvent_t tev[4][64];
    if (trace) for (int a = 0; a < 4; ++a) for (int c = 0; c < nchunks && c < 64; ++c) cudaEventCreate(&tev[a][c]);
    HX_CUDA(cudaEventRecord(ctx->ev_fork, s));
    HX_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    if (int e = timing_begin(ctx, s)) return e;
    for (int c = 0; c < nchunks; ++c) {
      const int r0 = c * chunk;
      const int rc = (rows - r0) < chunk ? (rows - r0) : chunk;
      const int b = c & 1;
      if (c >= 2) HX_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_gemm[b], 0));     // buffer b free again
      // at most 4 CTAs (1024 threads) per SM so a GEMM CTA (192 threads, ~215 KB smem) always fits beside them
      int rng_grid = grid_for((size_t)chunk * nK / 8, 256);
      { const char* e_ = getenv("HX_RNG_CTAS"); const int per = e_ ? atoi(e_) : 4; if (rng_grid > ctx->sm_count * per) rng_grid = ctx->sm_count * per; }
      if (trace && c < 64) cudaEventRecord(tev[0][c], ctx->side);
      k_rng_planes<<<rng_grid, 256, 0, ctx->side>>>(
          key, key_ptr, impl == HX_RNG_LEGACY ? 1 : 0, n, row0 + r0, rc, chunk, nK, Pbuf[b]);
      HX_LAUNCH_CHECK("k_rng_planes");
      if (trace && c < 64) cudaEventRecord(tev[1][c], ctx->side);
      HX_CUDA(cudaEventRecord(ctx->ev_rng[b], ctx->side));
      HX_CUDA(cudaStreamWaitEvent(s, ctx->ev_rng[b], 0));
      if (trace && c < 64) cudaEventRecord(tev[2][c], s);
      ga.M = rc; ga.N = d; ga.out = S + (size_t)r0 * d; ga.ldo = d;
      if (int e = launch_gemm<EPI_STORE>(ctx, Pbuf[b], chunk, (const bf*)pCT, dAB, nK, ga, s)) return e;
      HX_CUDA(cudaEventRecord(ctx->ev_gemm[b], s));
      if (trace && c < 64) cudaEventRecord(tev[3][c], s);
      ctx->launches += 1;
    }
    if (trace) {
      cudaStreamSynchronize(s); cudaStreamSynchronize(ctx->side);
      for (int c = 0; c < nchunks && c < 64; ++c) {
        float t[4];
        for (int a = 0; a < 4; ++a) cudaEventElapsedTime(&t[a], tev[0][0], tev[a][c]);
        fprintf(stderr, "[hx trace] chunk %d rng %.3f-%.3f gemm %.3f-%.3f ms\n", c, t[0], t[1], t[2], t[3]);
      }
      for (int a = 0; a < 4; ++a) for (int c = 0; c < nchunks && c < 64; ++c) cudaEventDestroy(tev[a][c]);
    }
    if (int e = timing_end(ctx, s)) return e;
  }

  // (3) leapfrog: L+1 kicks, every kick but the last followed by a drift (hmc_walk.py:85-102)
  const float* mu = (const float*)tgt->mean;
  k_tc_init<<<grid_for((size_t)rA * dK, 256), 256, 0, s>>>(X1, mu, rows, d, rA, dK, Q, (bf*)pQC, (float*)pDKP, NT);
  HX_LAUNCH_CHECK("k_tc_init");
  HX_CUDA(cudaMemsetAsync(pGP, 0, (size_t)2 * rA * dK * sizeof(bf), s));
  GemmArgs gg;
  memset(&gg, 0, sizeof(gg));
  gg.nprod = nprod; gg.M = rows; gg.N = d; gg.S = S; gg.Q = Q; gg.G = (float*)pG; gg.mu = mu;
  gg.rows_out_pad = rA; gg.Kout_p = dK; gg.NT = NT; gg.eps = (float)eps;
  auto grad = [&](int want_lp) -> int {
    GemmArgs x = gg;
    x.planes_out = (bf*)pGP; x.part = (float*)pLPP; x.want_lp = want_lp;
    return launch_gemm<EPI_GRAD>(ctx, (const bf*)pQC, rA, (const bf*)pPP, dB, dK, x, s);
  };
  if (int e = grad(0)) return e;
  for (int k = 0; k <= L; ++k) {
    GemmArgs x = gg;
    x.planes_out = (bf*)pQC; x.part = (float*)pDKP;
    x.a = (float)((k == 0 || k == L) ? 0.5 * (float)eps : (float)eps);
    x.drift = k < L ? 1 : 0;
    if (int e = launch_gemm<EPI_KICK>(ctx, (const bf*)pGP, rA, (const bf*)pMP, dB, dK, x, s)) return e;
    if (k < L) { if (int e = grad(k == L - 1 ? 1 : 0)) return e; }
  }

  // (4) energies, log-accept, optional fused accept + chain emit
  FinishArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.lpp = (const float*)pLPP; fa.dkp = (const float*)pDKP; fa.NT = NT; fa.lp1 = lp1; fa.Q = Q; fa.logacc = logacc;
  fa.lp_out = lp_out; fa.rows = rows; fa.d = d; fa.n = n; fa.row0 = row0; fa.has_lower0 = tgt->has_lower0;
  fa.lower0 = (float)tgt->lower0; fa.legacy = impl == HX_RNG_LEGACY ? 1 : 0;
  if (fuse) {
    fa.fuse = 1; fa.akey_ptr = fuse->akey_ptr; fa.X1w = const_cast<float*>(X1); fa.lp1w = fuse->lp1w;
    fa.accepts = fuse->accepts; fa.chain_out = fuse->chain_out; fa.chain_lp = fuse->chain_lp; fa.nonfinite = fuse->nonfinite;
  }
  k_tc_finish<<<(rows + 7) / 8, 256, 0, s>>>(fa);
  HX_LAUNCH_CHECK("k_tc_finish");
  ctx->launches += 6;
  return 0;
}

}  // namespace tc
}  // namespace hx
