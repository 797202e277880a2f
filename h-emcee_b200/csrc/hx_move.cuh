// hx_move.cuh — fused half-move kernels (exact SIMT path, f32 / f64).
//
// One CTA owns BM walkers of the group being updated and carries them through the whole
// half-move: momentum draw + projection, L leapfrog steps (target gradient, kick, drift), final
// log-density, energy difference and (optionally) the Metropolis accept.  Walkers of a group are
// independent given the complement, so no inter-CTA synchronisation is needed.
//
// Walk move in PROJECTED form.  The reference integrates the
// n x n momentum p (moves/hamiltonian/hmc_walk.py:61-107); every use of p is through
// S = p @ C (n x d) and the kick `p -= a * g @ C^T` becomes `S -= a * g @ M`, M = C^T C.  The
// kinetic-energy change of one kick is  -a/2 * <g, S_before + S_after>  per walker, so
//   dH = (U1 - U0) - 1/2 * sum_kicks a_k <g_k, S_k^- + S_k^+>
// replaces hmc_walk.py:47-54 without ever forming p.  P0 = normal(key,(n,n)) (hmc_walk.py:36) is
// generated tile-wise in registers from the counter-based RNG and consumed immediately.
#pragma once
#include "hx_common.cuh"
#include "hx_rng.cuh"

namespace hx {

enum : int {
  kFlagFrozen = 1,      // HX_MOVE_FROZEN_GRAD
  kFlagGivenMomentum = 2,  // integrator-only entry: S (walk) / p,D (side) supplied by the caller
  kFlagInPlaceQ = 4,    // q already sits in Q (no copy from X1)
};

constexpr int kBK = 16;

template <typename T, int BM, int TM>
struct TileMap {
  static constexpr int RG = BM / TM;
  static constexpr int LDA = BM + 4;
  int tid, nthreads, CG, CGP, cg, rg;
  bool active;
  __device__ __forceinline__ void init(int ldd) {
    CG = ldd >> 2;
    CGP = (CG >= kWarp) ? ((CG + kWarp - 1) / kWarp) * kWarp : CG;
    tid = threadIdx.x;
    nthreads = blockDim.x;
    rg = tid / CGP;
    cg = tid - rg * CGP;
    active = (rg < RG) && (cg < CG);
  }
};

template <int BM, int TM>
inline int tile_threads(int ldd) {
  const int CG = ldd / 4;
  const int CGP = (CG >= 32) ? ((CG + 31) / 32) * 32 : CG;
  int t = (BM / TM) * CGP;
  t = ((t + 31) / 32) * 32;
  if (t < 128) t = 128;
  return t;
}

// ---- A-operand loaders: fill As[kBK][LDA] (k-major, rows contiguous) for chunk k0 ---------------
template <typename T, int BM>
struct RngLoader {
  Key key;
  int legacy;
  uint64_t n;        // group size (P0 is n x n)
  int grow0;         // global row of the tile's first walker
  int rows_valid;
  __device__ __forceinline__ void operator()(T* As, int k0, int tid, int nthreads) const {
    constexpr int LDA = BM + 4;
    for (int idx = tid; idx < BM * kBK; idx += nthreads) {
      const int r = idx % BM, kk = idx / BM;
      const uint64_t j = (uint64_t)(k0 + kk);
      T v = T(0);
      if (r < rows_valid && j < n) v = normal_rt<T>(key, (uint64_t)(grow0 + r) * n + j, n * n, legacy);
      As[kk * LDA + r] = v;
    }
  }
};

template <typename T, int BM>
struct RowsLoader {
  const T* src;  // tile rows, row-major, stride lds
  int lds;
  const T* mu;   // optional per-column shift
  int rows_valid, K;
  __device__ __forceinline__ void operator()(T* As, int k0, int tid, int nthreads) const {
    constexpr int LDA = BM + 4;
    for (int idx = tid; idx < BM * kBK; idx += nthreads) {
      const int kk = idx % kBK, r = idx / kBK;
      const int k = k0 + kk;
      T v = T(0);
      if (r < rows_valid && k < K) {
        v = src[(size_t)r * lds + k];
        if (mu) v -= mu[k];
      }
      As[kk * LDA + r] = v;
    }
  }
};

// acc[TM][4] = Atile[BM x K] @ B[K x ldb]  for this thread's (row group, 4 columns).
template <typename T, int BM, int TM, class Loader>
__device__ __forceinline__ void gemm_rows(T (&acc)[TM][4], const TileMap<T, BM, TM>& tm, T* As, const Loader& load,
                                          const T* __restrict__ B, int ldb, int K) {
  constexpr int LDA = BM + 4;
#pragma unroll
  for (int r = 0; r < TM; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = T(0); }
  // Small row tiles (TM <= 4) are latency-bound: the B rows of a K-chunk are fetched into registers BEFORE the A tile is
  // produced (for the momentum projection that is the RNG), so the L2 latency of the kBK independent loads overlaps the fill
  // instead of being paid 4 rows at a time inside the FMA loop.  Same FMA order => bit-identical results.
  constexpr int PRE = (TM <= 4 && BM >= 8) ? ((sizeof(T) == 4) ? kBK / 2 : kBK / 4) : 0;   // not the 1024-thread wide-dim tiles
  for (int k0 = 0; k0 < K; k0 += kBK) {
    const T* Bp = B + 4 * (tm.active ? tm.cg : 0);
    T breg[PRE > 0 ? PRE : 1][4];
    if constexpr (PRE > 0) {
      if (tm.active) {
#pragma unroll
        for (int kk = 0; kk < PRE; ++kk) {
          int k = k0 + kk;
          k = (k < K) ? k : (K - 1);
          ld4_ro(Bp + (size_t)k * ldb, breg[kk]);
        }
      }
    }
    __syncthreads();
    load(As, k0, tm.tid, tm.nthreads);
    __syncthreads();
    if (tm.active) {
      const T* Ap = As + tm.rg * TM;
      if constexpr (PRE > 0) {
#pragma unroll
        for (int h = 0; h < kBK / PRE; ++h) {
          if (h > 0) {
#pragma unroll
            for (int kk = 0; kk < PRE; ++kk) {
              int k = k0 + h * PRE + kk;
              k = (k < K) ? k : (K - 1);
              ld4_ro(Bp + (size_t)k * ldb, breg[kk]);
            }
          }
#pragma unroll
          for (int kk = 0; kk < PRE; ++kk) {
            const int ka = h * PRE + kk;
            if constexpr (TM % 4 == 0) {
#pragma unroll
              for (int r4 = 0; r4 < TM / 4; ++r4) {
                T a[4];
                ld4(Ap + ka * LDA + 4 * r4, a);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) acc[4 * r4 + i][j] = fma_t(a[i], breg[kk][j], acc[4 * r4 + i][j]);
                }
              }
            } else {
#pragma unroll
              for (int r = 0; r < TM; ++r) {
                const T a = Ap[ka * LDA + r];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[r][j] = fma_t(a, breg[kk][j], acc[r][j]);
              }
            }
          }
        }
      } else {
#pragma unroll 4
        for (int kk = 0; kk < kBK; ++kk) {
          int k = k0 + kk;
          k = (k < K) ? k : (K - 1);  // tail rows multiply zero-filled A
          T b[4];
          ld4_ro(Bp + (size_t)k * ldb, b);
#pragma unroll
          for (int r4 = 0; r4 < TM / 4; ++r4) {
            T a[4];
            ld4(Ap + kk * LDA + 4 * r4, a);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[4 * r4 + i][j] = fma_t(a[i], b[j], acc[4 * r4 + i][j]);
            }
          }
        }
      }
    }
  }
  __syncthreads();
}

// Deterministic per-row reduction of per-thread partials -> out[BM] (shared).  red: [BM*32] shared.
template <typename T, int BM, int TM>
__device__ __forceinline__ void row_reduce(const T (&part)[TM], const TileMap<T, BM, TM>& tm, T* red, T* out) {
  const int lane = tm.tid & 31;
  int RW;
  if (tm.CG >= kWarp) {
    RW = tm.CGP / kWarp;
    if (tm.rg < TileMap<T, BM, TM>::RG) {  // warp-uniform: CGP is a multiple of 32
      const int w = tm.cg / kWarp;
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        const T v = warp_sum(tm.active ? part[r] : T(0));
        if (lane == 0) red[(tm.rg * TM + r) * RW + w] = v;
      }
    }
  } else {
    RW = tm.CG;
    if (tm.active) {
#pragma unroll
      for (int r = 0; r < TM; ++r) red[(tm.rg * TM + r) * RW + tm.cg] = part[r];
    }
  }
  __syncthreads();
  if (tm.tid < BM) {
    T s = T(0);
    for (int w = 0; w < RW; ++w) s += red[tm.tid * RW + w];
    out[tm.tid] = s;
  }
  __syncthreads();
}

// ---- targets -----------------------------------------------------------------------------------
// Row-local targets (iso Gaussian, Rosenbrock, funnel): one warp per walker, lanes over dim.
// Writes gU = grad U = -grad log pi into G and log pi into lp_rows (shared, [BM]).
template <typename T>
__device__ __forceinline__ void target_rows_local(const TargetDev<T>& tg, const T* q, T* G, T* lp_rows, int rows_valid,
                                                  int tid, int nthreads) {
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  const int d = tg.dim;
  for (int r = warp; r < rows_valid; r += nwarps) {
    const T* x = q + (size_t)r * d;
    T* g = G + (size_t)r * d;
    T U = T(0);
    if (tg.kind == kGaussIso) {
      for (int i = lane; i < d; i += 32) {
        const T c = x[i] - (tg.mean ? tg.mean[i] : T(0));
        g[i] = c;
        U = fma_t(c, c, U);
      }
      U = T(0.5) * warp_sum(U);
    } else if (tg.kind == kRosenbrock) {
      const T a = tg.a, b = tg.b;
      for (int i = lane; i < d; i += 32) {
        const T xi = x[i];
        T gi = T(0);
        if (i + 1 < d) {
          const T ri = x[i + 1] - xi * xi;
          const T ai = a - xi;
          U += b * ri * ri + ai * ai;
          gi += T(-4) * b * ri * xi - T(2) * ai;
        }
        if (i > 0) {
          const T xm = x[i - 1];
          gi += T(2) * b * (xi - xm * xm);
        }
        g[i] = gi;
      }
      U = warp_sum(U);
    } else {  // kFunnel: x[0] = nu
      T s = T(0);
      for (int i = 1 + lane; i < d; i += 32) s = fma_t(x[i], x[i], s);
      s = warp_sum(s);
      const T nu = x[0];
      const T e = exp_t(-nu);
      const T m = T(d - 1);
      const T sig2 = tg.a * tg.a;
      U = nu * nu / (T(2) * sig2) + T(0.5) * e * s + T(0.5) * m * nu;
      for (int i = lane; i < d; i += 32) g[i] = (i == 0) ? (nu / sig2 - T(0.5) * e * s + T(0.5) * m) : e * x[i];
    }
    if (lane == 0) {
      T lp = -U;
      if (tg.has_lower0 && x[0] < tg.lower0) lp = -inf_v<T>();
      lp_rows[r] = lp;
    }
  }
  __syncthreads();
}

// Evaluate the target at the tile's positions q: G <- grad U, lp_rows[BM] <- log pi (if want_lp).
template <typename T, int BM, int TM>
__device__ __forceinline__ void eval_target(const TargetDev<T>& tg, const TileMap<T, BM, TM>& tm, const T* q, T* G,
                                            T* As, T* red, T* lp_rows, int rows_valid, bool want_lp) {
  const int d = tg.dim;
  if (tg.kind == kGaussFull) {
    T acc[TM][4];
    RowsLoader<T, BM> ld{q, d, tg.mean, rows_valid, d};
    gemm_rows<T, BM, TM>(acc, tm, As, ld, tg.prec, tg.ld, d);   // (q - mu) @ P
    T part[TM];
#pragma unroll
    for (int r = 0; r < TM; ++r) part[r] = T(0);
    if (tm.active) {
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        const int row = tm.rg * TM + r;
        if (row < rows_valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = 4 * tm.cg + j;
            if (c < d) {
              const size_t idx = (size_t)row * d + c;
              G[idx] = acc[r][j];
              if (want_lp) part[r] = fma_t(acc[r][j], q[idx] - (tg.mean ? tg.mean[c] : T(0)), part[r]);
            }
          }
        }
      }
    }
    if (want_lp) {
      row_reduce<T, BM, TM>(part, tm, red, lp_rows);
      if (tm.tid < rows_valid) {
        T lp = T(-0.5) * lp_rows[tm.tid];
        if (tg.has_lower0 && q[(size_t)tm.tid * d] < tg.lower0) lp = -inf_v<T>();
        lp_rows[tm.tid] = lp;
      }
    }
    __syncthreads();
  } else {
    target_rows_local<T>(tg, q, G, lp_rows, rows_valid, tm.tid, tm.nthreads);
  }
}

// ---- walk half-move ------------------------------------------------------------------------------
template <typename T>
struct WalkParams {
  TargetDev<T> tg;
  const T* X1;      // [rows, d] walkers being updated (local shard)
  const T* C;       // [n, ldd] centred complement / sqrt(n), zero padded
  const T* M;       // [ldd, ldd] = C^T C
  const T* lp1;     // [rows]
  T* Q;             // [rows, d]  out: proposed positions (working q)
  T* S;             // [rows, d]  out: projected momentum p @ C (working S)
  T* G;             // [rows, d]  scratch: grad U
  T* logacc;        // [rows] out
  T* lp_out;        // [rows] out
  T* dK_out;        // [rows] out, nullable
  int n, d, ldd, row0, rows, L, flags, legacy;
  T eps;
  Key key;
  const uint32_t* key_ptr;   // if non-null the move key is read from device memory
  // fused Metropolis accept + chain emit (batch level); off when fuse_accept == 0
  const T* eps_ptr;      // device-resident step size / trajectory length (adaptive warm-up); override eps / L when set
  const int* L_ptr;
  int fuse_accept;
  Key akey;
  const uint32_t* akey_ptr;
  T* X1w;           // [rows, d] state updated in place (aliases X1)
  T* lp1w;          // [rows]
  int32_t* accepts; // [rows]
  T* chain_out;     // [rows, d] nullable
  T* chain_lp;      // [rows] nullable
  int32_t* nonfinite;
};

template <typename T>
__device__ __forceinline__ void accept_and_emit(int legacy, T* Q, T* X1w, T* lp1w, const T* lp1, int32_t* accepts, T* chain_out,
                                                T* chain_lp, int32_t* nonfinite, Key akey, int n, int d, int grow0,
                                                int rows_valid, const T* lp_rows, const T* la_rows, int* mask,
                                                int tid, int nthreads) {
  // samplers/sampler_utils.py:114-122: log(uniform(key,(n,),1e-10,1)) < log_accept_prob (strict)
  if (tid < rows_valid) {
    const T lo = T(1e-10);
    const T span = T(1.0) - lo;
    const T u = uniform_rt<T>(akey, (uint64_t)(grow0 + tid), (uint64_t)n, lo, span, legacy);
    const bool acc = log_t(u) < la_rows[tid];
    mask[tid] = acc ? 1 : 0;
    const T lp = acc ? lp_rows[tid] : lp1[tid];
    lp1w[tid] = lp;
    if (chain_lp) chain_lp[tid] = lp;
    accepts[tid] += acc ? 1 : 0;
    if (!isfinite(lp) && nonfinite) atomicOr(nonfinite, 1);
  }
  __syncthreads();
  const int tot = rows_valid * d;
  for (int idx = tid; idx < tot; idx += nthreads) {
    const int r = idx / d;
    const T v = mask[r] ? Q[idx] : X1w[idx];
    X1w[idx] = v;
    if (chain_out) chain_out[idx] = v;
  }
}


// ---- momentum projection for narrow targets (dim <= 64): S0[i][:] = sum_j normal(key,(n,n))[row0+i][j] * C[j][:] ------------------
// RNG-bound (119 instructions per draw vs dim FMAs).  A CTA of 8 warps owns 8*WPW walkers; the centred complement streams through
// shared memory in tiles of TJ rows (each tile is read once per 8*WPW walkers).  Inside a warp the LANES split the complement
// index j and every lane keeps WPW x dim partial sums in registers, so there is no barrier per draw and WPW independent
// hash/erf_inv chains per lane; a butterfly adds the 32 partials at the end (fixed order: results depend only on n).
template <typename T, int DV /* float4/double4 groups per row: ldd/4 <= 16 */, int WPW>
__global__ void __launch_bounds__(256) k_project_narrow(Key key, const uint32_t* key_ptr, int legacy, const T* __restrict__ C, int n,
                                                        int ldd, int d, int row0, int rows, T* __restrict__ S0) {
  constexpr int TJ = 64;                       // complement rows per shared tile
  extern __shared__ __align__(16) unsigned char smem_raw_pn[];
  T* tile = reinterpret_cast<T*>(smem_raw_pn);  // [TJ][sld]
  const int sld = (DV | 1) * 4;                 // row stride: an odd number of 16-byte groups => conflict-free vector reads
  if (key_ptr) { key.k0 = key_ptr[0]; key.k1 = key_ptr[1]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w0 = (blockIdx.x * 8 + warp) * WPW;            // first local walker of this warp
  T acc[WPW][DV * 4];
#pragma unroll
  for (int w = 0; w < WPW; ++w)
#pragma unroll
    for (int c = 0; c < DV * 4; ++c) acc[w][c] = T(0);
  const uint64_t nn = (uint64_t)n * (uint64_t)n;
  for (int j0 = 0; j0 < n; j0 += TJ) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < TJ * DV; idx += 256) {
      const int jj = idx / DV, g = idx - jj * DV;
      T v[4] = {T(0), T(0), T(0), T(0)};
      if (j0 + jj < n && 4 * g < ldd) ld4_ro(C + (size_t)(j0 + jj) * ldd + 4 * g, v);
      st4(tile + jj * sld + 4 * g, v);
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < TJ / 32; ++h) {
      const int jj = h * 32 + lane, j = j0 + jj;
      T x[WPW];
#pragma unroll
      for (int w = 0; w < WPW; ++w) {
        const int i = w0 + w;
        x[w] = (i < rows && j < n) ? normal_rt<T>(key, (uint64_t)(row0 + i) * (uint64_t)n + (uint64_t)j, nn, legacy) : T(0);
      }
#pragma unroll
      for (int g = 0; g < DV; ++g) {
        T v[4];
        ld4(tile + jj * sld + 4 * g, v);
#pragma unroll
        for (int w = 0; w < WPW; ++w)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[w][4 * g + k] = fma_t(x[w], v[k], acc[w][4 * g + k]);
      }
    }
  }
#pragma unroll
  for (int w = 0; w < WPW; ++w) {
    const int i = w0 + w;
#pragma unroll
    for (int c = 0; c < DV * 4; ++c) {
      T v = acc[w][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == (c & 31) && i < rows && c < d) S0[(size_t)i * d + c] = v;
    }
  }
}

template <typename T, int BM, int TM, int MAXT>
__global__ void __launch_bounds__(MAXT) k_walk_half(WalkParams<T> P) {
  using TM_t = TileMap<T, BM, TM>;
  constexpr int LDA = BM + 4;
  __shared__ __align__(16) T As[kBK * LDA];
  __shared__ T red[BM * 32];
  __shared__ T lp_rows[BM];
  __shared__ T dk_rows[BM];
  __shared__ T la_rows[BM];
  __shared__ int mask[BM];

  TM_t tm;
  tm.init(P.ldd);
  const int d = P.d;
  const int r0 = blockIdx.x * BM;
  const int rows_valid = min(BM, P.rows - r0);
  const int grow0 = P.row0 + r0;
  const size_t off = (size_t)r0 * d;
  T* Q = P.Q + off;
  T* S = P.S + off;
  T* G = P.G + off;
  const bool frozen = (P.flags & kFlagFrozen) != 0;
  const T eps_ = P.eps_ptr ? *P.eps_ptr : P.eps;
  const int L_ = P.L_ptr ? *P.L_ptr : P.L;

  Key key = P.key;
  if (P.key_ptr) { key.k0 = P.key_ptr[0]; key.k1 = P.key_ptr[1]; }

  if (!(P.flags & kFlagInPlaceQ)) {
    const T* X1 = P.X1 + off;
    for (int idx = tm.tid; idx < rows_valid * d; idx += tm.nthreads) Q[idx] = X1[idx];
  }

  T acc[TM][4];
  // S0 = P0[rows,:] @ C  (hmc_walk.py:36 fused with the first use of p)
  if (!(P.flags & kFlagGivenMomentum)) {
    RngLoader<T, BM> ld{key, P.legacy, (uint64_t)P.n, grow0, rows_valid};
    gemm_rows<T, BM, TM>(acc, tm, As, ld, P.C, P.ldd, P.n);
    if (tm.active) {
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        const int row = tm.rg * TM + r;
        if (row < rows_valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = 4 * tm.cg + j;
            if (c < d) S[(size_t)row * d + c] = acc[r][j];
          }
        }
      }
    }
  }
  __syncthreads();

  T dK[TM];
#pragma unroll
  for (int r = 0; r < TM; ++r) dK[r] = T(0);

  eval_target<T, BM, TM>(P.tg, tm, Q, G, As, red, lp_rows, rows_valid, L_ == 0);
  // L+1 kicks (half, L-1 full, half), every kick but the last followed by a drift (hmc_walk.py:85-102)
  for (int k = 0; k <= L_; ++k) {
    const T a = (k == 0 || k == L_) ? T(0.5) * eps_ : eps_;
    const bool drift = k < L_;
    RowsLoader<T, BM> ld{G, d, nullptr, rows_valid, d};
    gemm_rows<T, BM, TM>(acc, tm, As, ld, P.M, P.ldd, d);   // g @ M
    if (tm.active) {
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        const int row = tm.rg * TM + r;
        if (row < rows_valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = 4 * tm.cg + j;
            if (c < d) {
              const size_t idx = (size_t)row * d + c;
              const T s_old = S[idx];
              const T s_new = s_old - a * acc[r][j];
              dK[r] = fma_t(a * G[idx], s_old + s_new, dK[r]);
              S[idx] = s_new;
              if (drift) Q[idx] = Q[idx] + eps_ * s_new;
            }
          }
        }
      }
    }
    __syncthreads();
    if (drift && !frozen) eval_target<T, BM, TM>(P.tg, tm, Q, G, As, red, lp_rows, rows_valid, k == L_ - 1);
  }
  if (frozen) eval_target<T, BM, TM>(P.tg, tm, Q, G, As, red, lp_rows, rows_valid, true);

  row_reduce<T, BM, TM>(dK, tm, red, dk_rows);
  if (tm.tid < rows_valid) {
    const int r = tm.tid;
    const T dKr = T(-0.5) * dk_rows[r];
    const T lpn = lp_rows[r];
    T dH = (P.lp1[r0 + r] - lpn) + dKr;          // (U1 + K1) - (U0 + K0)   hmc_walk.py:54
    if (isnan(dH)) dH = inf_v<T>();               // hmc_walk.py:55
    const T la = min_t(T(0), -dH);                // hmc_walk.py:56
    la_rows[r] = la;
    P.logacc[r0 + r] = la;
    P.lp_out[r0 + r] = lpn;
    if (P.dK_out) P.dK_out[r0 + r] = dKr;
  }
  __syncthreads();
  if (P.fuse_accept) {
    Key akey = P.akey;
    if (P.akey_ptr) { akey.k0 = P.akey_ptr[0]; akey.k1 = P.akey_ptr[1]; }
    accept_and_emit<T>(P.legacy, Q, P.X1w + off, P.lp1w + r0, P.lp1 + r0, P.accepts + r0,
                               P.chain_out ? P.chain_out + off : nullptr, P.chain_lp ? P.chain_lp + r0 : nullptr,
                               P.nonfinite, akey, P.n, d, grow0, rows_valid, lp_rows, la_rows, mask, tm.tid,
                               tm.nthreads);
  }
}

// ---- side half-move ------------------------------------------------------------------------------
template <typename T>
struct SideParams {
  TargetDev<T> tg;
  const T* X1;        // [rows, d]
  const T* X2;        // [n, d] complement (uncentred)
  const int32_t* idx2;  // [rows, 2] complement indices for the local rows (k_choice2)
  const T* lp1;
  T* Q;               // [rows, d] out
  T* D;               // [rows, d] out: direction rows, scaled by p at the end (momentum_projected)
  T* G;               // scratch
  T* p_io;            // [rows] given/returned momentum (integrator-only entry), nullable
  T* logacc;
  T* lp_out;
  int n, d, ldd, row0, rows, L, flags, legacy;
  T eps;
  Key key;
  const uint32_t* key_ptr;
  const T* eps_ptr;      // device-resident step size / trajectory length (adaptive warm-up); override eps / L when set
  const int* L_ptr;
  int fuse_accept;
  Key akey;
  const uint32_t* akey_ptr;
  T* X1w;
  T* lp1w;
  int32_t* accepts;
  T* chain_out;
  T* chain_lp;
  int32_t* nonfinite;
};

template <typename T, int BM, int TM, int MAXT>
__global__ void __launch_bounds__(MAXT) k_side_half(SideParams<T> P) {
  using TM_t = TileMap<T, BM, TM>;
  constexpr int LDA = BM + 4;
  __shared__ __align__(16) T As[kBK * LDA];
  __shared__ T red[BM * 32];
  __shared__ T lp_rows[BM];
  __shared__ T pr_rows[BM];   // <g, D> per walker
  __shared__ T p_rows[BM];    // scalar momentum
  __shared__ T p0_rows[BM];
  __shared__ T la_rows[BM];
  __shared__ int mask[BM];

  TM_t tm;
  tm.init(P.ldd);
  const int d = P.d;
  const int r0 = blockIdx.x * BM;
  const int rows_valid = min(BM, P.rows - r0);
  const int grow0 = P.row0 + r0;
  const size_t off = (size_t)r0 * d;
  T* Q = P.Q + off;
  T* D = P.D + off;
  T* G = P.G + off;
  const bool frozen = (P.flags & kFlagFrozen) != 0;
  const T eps_ = P.eps_ptr ? *P.eps_ptr : P.eps;
  const int L_ = P.L_ptr ? *P.L_ptr : P.L;
  const bool given = (P.flags & kFlagGivenMomentum) != 0;

  Key key = P.key;
  if (P.key_ptr) { key.k0 = P.key_ptr[0]; key.k1 = P.key_ptr[1]; }

  if (!(P.flags & kFlagInPlaceQ)) {
    const T* X1 = P.X1 + off;
    for (int idx = tm.tid; idx < rows_valid * d; idx += tm.nthreads) Q[idx] = X1[idx];
  }
  if (!given) {
    // hmc_side.py:52: D = (X2[j1] - X2[j2]) / sqrt(2 d); hmc_side.py:54: p0 = normal(keys[n], (n,))
    const T sc = sqrt_t(T(2 * d));
    for (int idx = tm.tid; idx < rows_valid * d; idx += tm.nthreads) {
      const int r = idx / d, c = idx - r * d;
      const int j1 = P.idx2[2 * (r0 + r)], j2 = P.idx2[2 * (r0 + r) + 1];
      D[idx] = (P.X2[(size_t)j1 * d + c] - P.X2[(size_t)j2 * d + c]) / sc;
    }
    if (tm.tid < rows_valid) {
      const Key kmom = split_key_rt(key, (uint64_t)P.n, (uint64_t)P.n + 2, P.legacy);
      const T p0 = normal_rt<T>(kmom, (uint64_t)(grow0 + tm.tid), (uint64_t)P.n, P.legacy);
      p_rows[tm.tid] = p0;
      p0_rows[tm.tid] = p0;
    }
  } else if (tm.tid < rows_valid) {
    p_rows[tm.tid] = P.p_io[r0 + tm.tid];
    p0_rows[tm.tid] = p_rows[tm.tid];
  }
  __syncthreads();

  eval_target<T, BM, TM>(P.tg, tm, Q, G, As, red, lp_rows, rows_valid, L_ == 0);
  for (int k = 0; k <= L_; ++k) {
    const T a = (k == 0 || k == L_) ? T(0.5) * eps_ : eps_;
    const bool drift = k < L_;
    T part[TM];
#pragma unroll
    for (int r = 0; r < TM; ++r) part[r] = T(0);
    if (tm.active) {
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        const int row = tm.rg * TM + r;
        if (row < rows_valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = 4 * tm.cg + j;
            if (c < d) part[r] = fma_t(G[(size_t)row * d + c], D[(size_t)row * d + c], part[r]);
          }
        }
      }
    }
    row_reduce<T, BM, TM>(part, tm, red, pr_rows);
    if (tm.tid < rows_valid) p_rows[tm.tid] = p_rows[tm.tid] - a * pr_rows[tm.tid];   // hmc_side.py:110,119,130
    __syncthreads();
    if (drift) {
      for (int idx = tm.tid; idx < rows_valid * d; idx += tm.nthreads) {
        const int r = idx / d;
        Q[idx] = Q[idx] + eps_ * (p_rows[r] * D[idx]);                                 // hmc_side.py:115,124
      }
      __syncthreads();
      if (!frozen) eval_target<T, BM, TM>(P.tg, tm, Q, G, As, red, lp_rows, rows_valid, k == L_ - 1);
    }
  }
  if (frozen) eval_target<T, BM, TM>(P.tg, tm, Q, G, As, red, lp_rows, rows_valid, true);

  if (tm.tid < rows_valid) {
    const int r = tm.tid;
    const T pL = p_rows[r], p0 = p0_rows[r];
    const T lpn = lp_rows[r];
    T dH = (-lpn + T(0.5) * pL * pL) - (-P.lp1[r0 + r] + T(0.5) * p0 * p0);   // hmc_side.py:65-72
    if (isnan(dH)) dH = inf_v<T>();
    const T la = min_t(T(0), -dH);
    la_rows[r] = la;
    P.logacc[r0 + r] = la;
    P.lp_out[r0 + r] = lpn;
    if (P.p_io) P.p_io[r0 + r] = pL;
  }
  __syncthreads();
  // momentum_projected = D * p (hmc_side.py:133)
  for (int idx = tm.tid; idx < rows_valid * d; idx += tm.nthreads) D[idx] = D[idx] * p_rows[idx / d];
  if (P.fuse_accept) {
    Key akey = P.akey;
    if (P.akey_ptr) { akey.k0 = P.akey_ptr[0]; akey.k1 = P.akey_ptr[1]; }
    accept_and_emit<T>(P.legacy, Q, P.X1w + off, P.lp1w + r0, P.lp1 + r0, P.accepts + r0,
                               P.chain_out ? P.chain_out + off : nullptr, P.chain_lp ? P.chain_lp + r0 : nullptr,
                               P.nonfinite, akey, P.n, d, grow0, rows_valid, lp_rows, la_rows, mask, tm.tid,
                               tm.nthreads);
  }
}

// ---- log-density only ------------------------------------------------------------------------------
template <typename T>
struct LogProbParams {
  TargetDev<T> tg;
  const T* X;   // [rows, d]
  T* G;         // [rows, d] grad U out (scratch or user buffer)
  T* lp;        // [rows]
  int rows, ldd;
  int negate_grad;   // write grad log pi = -grad U
};

template <typename T, int BM, int TM, int MAXT>
__global__ void __launch_bounds__(MAXT) k_log_prob(LogProbParams<T> P) {
  using TM_t = TileMap<T, BM, TM>;
  constexpr int LDA = BM + 4;
  __shared__ __align__(16) T As[kBK * LDA];
  __shared__ T red[BM * 32];
  __shared__ T lp_rows[BM];
  TM_t tm;
  tm.init(P.ldd);
  const int d = P.tg.dim;
  const int r0 = blockIdx.x * BM;
  const int rows_valid = min(BM, P.rows - r0);
  const size_t off = (size_t)r0 * d;
  eval_target<T, BM, TM>(P.tg, tm, P.X + off, P.G + off, As, red, lp_rows, rows_valid, true);
  if (tm.tid < rows_valid) P.lp[r0 + tm.tid] = lp_rows[tm.tid];
  if (P.negate_grad) {
    __syncthreads();
    for (int idx = tm.tid; idx < rows_valid * d; idx += tm.nthreads) P.G[off + idx] = -P.G[off + idx];
  }
}

}  // namespace hx
