// hx_tc_gemm.cuh — tcgen05 / TMEM / TMA GEMM with split-bf16 operands and fused leapfrog epilogues.
//
//   D[M x N] (fp32, TMEM) = A[M x K] * B[N x K]^T
//
// A and B live in HBM as two bf16 "planes" each (hi = bf16(x), lo = bf16(x - hi)), K-major, zero padded:
// plane p of a matrix with R_pad rows is rows [p*R_pad, (p+1)*R_pad) of one [2*R_pad, K_pad] bf16 array.
// With nprod = 3 the kernel accumulates A_hi*B_hi + A_hi*B_lo + A_lo*B_hi in the same fp32 TMEM
// accumulator (~2^-16 relative product error, fp32-class); nprod = 1 uses the hi planes only.
//
// One CTA per 128 x BN output tile; warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread),
// warps 2-5 = epilogue (TMEM -> registers; one thread per output row, 128-byte runs per thread).
#pragma once
#include "hx_tc.cuh"
#include "hx_common.cuh"
#include <cuda_bf16.h>

namespace hx {
namespace tc {

constexpr int kBM = 128;    // UMMA M
constexpr int kBKe = 64;    // K elements per stage (= 128 B of bf16 = one swizzle atom)
constexpr int kThreads = 192;

enum : int { EPI_STORE = 0, EPI_GRAD = 1, EPI_KICK = 2 };

struct GemmArgs {
  int M, N;          // valid output rows / cols
  int Kp;            // padded K (multiple of 64)
  int rowsA_pad;     // rows per plane of A (multiple of 128)
  int rowsB_pad;     // rows per plane of B (multiple of BN)
  int nprod;         // 1 or 3
  // EPI_STORE
  float* out; int ldo;
  // leapfrog epilogues (row-major [M x N] fp32 state, ld = N)
  float* S; float* Q; float* G;
  const float* mu;                 // [N] or nullptr
  __nv_bfloat16* planes_out;       // EPI_GRAD: G planes; EPI_KICK: (Q - mu) planes; [2, rows_out_pad, Kout_p]
  int rows_out_pad, Kout_p;
  float* part;                     // EPI_GRAD: lp partial [M x NT] (written); EPI_KICK: dK partial [M x NT] (accumulated)
  int NT;                          // number of N tiles (row stride of `part`)
  float a, eps;
  int drift, want_lp;
};

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

template <int BN>
struct SmemLayout {
  static constexpr int A_PLANE = kBM * 128;          // 16 KB
  static constexpr int B_PLANE = BN * 128;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int STAGES = (BN >= 256) ? 2 : 3;
  static constexpr int STG = 0;
  static constexpr int BYTES = STAGES * STAGE + STG + 256 + 1024;   // + barriers + alignment slack
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* bars = (uint64_t*)(smem + L::STAGES * L::STAGE + L::STG);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + L::STAGES;     // [STAGES]
  uint64_t* tmem_full = bars + 2 * L::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * L::STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int num_kb = g.Kp / kBKe;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < L::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bytes = (g.nprod == 3 ? 2u : 1u) * (uint32_t)(L::A_PLANE + L::B_PLANE);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % L::STAGES;
        const uint32_t ph = (uint32_t)(kb / L::STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* st = stage_base + s * L::STAGE;
        mbar_expect_tx(&full[s], bytes);
        tma_load_2d(st, &tmA, &full[s], kb * kBKe, m0);
        tma_load_2d(st + 2 * L::A_PLANE, &tmB, &full[s], kb * kBKe, n0);
        if (g.nprod == 3) {
          tma_load_2d(st + L::A_PLANE, &tmA, &full[s], kb * kBKe, g.rowsA_pad + m0);
          tma_load_2d(st + 2 * L::A_PLANE + L::B_PLANE, &tmB, &full[s], kb * kBKe, g.rowsB_pad + n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % L::STAGES;
        const uint32_t ph = (uint32_t)(kb / L::STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(stage_base + s * L::STAGE);
        const uint32_t a_lo = a_hi + L::A_PLANE;
        const uint32_t b_hi = a_hi + 2 * L::A_PLANE;
        const uint32_t b_lo = b_hi + L::B_PLANE;
        const int np = g.nprod;
#pragma unroll 1
        for (int p = 0; p < np; ++p) {
          const uint64_t ad = umma_desc_k128(p == 2 ? a_lo : a_hi);
          const uint64_t bd = umma_desc_k128(p == 1 ? b_lo : b_hi);
#pragma unroll
          for (int k = 0; k < kBKe / 16; ++k) {
            umma_bf16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | p | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);     // frees the smem stage when these MMAs have read it
      }
      umma_commit(tmem_full);       // accumulator complete
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32); thread = one output row ----------
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int row = m0 + 32 * q + lane;
    const bool rok = row < g.M;
    const bool vec = (g.N & 3) == 0 && (EPI == EPI_STORE ? (g.ldo & 3) == 0 : true);
    float rowacc = 0.f;
    const int NC = BN / 32;
#pragma unroll 1
    for (int cc = 0; cc < NC; ++cc) {
      const int col0 = n0 + cc * 32;
      if (col0 >= g.N) break;           // warp-uniform
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(cc * 32), v);
      if (!rok) continue;
      const int nval = min(32, g.N - col0);
      const bool fast = vec && nval == 32;
      if (EPI == EPI_STORE) {
        float* o = g.out + (size_t)row * g.ldo + col0;
        if (fast) {
#pragma unroll
          for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nval) o[j] = v[j];
        }
      } else {
        const size_t base = (size_t)row * g.N + col0;
        const size_t pbase = (size_t)row * g.Kout_p + col0;
        const size_t plane = (size_t)g.rows_out_pad * g.Kout_p;
        float mu[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) mu[j] = 0.f;
        if (g.mu) {
          if (fast) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(g.mu + col0) + j);
              mu[4 * j] = t.x; mu[4 * j + 1] = t.y; mu[4 * j + 2] = t.z; mu[4 * j + 3] = t.w;
            }
          } else {
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nval) mu[j] = g.mu[col0 + j];
          }
        }
        if (EPI == EPI_GRAD) {
          // g = acc: store fp32 + bf16 planes; optional lp partial sum_j g_j (q_j - mu_j)
          float* Gp = g.G + base;
          if (fast) {
#pragma unroll
            for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(Gp)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nval) Gp[j] = v[j];
          }
          __align__(16) __nv_bfloat16 hi[32], lo[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) split_bf16(j < nval ? v[j] : 0.f, hi[j], lo[j]);
          if (fast) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              reinterpret_cast<uint4*>(g.planes_out + pbase)[j] = reinterpret_cast<const uint4*>(hi)[j];
              reinterpret_cast<uint4*>(g.planes_out + plane + pbase)[j] = reinterpret_cast<const uint4*>(lo)[j];
            }
          } else {
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nval) { g.planes_out[pbase + j] = hi[j]; g.planes_out[plane + pbase + j] = lo[j]; }
          }
          if (g.want_lp) {
            const float* Qp = g.Q + base;
            if (fast) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = reinterpret_cast<const float4*>(Qp)[j];
                rowacc = fmaf(v[4 * j], t.x - mu[4 * j], rowacc);
                rowacc = fmaf(v[4 * j + 1], t.y - mu[4 * j + 1], rowacc);
                rowacc = fmaf(v[4 * j + 2], t.z - mu[4 * j + 2], rowacc);
                rowacc = fmaf(v[4 * j + 3], t.w - mu[4 * j + 3], rowacc);
              }
            } else {
              _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nval) rowacc = fmaf(v[j], Qp[j] - mu[j], rowacc);
            }
          }
        } else {
          // kick: S -= a * acc ; dK += a * g * (S- + S+) ; optional drift q += eps * S+ (+ centred planes)
          float sv[32], gv[32], qv[32];
          float* Sp = g.S + base;
          const float* Gp = g.G + base;
          float* Qp = g.Q + base;
          if (fast) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = reinterpret_cast<const float4*>(Sp)[j];
              sv[4 * j] = t.x; sv[4 * j + 1] = t.y; sv[4 * j + 2] = t.z; sv[4 * j + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = reinterpret_cast<const float4*>(Gp)[j];
              gv[4 * j] = t.x; gv[4 * j + 1] = t.y; gv[4 * j + 2] = t.z; gv[4 * j + 3] = t.w;
            }
            if (g.drift) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = reinterpret_cast<const float4*>(Qp)[j];
                qv[4 * j] = t.x; qv[4 * j + 1] = t.y; qv[4 * j + 2] = t.z; qv[4 * j + 3] = t.w;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { sv[j] = 0.f; gv[j] = 0.f; qv[j] = 0.f; }
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nval) { sv[j] = Sp[j]; gv[j] = Gp[j]; if (g.drift) qv[j] = Qp[j]; }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float s_new = sv[j] - g.a * v[j];
            rowacc = fmaf(g.a * gv[j], sv[j] + s_new, rowacc);
            sv[j] = s_new;
            qv[j] = qv[j] + g.eps * s_new;
          }
          if (fast) {
#pragma unroll
            for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(Sp)[j] = make_float4(sv[4 * j], sv[4 * j + 1], sv[4 * j + 2], sv[4 * j + 3]);
          } else {
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nval) Sp[j] = sv[j];
          }
          if (g.drift) {
            __align__(16) __nv_bfloat16 hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) split_bf16(j < nval ? qv[j] - mu[j] : 0.f, hi[j], lo[j]);
            if (fast) {
#pragma unroll
              for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(Qp)[j] = make_float4(qv[4 * j], qv[4 * j + 1], qv[4 * j + 2], qv[4 * j + 3]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                reinterpret_cast<uint4*>(g.planes_out + pbase)[j] = reinterpret_cast<const uint4*>(hi)[j];
                reinterpret_cast<uint4*>(g.planes_out + plane + pbase)[j] = reinterpret_cast<const uint4*>(lo)[j];
              }
            } else {
              _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nval) { Qp[j] = qv[j]; g.planes_out[pbase + j] = hi[j]; g.planes_out[plane + pbase + j] = lo[j]; }
            }
          }
        }
      }
    }
    if (EPI != EPI_STORE && rok) {
      float* p = g.part + (size_t)row * g.NT + blockIdx.x;
      if (EPI == EPI_GRAD) { if (g.want_lp) *p = rowacc; }
      else *p += rowacc;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

}  // namespace tc
}  // namespace hx
