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
// warps 2-5 = epilogue (TMEM -> registers -> smem transpose -> coalesced global access).
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
  static constexpr int STG = 4 * 32 * 33 * 4;        // epilogue transpose staging (4 warps)
  static constexpr int BYTES = STAGES * STAGE + STG + 256 + 1024;   // + barriers + alignment slack
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  float* stg_all = (float*)(smem + L::STAGES * L::STAGE);
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
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) ------------------------------------
    const int q = warp & 3;
    float* stg = stg_all + (warp - 2) * 32 * 33;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int row_base = m0 + 32 * q;
    float rowacc = 0.f;                 // lane r holds the partial of row (row_base + r)
    const int NC = BN / 32;
#pragma unroll 1
    for (int cc = 0; cc < NC; ++cc) {
      const int col0 = n0 + cc * 32;
      if (col0 >= g.N) break;           // warp-uniform
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(cc * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
      __syncwarp();
      const int col = col0 + lane;
      const bool cok = col < g.N;
      const float mu = (EPI != EPI_STORE && g.mu && cok) ? g.mu[col] : 0.f;
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int row = row_base + r;
        if (row >= g.M) break;          // warp-uniform
        const float acc = stg[r * 33 + lane];
        float x = 0.f;
        if (cok) {
          if (EPI == EPI_STORE) {
            g.out[(size_t)row * g.ldo + col] = acc;
          } else if (EPI == EPI_GRAD) {
            const size_t idx = (size_t)row * g.N + col;
            g.G[idx] = acc;
            __nv_bfloat16 hi, lo;
            split_bf16(acc, hi, lo);
            const size_t pidx = (size_t)row * g.Kout_p + col;
            g.planes_out[pidx] = hi;
            g.planes_out[(size_t)g.rows_out_pad * g.Kout_p + pidx] = lo;
            if (g.want_lp) x = acc * (g.Q[idx] - mu);
          } else {
            const size_t idx = (size_t)row * g.N + col;
            const float s_old = g.S[idx];
            const float s_new = s_old - g.a * acc;
            x = g.a * g.G[idx] * (s_old + s_new);
            g.S[idx] = s_new;
            if (g.drift) {
              const float qn = g.Q[idx] + g.eps * s_new;
              g.Q[idx] = qn;
              __nv_bfloat16 hi, lo;
              split_bf16(qn - mu, hi, lo);
              const size_t pidx = (size_t)row * g.Kout_p + col;
              g.planes_out[pidx] = hi;
              g.planes_out[(size_t)g.rows_out_pad * g.Kout_p + pidx] = lo;
            }
          }
        }
        if (EPI != EPI_STORE) {
          const float s = warp_sum(x);
          if (lane == r) rowacc += s;
        }
      }
      __syncwarp();
    }
    if (EPI != EPI_STORE) {
      const int row = row_base + lane;
      if (row < g.M) {
        float* p = g.part + (size_t)row * g.NT + blockIdx.x;
        if (EPI == EPI_GRAD) { if (g.want_lp) *p = rowacc; }
        else *p += rowacc;
      }
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
