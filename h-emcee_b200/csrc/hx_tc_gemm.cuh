// hx_tc_gemm.cuh — tcgen05 / TMEM / TMA GEMM with split operands and fused leapfrog epilogues.
//
//   D[M x N] (fp32, TMEM) = A[M x K] * B[N x K]^T
//
// A and B live in HBM as two "planes" each (hi, lo with x ~= hi + lo), K-major, zero padded: plane p of a
// matrix with R_pad rows is rows [p*R_pad, (p+1)*R_pad) of one [2*R_pad, K_pad] array.  The kernel
// accumulates A_hi*B_hi + A_hi*B_lo + A_lo*B_hi in one fp32 TMEM accumulator (nprod = 3), or only the
// first product (nprod = 1).  Two operand kinds:
//   KIND_BF16  hi = bf16(x), lo = bf16(x - hi)            kind::f16, ~2^-16 per product   (momentum projection, Gram)
//   KIND_TF32  hi = x with 13 low mantissa bits cleared,   kind::tf32, ~2^-21 per product  (leapfrog kicks: the
//              lo = x - hi (fp32 containers)                                                energy error budget needs it)
//
// One CTA per 128 x BN output tile; warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread),
// warps 2-5 = epilogue (TMEM -> registers; one thread per output row, 128-byte runs per thread).
#pragma once
#include "hx_tc.cuh"
#include "hx_common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace hx {
namespace tc {

constexpr int kBM = 128;      // UMMA M
constexpr int kThreads = 192;

enum : int { KIND_BF16 = 0, KIND_TF32 = 1, KIND_F16F8 = 2, KIND_F16 = 3 };
enum : int { EPI_STORE = 0, EPI_KICKB = 1, EPI_DOT = 2, EPI_PROP = 3 };

template <int KIND> struct Kind;
template <> struct Kind<KIND_BF16> {
  typedef __nv_bfloat16 elem;
  static constexpr int ROW_ELEMS = 64;   // elements per 128-byte swizzle row = K per stage
  static constexpr int UMMA_K = 16;
};
// fp16 main operand + two e4m3 correction operands (EPI_STORE only; see k_tc_gemm): plane 0 = fp16(x), plane 1 = per
// 64-element K-block 64 bytes of e4m3(x) followed by 64 bytes of e4m3((x - fp16(x)) * 2^11), i.e. the same 128-byte rows
template <> struct Kind<KIND_F16F8> {
  typedef uint16_t elem;
  static constexpr int ROW_ELEMS = 64;
  static constexpr int UMMA_K = 16;
};
// fp16 hi/lo (22 significand bits, 2^-22 per product at the kind::f16 rate, twice the tf32 rate): hi = fp16(x s), lo = fp16(x s - hi)
// with a power-of-two scale s per operand from its max-abs (f16_scale) so that the pair stays in fp16's normal range; the
// epilogues divide the scales out again.  Used for the propagator-form energy products.
template <> struct Kind<KIND_F16> {
  typedef __half elem;
  static constexpr int ROW_ELEMS = 64;
  static constexpr int UMMA_K = 16;
};
constexpr float kF8ResidualScale = 2048.0f;     // 2^11: the fp16 rounding residual of |x| < 64 times this is < 32
template <> struct Kind<KIND_TF32> {
  typedef float elem;
  static constexpr int ROW_ELEMS = 32;
  static constexpr int UMMA_K = 8;
};

struct GemmArgs {
  int M, N;          // valid output rows / cols
  int Kp;            // padded K (multiple of 64)
  int rowsA_pad;     // rows per plane of A (multiple of 128)
  int rowsB_pad;     // rows per plane of B (multiple of BN)
  int nprod;         // 1 or 3
  // EPI_STORE: out[row*ldo + col] = acc
  float* out; int ldo;
  // EPI_KICKB (acc = (q - mu) @ Bop): S = S_in - a*acc ; Qs (+)= a*(Q - mu) ; if drift: Q += eps*S, planes <- split(Q - mu)
  const float* S_in; float* S; float* Q; float* Qs;
  const float* mu;                 // [N] or nullptr
  void* planes_out;                // [2, rows_out_pad, Kout_p] in the kernel's operand kind
  int rows_out_pad, Kout_p;
  float a, eps;
  int drift, first;
  // EPI_DOT: part[row*NT + ntile] = sum_cols acc * (X + Y - mu)   (Y, mu optional)
  const float* X; const float* Y;
  float* part; int NT;
  // EPI_PROP (acc = [q0 - mu | S0] @ [Tq | Ts], the L-step leapfrog propagator of a Gaussian target): the N axis holds
  // two column blocks of blk_cols (padded) columns each; block 0: Q = acc + mu, planes_out (tf32 hi/lo) <- split(acc);
  // block 1: S = acc.  blk_cols = 0 for the other epilogues (one block).
  int blk_cols;
  int planes_bf16;   // EPI_PROP: kind of planes_out (0 tf32 hi/lo, 1 bf16 hi/lo, 2 fp16 hi/lo scaled by row_scale)
  // split-K (EPI_STORE only): grid.z slices of kb_per_slice K-blocks each, slice z stores to out + z * slice_stride; the
  // caller adds the slices in a fixed order (also shortens the tensor core's truncating accumulation chains).  0 = no split.
  int kb_per_slice; size_t slice_stride;
  // KIND_F16F8: out = (main + corr / 2^11) * scale_ptr[1]   (scale_ptr = {sc, 1/sc}, the power-of-two scale of the B operand)
  const float* scale_ptr;
  // fp16 energy operands.  EPI_STORE with amax_out: atomicMax of |out| into it (max-abs of W P);
  // EPI_DOT with dot_pair = 1 / 2: partial sums are divided by row_scale[row] * f16_scale(amax[1] / amax[0]) (operands Z x WP / Z x P);
  // EPI_PROP with planes_bf16 = 2: planes_out = fp16 hi/lo of acc * row_scale[row]
  uint32_t* amax_out; const uint32_t* amax; const float* row_scale; int dot_pair;
  // KIND_F16 with EPI_KICKB / EPI_PROP: acc is divided by row_scale[row] * f16_scale(amax[unscale_idx]) before use
  int unscale_idx;
  int n_cover;       // pair kernel: padded N the grid covers (0: N); EPI_PROP covers two column blocks
  int epi_force_row; // launch-site override: 1 = row-per-lane epilogue for this launch whatever the ctx mask says
  int row_epilogue;  // bit EPI set: row-per-lane epilogue for that epilogue kind, clear: coalesced chunks through shared memory (hx_ctx_set_tuning(23, mask))
  int one_acc;       // single-CTA kernel, bf16 hi/lo: all three products into ONE accumulator in the order (hi hi, lo hi, hi lo) per
                     // K-block -- the accumulation structure of the fused draw + projection kernel (hx_tc_fused.cuh), kept here
                     // so that the two can be compared bit for bit (hx_ctx_set_tuning(18, 1))
};

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// hi = RN_tf32(x), lo = RN_tf32(x - hi): both parts are exact tf32 values, so the tensor core's own truncation of the fp32
// container is a no-op.  (Masking the low 13 bits instead rounds hi AND the truncated lo toward zero: a systematic
// -2^-22 relative bias per product that shows up as a ~1e-2 bias of the energy error on kappa = 1e4 targets.)
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = rn_tf32(x);
  lo = rn_tf32(x - hi);
}

__device__ __forceinline__ void split_f16(float xs, __half& hi, __half& lo) {
  hi = __float2half_rn(xs);
  lo = __float2half_rn(xs - __half2float(hi));
}
// power-of-two scale that puts a max-abs value (given by its bit pattern) into [2^(top-1), 2^top)
__device__ __forceinline__ float f16_scale(uint32_t amax_bits, int top) {
  const float amax = __uint_as_float(amax_bits);
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);
  return ldexpf(1.f, min(top - e, 100));      // (nearly) all-zero operands: keep the scale and its reciprocal finite
}
// A operands (z0 = [q0 - mu | S0] and q_L - mu) are scaled per ROW (walker) from the row's own max-abs -- row-local, so a
// row-sharded run reproduces the unsharded one bit for bit; q_L - mu reuses its walker's z0 scale, which leaves 16x head-room.
// B operands (P, W P) are replicated on every rank and use one scale each from max-abs words [0] P, [1] W P.
constexpr int kF16TopZ = 12, kF16TopB = 15;

// 32 contiguous floats of one row: vector path when `fast`, predicated scalar path otherwise
__device__ __forceinline__ void load32(const float* p, float (&v)[32], int nval, bool fast) {
  if (fast) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = reinterpret_cast<const float4*>(p)[j];
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (j < nval) ? p[j] : 0.f;
  }
}
__device__ __forceinline__ void store32(float* p, const float (&v)[32], int nval, bool fast) {
  if (fast) {
#pragma unroll
    for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j < nval) p[j] = v[j];
  }
}
// split 32 values into the hi/lo planes of the given kind (plane stride in elements)
template <int KIND>
__device__ __forceinline__ void store_planes32(void* planes, size_t pbase, size_t plane, const float (&v)[32], int nval, bool fast,
                                               float scale = 1.f) {
  if (KIND == KIND_F16) {
    __half* P = reinterpret_cast<__half*>(planes);
    __align__(16) __half hi[32], lo[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) split_f16(j < nval ? v[j] * scale : 0.f, hi[j], lo[j]);
    if (fast) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        reinterpret_cast<uint4*>(P + pbase)[j] = reinterpret_cast<const uint4*>(hi)[j];
        reinterpret_cast<uint4*>(P + plane + pbase)[j] = reinterpret_cast<const uint4*>(lo)[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) { P[pbase + j] = hi[j]; P[plane + pbase + j] = lo[j]; }
    }
  } else if (KIND == KIND_BF16) {
    __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(planes);
    __align__(16) __nv_bfloat16 hi[32], lo[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) split_bf16(j < nval ? v[j] : 0.f, hi[j], lo[j]);
    if (fast) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        reinterpret_cast<uint4*>(P + pbase)[j] = reinterpret_cast<const uint4*>(hi)[j];
        reinterpret_cast<uint4*>(P + plane + pbase)[j] = reinterpret_cast<const uint4*>(lo)[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) { P[pbase + j] = hi[j]; P[plane + pbase + j] = lo[j]; }   // zero K-padding
    }
  } else {
    float* P = reinterpret_cast<float*>(planes);
    float hi[32], lo[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) split_tf32(j < nval ? v[j] : 0.f, hi[j], lo[j]);
    store32(P + pbase, hi, 32, fast);        // values beyond nval are zero: keeps the K-padding of the planes zero
    store32(P + plane + pbase, lo, 32, fast);
  }
}

template <int BN>
struct SmemLayout {
  static constexpr int A_PLANE = kBM * 128;          // 16 KB
  static constexpr int B_PLANE = BN * 128;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int STAGES = (BN >= 256) ? 2 : 3;
  static constexpr int BYTES = STAGES * STAGE + 256 + 1024;   // + barriers + alignment slack
};

template <int KIND>
__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (KIND != KIND_TF32) {
    umma_bf16(tmem_d, ad, bd, idesc, acc);
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
        : "memory");
  }
}

// ---- coalesced epilogue (round 2) ---------------------------------------------------------------------------------------------
// tcgen05.ld hands every lane 32 consecutive columns of ITS row, so a row-per-lane epilogue issues global accesses whose 32
// lanes touch 32 different cache lines (16 bytes each): the kick GEMM of the propagator basis run spent ~25 of its 43 us in
// those 32-wavefront requests (ncu: tensor pipe 13 % active, stalls on L1TEX scoreboards).  The warp therefore transposes each
// 32 x 32 chunk through a 4.6 KB tile of the (by then idle) operand stages: afterwards lane l holds, for i < 8, the four values
// of local row (l >> 3) + 4 i at columns 4 (l & 7) .. + 3, i.e. eight lanes cover 128 contiguous bytes of one row and a warp
// request touches 4 full lines.  Ragged chunks (fewer than 32 valid columns) keep the row-per-lane path.
// Measured at c5 (hx_ctx_set_tuning(23, mask), bit EPI = row path): the kick GEMM gains 6 % (basis run 0.647 -> 0.605 ms per
// half-move) -- its epilogue turned out to be bound by the latency of four lone warps, not by LSU wavefronts --, EPI_STORE and
// EPI_PROP are unchanged (stores do not stall), EPI_DOT loses 0.4 ms per half-move (eight shuffle hand-overs per row and chunk on
// warps that have no one to hide behind), so the default keeps DOT and PROP on the row path.
constexpr int kTilePitch = 36;                               // floats per tile row: 16-byte aligned, conflict-free both ways
constexpr int kTileBytes = 32 * kTilePitch * 4;
__device__ __forceinline__ void warp_transpose32(const float (&v)[32], float* tile, int lane, float4 (&o)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(tile + lane * kTilePitch + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = *reinterpret_cast<const float4*>(tile + ((lane >> 3) + 4 * i) * kTilePitch + 4 * (lane & 7));
  __syncwarp();
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
// hi/lo planes of four consecutive values (same splits as store_planes32)
template <int KIND>
__device__ __forceinline__ void store_planes4(void* planes, size_t pbase, size_t plane, const float4& v, float scale = 1.f) {
  const float x[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
  if (KIND == KIND_F16) {
    __half* P = reinterpret_cast<__half*>(planes);
    __align__(8) __half hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_f16(x[j], hi[j], lo[j]);
    *reinterpret_cast<uint2*>(P + pbase) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(P + plane + pbase) = *reinterpret_cast<const uint2*>(lo);
  } else if (KIND == KIND_BF16) {
    __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(planes);
    __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(x[j], hi[j], lo[j]);
    *reinterpret_cast<uint2*>(P + pbase) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(P + plane + pbase) = *reinterpret_cast<const uint2*>(lo);
  } else {
    float* P = reinterpret_cast<float*>(planes);
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_tf32(x[j], hi[j], lo[j]);
    st4(P + pbase, make_float4(hi[0], hi[1], hi[2], hi[3]));
    st4(P + plane + pbase, make_float4(lo[0], lo[1], lo[2], lo[3]));
  }
}

// Epilogue of one 128 x BN accumulator tile held in this CTA's TMEM (main product at columns [0, BN), correction products at
// [BN, 2 BN)): warp quadrant q owns TMEM lanes [32 q, 32 q + 32), one thread per output row.
template <int BN, int EPI, int KIND>
__device__ __forceinline__ void epilogue_rows(const GemmArgs& g, uint32_t tmem_base, int m0, int n0, int ntile, int q, int lane,
                                              float* tile = nullptr) {
  const int row = m0 + 32 * q + lane;
  const bool rok = row < g.M;
  const bool vec = (g.N & 3) == 0 && (EPI == EPI_STORE ? (g.ldo & 3) == 0 : true);
  float rowacc = 0.f;
  float dot8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // EPI_DOT, coalesced chunks: partial sums of local rows (lane >> 3) + 4 i
  float amax4 = 0.f;
  float f16_inv = 1.f;
  if (KIND == KIND_F16 && (EPI == EPI_KICKB || EPI == EPI_PROP) && rok)
    f16_inv = 1.f / (g.row_scale[row] * f16_scale(g.amax[g.unscale_idx], kF16TopB));
  // coalesced path: power-of-two operand scales of this lane's eight rows (fp16 hi/lo planes), read once
  float rs8[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
  if (tile && ((EPI == EPI_KICKB && KIND == KIND_F16 && g.drift) || (EPI == EPI_PROP && g.planes_bf16 == 2))) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = m0 + 32 * q + (lane >> 3) + 4 * i;
      if (r < g.M) rs8[i] = g.row_scale[r];
    }
  }
  const int NC = BN / 32;
  const int blk = (EPI == EPI_PROP) ? n0 / g.blk_cols : 0;
  const int cn0 = (EPI == EPI_PROP) ? n0 - blk * g.blk_cols : n0;
#pragma unroll 1
  for (int cc = 0; cc < NC; ++cc) {
    const int col0 = cn0 + cc * 32;
    if (col0 >= g.N) break;           // warp-uniform
    float v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(cc * 32), v);
    if (g.nprod == 3 && !g.one_acc) {
      float vc[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(BN + cc * 32), vc);
      if (KIND == KIND_F16F8) {
        const float so = g.scale_ptr[1];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(vc[j], 1.0f / kF8ResidualScale, v[j]) * so;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += vc[j];
      }
    }
    if (KIND == KIND_F16 && (EPI == EPI_KICKB || EPI == EPI_PROP)) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= f16_inv;
    }
    const int nval = min(32, g.N - col0);
    const bool fast = vec && nval == 32;
    if (tile && fast) {
      // ---- coalesced path: full 32-column chunk, eight lanes per row ----------------------------------------------------------
      float4 o[8];
      warp_transpose32(v, tile, lane, o);
      const int c4 = col0 + 4 * (lane & 7);
      float4 mu4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (EPI != EPI_STORE && g.mu && !(EPI == EPI_PROP && blk != 0)) mu4 = ld4(g.mu + c4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = m0 + 32 * q + (lane >> 3) + 4 * i;
        if (r >= g.M) continue;
        const float4 a = o[i];
        if (EPI == EPI_STORE) {
          st4(g.out + (size_t)blockIdx.z * g.slice_stride + (size_t)r * g.ldo + c4, a);
          if (g.amax_out) amax4 = fmaxf(fmaxf(amax4, fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
        } else if (EPI == EPI_PROP) {
          const size_t base = (size_t)r * g.N + c4;
          if (blk == 0) {
            st4(g.Q + base, make_float4(a.x + mu4.x, a.y + mu4.y, a.z + mu4.z, a.w + mu4.w));
            const size_t pb = (size_t)r * g.Kout_p + c4, pl = (size_t)g.rows_out_pad * g.Kout_p;
            if (g.planes_bf16 == 2) store_planes4<KIND_F16>(g.planes_out, pb, pl, a, rs8[i]);
            else if (g.planes_bf16) store_planes4<KIND_BF16>(g.planes_out, pb, pl, a);
            else store_planes4<KIND_TF32>(g.planes_out, pb, pl, a);
          } else {
            st4(g.S + base, a);
          }
        } else if (EPI == EPI_DOT) {
          const size_t base = (size_t)r * g.N + c4;
          float4 x = ld4(g.X + base);
          if (g.Y) {
            const float4 y = ld4(g.Y + base);
            x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
          }
          float t = dot8[i];
          t = fmaf(a.x, x.x - mu4.x, t); t = fmaf(a.y, x.y - mu4.y, t); t = fmaf(a.z, x.z - mu4.z, t); t = fmaf(a.w, x.w - mu4.w, t);
          dot8[i] = t;
        } else {
          // EPI_KICKB
          const size_t base = (size_t)r * g.N + c4;
          float4 sv = ld4(g.S_in + base), qv = ld4(g.Q + base);
          float4 ws = g.first ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(g.Qs + base);
          ws.x = fmaf(g.a, qv.x - mu4.x, ws.x); ws.y = fmaf(g.a, qv.y - mu4.y, ws.y);
          ws.z = fmaf(g.a, qv.z - mu4.z, ws.z); ws.w = fmaf(g.a, qv.w - mu4.w, ws.w);
          sv.x = sv.x - g.a * a.x; sv.y = sv.y - g.a * a.y; sv.z = sv.z - g.a * a.z; sv.w = sv.w - g.a * a.w;
          qv.x = qv.x + g.eps * sv.x; qv.y = qv.y + g.eps * sv.y; qv.z = qv.z + g.eps * sv.z; qv.w = qv.w + g.eps * sv.w;
          st4(g.S + base, sv);
          st4(g.Qs + base, ws);
          if (g.drift) {
            st4(g.Q + base, qv);
            const float4 qc = make_float4(qv.x - mu4.x, qv.y - mu4.y, qv.z - mu4.z, qv.w - mu4.w);
            const size_t pb = (size_t)r * g.Kout_p + c4, pl = (size_t)g.rows_out_pad * g.Kout_p;
            if (KIND == KIND_F16) store_planes4<KIND_F16>(g.planes_out, pb, pl, qc, rs8[i]);
            else store_planes4<KIND == KIND_F16F8 ? KIND_BF16 : KIND>(g.planes_out, pb, pl, qc);
          }
        }
      }
      continue;
    }
    if (!rok) continue;
    if (EPI == EPI_STORE) {
      store32(g.out + (size_t)blockIdx.z * g.slice_stride + (size_t)row * g.ldo + col0, v, nval, fast);
      if (g.amax_out) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < nval) rowacc = fmaxf(rowacc, fabsf(v[j]));
      }
    } else if (EPI == EPI_PROP) {
      const size_t base = (size_t)row * g.N + col0;
      if (blk == 0) {
        float qv[32];
        if (g.mu) {
          load32(g.mu + col0, qv, nval, fast);
#pragma unroll
          for (int j = 0; j < 32; ++j) qv[j] += v[j];
          store32(g.Q + base, qv, nval, fast);
        } else {
          store32(g.Q + base, v, nval, fast);
        }
        if (g.planes_bf16 == 2)
          store_planes32<KIND_F16>(g.planes_out, (size_t)row * g.Kout_p + col0, (size_t)g.rows_out_pad * g.Kout_p, v, nval, fast,
                                   g.row_scale[row]);
        else if (g.planes_bf16)
          store_planes32<KIND_BF16>(g.planes_out, (size_t)row * g.Kout_p + col0, (size_t)g.rows_out_pad * g.Kout_p, v, nval, fast);
        else
          store_planes32<KIND_TF32>(g.planes_out, (size_t)row * g.Kout_p + col0, (size_t)g.rows_out_pad * g.Kout_p, v, nval, fast);
      } else {
        store32(g.S + base, v, nval, fast);
      }
    } else {
      const size_t base = (size_t)row * g.N + col0;
      float mu[32];
      if (g.mu) load32(g.mu + col0, mu, nval, fast);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) mu[j] = 0.f;
      }
      if (EPI == EPI_DOT) {
        float x[32];
        load32(g.X + base, x, nval, fast);
        if (g.Y) {
          float y[32];
          load32(g.Y + base, y, nval, fast);
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] += y[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) rowacc = fmaf(v[j], x[j] - mu[j], rowacc);   // padded lanes hold acc = 0
      } else {
        // EPI_KICKB
        float sv[32], qv[32], ws[32];
        load32(g.S_in + base, sv, nval, fast);
        load32(g.Q + base, qv, nval, fast);
        if (g.first) {
#pragma unroll
          for (int j = 0; j < 32; ++j) ws[j] = 0.f;
        } else {
          load32(g.Qs + base, ws, nval, fast);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          ws[j] = fmaf(g.a, qv[j] - mu[j], ws[j]);        // impulse-weighted position sum (G_total = Qs @ P)
          sv[j] = sv[j] - g.a * v[j];
          qv[j] = qv[j] + g.eps * sv[j];
        }
        store32(g.S + base, sv, nval, fast);
        store32(g.Qs + base, ws, nval, fast);
        if (g.drift) {
          store32(g.Q + base, qv, nval, fast);
#pragma unroll
          for (int j = 0; j < 32; ++j) qv[j] -= mu[j];
          if (KIND == KIND_F16)
            store_planes32<KIND_F16>(g.planes_out, (size_t)row * g.Kout_p + col0, (size_t)g.rows_out_pad * g.Kout_p, qv, nval, fast,
                                     g.row_scale[row]);
          else
            store_planes32<KIND == KIND_F16F8 ? KIND_BF16 : KIND>(g.planes_out, (size_t)row * g.Kout_p + col0,
                                                                   (size_t)g.rows_out_pad * g.Kout_p, qv, nval, fast);
        }
      }
    }
  }
  if (EPI == EPI_DOT && tile) {
    // hand the coalesced partial sums to the row owners: the eight lanes of group (lane >> 3) hold local rows (lane >> 3) + 4 i
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = dot8[i];
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      const float mine = __shfl_sync(0xffffffffu, t, 8 * (lane & 3));      // local row `lane` lives in group lane & 3 at i = lane >> 2
      if ((lane >> 2) == i) rowacc += mine;
    }
  }
  if (EPI == EPI_STORE && g.amax_out) rowacc = fmaxf(rowacc, amax4);
  if (EPI == EPI_DOT && rok) {
    float sc = 1.f;
    if (g.dot_pair) sc = 1.f / (g.row_scale[row] * f16_scale(g.amax[g.dot_pair == 1 ? 1 : 0], kF16TopB));
    g.part[(size_t)row * g.NT + ntile] = rowacc * sc;
  }
  if (EPI == EPI_STORE && g.amax_out) {      // rows beyond M hold rowacc = 0: the max over the warp is the tile's max-abs
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rowacc = fmaxf(rowacc, __shfl_xor_sync(0xffffffffu, rowacc, o));
    if (lane == 0 && rowacc > 0.f) atomicMax(g.amax_out, __float_as_uint(rowacc));
  }
}

template <int BN, int EPI, int KIND>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using L = SmemLayout<BN>;
  using KD = Kind<KIND>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* bars = (uint64_t*)(smem + L::STAGES * L::STAGE);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + L::STAGES;     // [STAGES]
  uint64_t* tmem_full = bars + 2 * L::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * L::STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int num_kb_total = g.Kp / KD::ROW_ELEMS;
  const int kb0 = g.kb_per_slice ? (int)blockIdx.z * g.kb_per_slice : 0;
  const int num_kb = g.kb_per_slice ? min(g.kb_per_slice, num_kb_total - kb0) : num_kb_total;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < L::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  // two fp32 accumulators: columns [0, BN) take the hi*hi products, [BN, 2 BN) the two hi*lo correction products.  The tensor
  // core truncates at every accumulation step; keeping the 2^-9-sized corrections out of the main chain cuts the number of
  // truncations of the large partial sum by 3x, and the two parts are added once, round-to-nearest, in the epilogue.
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
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
        const int kc = (kb0 + kb) * KD::ROW_ELEMS;
        tma_load_2d(st, &tmA, &full[s], kc, m0);
        tma_load_2d(st + 2 * L::A_PLANE, &tmB, &full[s], kc, n0);
        if (g.nprod == 3) {
          tma_load_2d(st + L::A_PLANE, &tmA, &full[s], kc, g.rowsA_pad + m0);
          tma_load_2d(st + 2 * L::A_PLANE + L::B_PLANE, &tmB, &full[s], kc, g.rowsB_pad + n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // kind::f16 / kind::tf32 instruction descriptor: D = F32, A/B format (1 = BF16, 2 = TF32), K-major, N>>3, M>>4
      constexpr uint32_t fmt = (KIND == KIND_TF32) ? 2u : (KIND == KIND_F16 ? 0u : 1u);
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % L::STAGES;
        const uint32_t ph = (uint32_t)(kb / L::STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(stage_base + s * L::STAGE);
        const uint32_t a_lo = a_hi + L::A_PLANE;
        const uint32_t b_hi = a_hi + 2 * L::A_PLANE;
        const uint32_t b_lo = b_hi + L::B_PLANE;
        if (KIND == KIND_F16F8) {
          // main: fp16(a) * fp16(b) (kind::f16, 4 x K16); corrections into the second accumulator: e4m3(a) * e4m3(res b) and
          // e4m3(res a) * e4m3(b) (kind::f8f6f4, 4 x K32 at twice the rate) -- the 128-byte row of the second plane is
          // [value | residual], so K-step k of A pairs with K-step (k + 2) % 4 of B
          constexpr uint32_t idesc16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);   // A = B = F16
          constexpr uint32_t idesc8 = umma_idesc_e4m3(kBM, BN);
          const uint64_t ad = umma_desc_k128(a_hi), bd = umma_desc_k128(b_hi);
          const uint64_t ad8 = umma_desc_k128(a_lo), bd8 = umma_desc_k128(b_lo);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc16, (kb | k) != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f8(tmem_base + (uint32_t)BN, ad8 + (uint64_t)(2 * k), bd8 + (uint64_t)(2 * ((k + 2) & 3)), idesc8,
                    (kb | k) != 0 ? 1u : 0u);
        } else {
        const int np = g.nprod;
#pragma unroll 1
        for (int p0 = 0; p0 < np; ++p0) {
          const int p = g.one_acc ? (p0 == 0 ? 0 : 3 - p0) : p0;      // one accumulator: hi hi, lo hi, hi lo (the fused kernel's order)
          const uint64_t ad = umma_desc_k128(p == 2 ? a_lo : a_hi);
          const uint64_t bd = umma_desc_k128(p == 1 ? b_lo : b_hi);
#pragma unroll
          for (int k = 0; k < KD::ROW_ELEMS / KD::UMMA_K; ++k) {   // 4 steps of 32 bytes inside the 128-byte swizzle row
            const uint32_t acc = g.one_acc ? ((kb | p0 | k) != 0 ? 1u : 0u)
                                           : (p == 0 ? ((kb | k) != 0 ? 1u : 0u) : ((kb | (p - 1) | k) != 0 ? 1u : 0u));
            umma_issue<KIND>(tmem_base + ((p == 0 || g.one_acc) ? 0u : (uint32_t)BN), ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, acc);
          }
        }
        }
        umma_commit(&empty[s]);     // frees the smem stage when these MMAs have read it
      }
      umma_commit(tmem_full);       // accumulator complete
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32); thread = one output row ----------
    mbar_wait_backoff(tmem_full, 0, 512);
    tc_fence_after();
    epilogue_rows<BN, EPI, KIND>(g, tmem_base, m0, n0, (int)blockIdx.x, warp & 3, lane,
                                 ((g.row_epilogue >> EPI) & 1) ? nullptr : reinterpret_cast<float*>(stage_base + (warp & 3) * kTileBytes));
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---- CTA-pair variant (cta_group::2) ----------------------------------------------------------------------------------------
// Two CTAs of a (2, 1, 1) cluster own two consecutive 128-row tiles of one 256-column block and run ONE MMA of M = 256: each
// CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256 rows); the tensor cores read the other half from
// the peer's shared memory.  Per CTA and K-block that is 64 KB from L2 instead of 96 KB.  The single-CTA kernel is bound by the
// L2 -> SM operand stream (about 6.7 KB/clk over the chip, the measured TMA ceiling), not by the tensor pipe, so a third less
// operand traffic is a third more head-room; the smaller stage also buys a third pipeline stage in the same shared memory.
// Both CTAs issue TMA loads that complete on the LEADER's `full` barrier, the leader's elected thread issues the MMAs and its
// commits arrive on the `empty` / `tmem_full` barriers of both CTAs; every CTA drains its own 128 TMEM lanes.
struct SmemLayoutPair {
  static constexpr int A_PLANE = kBM * 128;          // 16 KB
  static constexpr int B_PLANE = 128 * 128;          // this CTA's half of the 256-row B tile
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int STAGES = 3;
  static constexpr int BYTES = STAGES * STAGE + 256 + 1024;
};

template <int EPI, int KIND>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_gemm_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  constexpr int BN = 256;
  using L = SmemLayoutPair;
  using KD = Kind<KIND>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* bars = (uint64_t*)(smem + L::STAGES * L::STAGE);
  uint64_t* full = bars;                  // [STAGES]  (used in the leader)
  uint64_t* empty = bars + L::STAGES;     // [STAGES]
  uint64_t* tmem_full = bars + 2 * L::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * L::STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D grid of (2, 1, 1) clusters: pair p = blockIdx.x / 2 owns column tile p % NT of row tiles 2 (p / NT) + {0, 1}, so
  // consecutive pairs sweep the column tiles of the same rows (A tile reuse in L2)
  const uint32_t rank = cluster_ctarank();           // 0 = leader
  const int pair_id = (int)blockIdx.x >> 1, ntiles = ((g.n_cover ? g.n_cover : g.N) + BN - 1) / BN;
  const int ntile = pair_id % ntiles;
  const int m0 = (2 * (pair_id / ntiles) + (int)rank) * kBM, n0 = ntile * BN;
  const int num_kb_total = g.Kp / KD::ROW_ELEMS;
  const int kb0 = g.kb_per_slice ? (int)blockIdx.z * g.kb_per_slice : 0;
  const int num_kb = g.kb_per_slice ? min(g.kb_per_slice, num_kb_total - kb0) : num_kb_total;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < L::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc_pair(tmem_slot, 2 * BN);
  tc_fence_before();
  cluster_sync_all();                     // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t planes = g.nprod == 3 ? 2u : 1u;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % L::STAGES;
        const uint32_t ph = (uint32_t)(kb / L::STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* st = stage_base + s * L::STAGE;
        if (rank == 0) mbar_expect_tx(&full[s], 2u * planes * (uint32_t)(L::A_PLANE + L::B_PLANE));   // both CTAs' bytes
        const uint32_t lbar = mapa_shared(smem_u32(&full[s]), 0);
        const int kc = (kb0 + kb) * KD::ROW_ELEMS;
        const int nb = n0 + (int)rank * 128;
        tma_load_2d_pair(st, &tmA, lbar, kc, m0);
        tma_load_2d_pair(st + 2 * L::A_PLANE, &tmB, lbar, kc, nb);
        if (g.nprod == 3) {
          tma_load_2d_pair(st + L::A_PLANE, &tmA, lbar, kc, g.rowsA_pad + m0);
          tma_load_2d_pair(st + 2 * L::A_PLANE + L::B_PLANE, &tmB, lbar, kc, g.rowsB_pad + nb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t fmt = (KIND == KIND_TF32) ? 2u : ((KIND == KIND_F16F8 || KIND == KIND_F16) ? 0u : 1u);
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * kBM) >> 4) << 24);
      constexpr uint32_t idesc8 = umma_idesc_e4m3(2 * kBM, BN);
      static_assert(KIND != KIND_TF32, "pair kernel: bf16 hi/lo or fp16 + e4m3 operands");
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % L::STAGES;
        const uint32_t ph = (uint32_t)(kb / L::STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(stage_base + s * L::STAGE);
        const uint32_t a_lo = a_hi + L::A_PLANE;
        const uint32_t b_hi = a_hi + 2 * L::A_PLANE;
        const uint32_t b_lo = b_hi + L::B_PLANE;
        if (KIND == KIND_F16F8) {
          const uint64_t ad = umma_desc_k128(a_hi), bd = umma_desc_k128(b_hi);
          const uint64_t ad8 = umma_desc_k128(a_lo), bd8 = umma_desc_k128(b_lo);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_pair_f16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_pair_f8(tmem_base + (uint32_t)BN, ad8 + (uint64_t)(2 * k), bd8 + (uint64_t)(2 * ((k + 2) & 3)), idesc8,
                         (kb | k) != 0 ? 1u : 0u);
        } else {
          const int np = g.nprod;
#pragma unroll 1
          for (int p = 0; p < np; ++p) {
            const uint64_t ad = umma_desc_k128(p == 2 ? a_lo : a_hi);
            const uint64_t bd = umma_desc_k128(p == 1 ? b_lo : b_hi);
#pragma unroll
            for (int k = 0; k < KD::ROW_ELEMS / KD::UMMA_K; ++k) {
              const uint32_t acc = p == 0 ? ((kb | k) != 0 ? 1u : 0u) : ((kb | (p - 1) | k) != 0 ? 1u : 0u);
              umma_pair_f16(tmem_base + (p == 0 ? 0u : (uint32_t)BN), ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, acc);
            }
          }
        }
        umma_commit_pair(&empty[s]);      // frees the stage in both CTAs
      }
      umma_commit_pair(tmem_full);        // accumulators of both CTAs complete
    }
  } else {
    mbar_wait_backoff(tmem_full, 0, 512);
    tc_fence_after();
    epilogue_rows<BN, EPI, KIND>(g, tmem_base, m0, n0, ntile, warp & 3, lane,
                                 ((g.row_epilogue >> EPI) & 1) ? nullptr : reinterpret_cast<float*>(smem + (warp & 3) * kTileBytes));
    tc_fence_before();
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();                     // the peer's shared memory and barriers stay alive until both CTAs are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 2 * BN);
  }
}

}  // namespace tc
}  // namespace hx
