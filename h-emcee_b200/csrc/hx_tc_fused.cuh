// hx_tc_fused.cuh — momentum projection S0 = P0 C with P0 = normal(key, (n, n)) generated INSIDE the GEMM
// (reference: moves/hamiltonian/hmc_walk.py:36 draws P0, :105 / the first kick consume it only through P0 @ C).
//
// The staged path (k_rng_planes + k_tc_gemm) writes the 4.3 GB operand planes of P0 to HBM and reads them back, and
// its ALU-bound draw and tensor-bound GEMM run as two kernels on two streams.  Here P0 never leaves the SM:
//
//   * one CTA owns a 128-row x 512-column tile of S0 (ONE fp32 accumulator = all 512 TMEM columns), so a 128-row
//     tile of P0 is shared by at most two CTAs (dim <= 1024) -- a CTA PAIR (cluster (2,1,1); pairs pack all 148 SMs,
//     clusters of four strand 16 of them).  Each CTA of the pair draws HALF of every 128 x 64 K-block of P0
//     (threefry2x32 + XLA erf_inv in registers, work items of 4 rows x 64 K dealt round-robin over the RNG warps),
//     splits it into bf16 hi/lo and stores the 16-byte chunks, already in the 128-byte-swizzled K-major layout the
//     UMMA descriptor expects, into its own A stage; one lane then hands the item's 2 x 512 contiguous bytes to the
//     TMA engine, which copies them into the peer's A stage (cp.async.bulk shared::cta -> shared::cluster) and counts
//     the bytes on the peer's `fullA` mbarrier;
//   * warp 0 streams the B operand (C^T planes, bf16 hi/lo) with TMA, one plane of 256 columns x 64 K per stage;
//   * warp 1 issues, per 256-column unit and K-block, the 12 tcgen05.mma of the three-product split in the order
//     hi*hi, lo*hi (B hi plane), hi*lo (B lo plane) into the single accumulator; its commits free the B stages locally
//     and the A stage in BOTH CTAs (multicast commit);
//   * the RNG warps drain TMEM at the end (one thread per output row, 128-byte runs);
//   * split-K over grid.y (a function of n only): shorter waves, each slice's share of C^T stays in L2.
//
// Counters: element (i, j) of the (n, n) draw is counter i * n + j of the partitionable threefry layout (one hash per
// 32-bit word), so any row range reproduces the rows of the full draw (row shards, App. B of SURVEY.md).  The legacy
// layout pairs element e with e + n^2/2 in one hash and stays on the staged path.
//
// Measured on B200 (c5: 32768 x 32768 draw, dim 1000; scripts/fused_bench.py, profiles/r02_*): 5.0 ms alone at 1965 MHz
// (staged draw + GEMM on two streams: 5.9), of which the draw arithmetic alone needs 3.5 ms (120 clk per warp and normal and
// scheduler: 96 instructions, alu pipe (SHF / LOP3 of threefry, 2 clk each) and issue slots saturate together) and the
// MMA + TMA side alone 3.4 ms (shared-memory bandwidth: 288 KB of operand reads + 192 KB of stage writes per K-block
// and CTA at 128 B/clk).  The two sides are balanced to 3 %, which is why they couple (5.0 rather than 3.5).
#pragma once
#include "hx_tc_gemm.cuh"
#include "hx_rng.cuh"
#include <cuda_fp16.h>
#include <cuda_fp8.h>

namespace hx {
namespace tc {

__constant__ uint32_t c_shr9mul = 1u << 23;      // hi32(bits * 2^23) = bits >> 9 as one IMAD.HI (constant bank: no strength reduction)

// y0 ^ y1 of threefry2x32-20(key, (x0, x1)) with x0 = hi + ks0, x1 = lo + ks1 already injected.  70 instructions per hash:
// 20 x (IADD/IMAD, SHF.L.W, LOP3) + 5 key injections folded into three-input adds.  (Routing rotations through the fma pipe as
// IMAD.WIDE by 2^r with the OR folded into the round's LOP3 was measured on B200: every such rotation made the kernel slower.)
#ifndef HX_FUSED_IMAD_INJ
#define HX_FUSED_IMAD_INJ 1      // hash adds as IMAD by a constant-bank 1 (fma pipe) instead of IADD3 / VIADD (alu pipe, which binds): 4.43 -> 4.36 ms at c5
#endif
__constant__ uint32_t c_one = 1u;                // multiplier from the constant bank: ptxas cannot fold a * c[..] + b back into an IADD3
__device__ __forceinline__ uint32_t add_fma_pipe(uint32_t a, uint32_t b) {
#if HX_FUSED_IMAD_INJ
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(c_one), "r"(b));
  return r;
#else
  return a + b;
#endif
}
__device__ __forceinline__ uint32_t tf_bits(uint32_t x0, uint32_t x1, uint32_t ks0, uint32_t ks1, uint32_t ks2) {
#define HX_R(r) x0 = add_fma_pipe(x0, x1); x1 = __funnelshift_l(x1, x1, r) ^ x0;
#define HX_INJ(ka, kb) x0 = add_fma_pipe(x0, ka); x1 = add_fma_pipe(x1, kb);
  HX_R(13) HX_R(15) HX_R(26) HX_R(6)
  HX_INJ(ks1, ks2 + 1u)
  HX_R(17) HX_R(29) HX_R(16) HX_R(24)
  HX_INJ(ks2, ks0 + 2u)
  HX_R(13) HX_R(15) HX_R(26) HX_R(6)
  HX_INJ(ks0, ks1 + 3u)
  HX_R(17) HX_R(29) HX_R(16) HX_R(24)
  HX_INJ(ks1, ks2 + 4u)
  HX_R(13) HX_R(15) HX_R(26) HX_R(6)
  HX_INJ(ks2, ks0 + 5u)
#undef HX_INJ
#undef HX_R
  return x0 ^ x1;
}

// out of line: used only where the low counter word wraps inside a group of eight elements
__device__ __noinline__ uint32_t tf_bits_at(Key key, uint32_t ks2, uint64_t e) {
  return tf_bits((uint32_t)(e >> 32) + key.k0, (uint32_t)e + key.k1, key.k0, key.k1, ks2);
}

// random bits of eight consecutive elements of the (n, n) draw starting at 64-bit counter e (partitionable layout): straight-line
// fast path that assumes the low counter word does not wrap inside the group, and the fix-up for the group where it does
// (once per 2^32 draws).  Kept apart so that the fast path has no branch and ptxas can interleave it with other work.
__device__ __forceinline__ void bits8_fast(Key key, uint32_t ks2, uint64_t e, uint32_t (&b)[8]) {
  const uint32_t a0 = (uint32_t)(e >> 32) + key.k0, a1 = (uint32_t)e + key.k1;
#pragma unroll
  for (int t = 0; t < 8; ++t) b[t] = tf_bits(a0, add_fma_pipe(a1, (uint32_t)t), key.k0, key.k1, ks2);
}
__device__ __forceinline__ void bits8_fix_wrap(Key key, uint32_t ks2, uint64_t e, uint32_t (&b)[8]) {
  if ((uint32_t)e > 0xFFFFFFF7u) {
#pragma unroll
    for (int t = 0; t < 8; ++t) b[t] = tf_bits_at(key, ks2, e + (uint64_t)t);
  }
}

// jax.random.normal(float32) values of eight random words.  Same arithmetic, operation by operation, as
// normal_at<float, false> (hx_rng.cuh) -- the staged path's draw.  The central branch of XLA's erf_inv (w < 5) is evaluated
// for all eight in straight-line code; the tail branch (|u| > 0.9966, 0.34 % of the draws) is a separate fix-up.
__device__ __forceinline__ bool normals8_central(const uint32_t (&b)[8], float (&x)[8], float (&u)[8], float (&w)[8]) {
  // The alu pipe (SHF / LOP3 of the hash) is what binds this kernel: every alu instruction per draw costs ~2 % (dropping the no-op
  // FMNMX of the uniform map: 4.52 -> 4.43 ms at c5).  Tail detection = running maximum of w (one FMNMX per draw; FSETP + PLOP3
  // took two, a sum of w^2 on the fma pipe is a superset that sends most batches through the fix-up: 4.49 ms)
  float wmax = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    // (bits >> 9) | 0x3F800000 as one IMAD.HI: hi32(bits * 2^23) + 0x3F800000 (the two fields do not overlap)
    const uint32_t m = __umulhi(b[t], c_shr9mul) + 0x3F800000u;
    const float f = __uint_as_float(m) - 1.0f;
    // jax.random.uniform's max(minval, .) is a no-op here: f >= 0 and rounding is monotonic, so fl(2 f + lo) >= lo always -- one
    // alu-pipe instruction (FMNMX) less per draw, same bits as normal_at (hx_rng.cuh) keeps the max for arbitrary minval
    u[t] = fmaf(f, 2.0f, -0.99999994f);                           // 2 f is exact: one rounding, like fmul_rn + fadd_rn
    w[t] = -0.693147182f * fast_lg2(__fsub_rn(1.0f, __fmul_rn(u[t], u[t])));
    wmax = fmaxf(wmax, w[t]);
    const float ww = w[t] - 2.5f;
    float p = 2.81022636e-08f;
    p = fmaf(p, ww, 3.43273939e-07f);
    p = fmaf(p, ww, -3.5233877e-06f);
    p = fmaf(p, ww, -4.39150654e-06f);
    p = fmaf(p, ww, 0.00021858087f);
    p = fmaf(p, ww, -0.00125372503f);
    p = fmaf(p, ww, -0.00417768164f);
    p = fmaf(p, ww, 0.246640727f);
    p = fmaf(p, ww, 1.50140941f);
    x[t] = 1.41421354f * (p * u[t]);
  }
  return !(wmax < 5.0f);
}
__device__ __forceinline__ void normals8_fix_tail(float (&x)[8], const float (&u)[8], const float (&w)[8]) {
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    if (!(w[t] < 5.0f)) {
      const float ww = fast_sqrt(w[t]) - 3.0f;
      float p = -0.000200214257f;
      p = fmaf(p, ww, 0.000100950558f);
      p = fmaf(p, ww, 0.00134934322f);
      p = fmaf(p, ww, -0.00367342844f);
      p = fmaf(p, ww, 0.00573950773f);
      p = fmaf(p, ww, -0.0076224613f);
      p = fmaf(p, ww, 0.00943887047f);
      p = fmaf(p, ww, 1.00167406f);
      p = fmaf(p, ww, 2.83297682f);
      x[t] = 1.41421354f * (p * u[t]);
    }
  }
}

// bf16 hi/lo split of eight values, packed: hi[0..3], lo[0..3] hold elements (2t, 2t+1) in the low / high half-words.
// Identical to split_bf16 (hi = rn(x), lo = rn(x - hi)), two elements per F2FP.
__device__ __forceinline__ void split8_bf16(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const __nv_bfloat162 hp = __floats2bfloat162_rn(x[2 * t], x[2 * t + 1]);
    h[t] = *reinterpret_cast<const uint32_t*>(&hp);
    const float h0 = __uint_as_float(h[t] << 16), h1 = __uint_as_float(h[t] & 0xFFFF0000u);
    const __nv_bfloat162 lp = __floats2bfloat162_rn(x[2 * t] - h0, x[2 * t + 1] - h1);
    l[t] = *reinterpret_cast<const uint32_t*>(&lp);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// fp16 + fp8 form of eight values at TRUE scale (FMT 1): p16 = fp16(x) (main operand), v8 = e4m3(x), r8 = e5m2(x - fp16(x)).
// The correction products e4m3(x) * e5m2(res b) and e5m2(res x) * e4m3(b) need no rescaling (e5m2 reaches down to 2^-16, the
// fp16 residual of |x| < 6 is < 2^-9.6), so they accumulate into the SAME accumulator as the main product.
__device__ __forceinline__ void split8_f16e(const float (&x)[8], uint4& p16, uint2& v8, uint2& r8) {
  uint32_t h[4];
  uint16_t v[4], r[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const __half2 hp = __floats2half2_rn(x[2 * t], x[2 * t + 1]);
    h[t] = *reinterpret_cast<const uint32_t*>(&hp);
    const float2 hf = __half22float2(hp);
    v[t] = __nv_cvt_float2_to_fp8x2(make_float2(x[2 * t], x[2 * t + 1]), __NV_SATFINITE, __NV_E4M3);
    r[t] = __nv_cvt_float2_to_fp8x2(make_float2(x[2 * t] - hf.x, x[2 * t + 1] - hf.y), __NV_SATFINITE, __NV_E5M2);
  }
  p16 = make_uint4(h[0], h[1], h[2], h[3]);
  v8 = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
  r8 = make_uint2((uint32_t)r[0] | ((uint32_t)r[1] << 16), (uint32_t)r[2] | ((uint32_t)r[3] << 16));
}

// try_wait with a suspend-time hint: the hardware parks the warp until the phase completes (or the hint expires) instead of
// returning after its default time slice -- the one-thread TMA / MMA warps and blocked RNG warps would otherwise spin on
// try_wait + branch and take issue slots from the RNG warps of their scheduler (13 % of all issued instructions, measured)
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  uint32_t ok, polls = 0;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (!ok && ++polls > (1u << 22)) __trap();
  } while (!ok);
}
// Wait of a warp that is AHEAD of its consumer (RNG warps on an `empty` barrier): non-blocking test + fixed sleep.  A parked
// try_wait is woken by every barrier event of the CTA (item arrivals, bulk-copy completions, TMA completions, MMA commits:
// ~60 per K-block), and with up to 28 RNG warps parked that was ~900 wake-ups x 6 instructions per K-block and CTA -- a third
// of all issue slots (ncu: 37 % of the warp samples on NANOSLEEP.SYNCS).  A warp that runs ahead loses nothing by oversleeping.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t ok, polls = 0;
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(300);
    if (++polls > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, const uint2& v) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Bulk copy (TMA engine) of `bytes` contiguous bytes of this CTA's shared memory into the peer CTA's shared memory; its
// completion is counted, in bytes, on an mbarrier of the peer (complete_tx).  No fence or release on the writer's side towards
// the peer: a cluster-scope release costs MEMBAR.ALL.GPU per warp and K-block (the dominant stall of the first version of this
// kernel), and per-thread st.async stores (second version) wake every parked waiter of the peer 64 times per item.
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes, uint32_t cluster_mbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
               "r"(src_cta_addr), "r"(bytes), "r"(cluster_mbar)
               : "memory");
}
// commit of a single-CTA MMA sequence that arrives on the same barrier offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// K-major SWIZZLE_128B operand whose 8-row groups (1 KB atoms) are `sbo` bytes apart (umma_desc_k128 fixes 1024)
__device__ __forceinline__ uint64_t umma_desc_k128_sbo(uint32_t smem_addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct FusedArgs {
  Key key; const uint32_t* key_ptr;   // key by value, or device-resident (batch loops / graph replays)
  int n;            // walkers per group: row length of the (n, n) draw and the K extent
  int row0;         // global row index of local row 0 (row shard offset)
  int rows;         // valid rows of this launch
  int d;            // valid output columns
  int num_kb;       // padded K / 64
  int kb_per_slice; // split-K: grid.y slices of this many K-blocks; slice z stores to out + z * slice_stride
  size_t slice_stride;
  int units;        // 256-column units per CTA (1 or 2)
  int rowsB_pad;    // plane stride of the B planes, in rows
  float* out; int ldo;
  int diag;         // diagnostics: 1 = no draw (A garbage), 2 = no MMA; results are garbage
  int relaxed_wait; // RNG warps poll their `empty` barrier with test_wait + nanosleep instead of a parked try_wait
  int stagger_ns;   // RNG warps assigned to the p-th K-block of a round start p * stagger_ns late (see the RNG loop)
  int atomic_out;   // split-K with exactly two slices: both add into the zeroed output with red.global.add (x + y is commutative and
                    // 0 + x exact, so the result does not depend on which slice lands first) -- no slice buffers, no sum kernel
  const float* scale_ptr;   // FMT 1: {sc, 1/sc}, the power-of-two scale of the B planes; the epilogue multiplies by scale_ptr[1]
};

#ifndef HX_FUSED_IR
#define HX_FUSED_IR 8          // rows of a work item of the RNG warps: 4 or 8 (measured at c5: 8 rows, A ring 4 / B ring 3: 15.96 ms per iteration; 4 rows, 3 / 4: 16.22)
#endif
#ifndef HX_FUSED_SA
#define HX_FUSED_SA 4
#define HX_FUSED_SB 3
#endif
struct FusedSmem {
  static constexpr int A_PLANE = kBM * 128;            // 16 KB: 128 rows x 64 bf16
  static constexpr int A_STAGE = 2 * A_PLANE;          // hi + lo, INTERLEAVED per 8-row group: [group g: plane 0 (1 KB) | plane 1 (1 KB)], so
                                                       // that an 8-row work item is 2 KB contiguous (one bulk copy to the peer, one byte-count
                                                       // event on its barrier instead of two); the UMMA descriptors step 2 KB between groups
  static constexpr int A_GROUP = 2048;
  static constexpr int B_STAGE = 256 * 128;            // 32 KB: ONE plane (hi or lo) of 256 columns x 64 bf16
  static constexpr int SA = HX_FUSED_SA, SB = HX_FUSED_SB;
  static constexpr int BAR_OFF = SA * A_STAGE + SB * B_STAGE;       // 224 KB
  static constexpr int BYTES = BAR_OFF + 256 + 1024;                // + barriers + alignment slack
};

// SHARE: the two CTAs of a (2,1,1) cluster own the two 512-column halves of one 128-row tile and each draws half of the
// A tile for both; otherwise every CTA is on its own (dim <= 512).  NW = RNG warps per CTA.
// FMT 0: bf16 hi/lo operands, three kind::f16 products.  FMT 1: fp16 main product + two fp8 correction products at twice the
// rate (A stage: plane 0 = fp16(x), plane 1 rows = [64 x e4m3(x) | 64 x e5m2(x - fp16(x))]; B planes (k_center_planes_T fmt 2):
// plane 0 = fp16(b sc), plane 1 rows = [64 x e5m2(res) | 64 x e4m3(b sc)]): 8 instead of 12 MMAs and 192 KB instead of 288 KB of
// shared-memory operand reads per K-block and CTA -- the resource that bounds the MMA side of this kernel.
template <int NW, bool SHARE, int FMT>
__global__ void __launch_bounds__(64 + 32 * NW, 1)
k_tc_proj_fused(const __grid_constant__ CUtensorMap tmB, const FusedArgs f) {
  using L = FusedSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stA = smem;
  uint8_t* stB = smem + L::SA * L::A_STAGE;
  uint64_t* bars = (uint64_t*)(smem + L::BAR_OFF);
  uint64_t* fullA = bars;                          // [SA]  RNG warps (local arrivals + the peer's bytes) -> MMA
  uint64_t* emptyA = bars + L::SA;                 // [SA]  MMA commits of the sharing CTAs -> RNG warps
  uint64_t* fullB = bars + 2 * L::SA;              // [SB]  TMA -> MMA
  uint64_t* emptyB = bars + 2 * L::SA + L::SB;     // [SB]  MMA commit -> TMA
  uint64_t* tmem_full = bars + 2 * L::SA + 2 * L::SB;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * L::SA + 2 * L::SB + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = SHARE ? cluster_ctarank() : 0u;
  const int tile = SHARE ? ((int)blockIdx.x >> 1) : (int)blockIdx.x;
  const int m0 = tile * kBM;
  const int n0 = SHARE ? (int)rank * 512 : 0;
  const int kb0 = f.kb_per_slice ? (int)blockIdx.y * f.kb_per_slice : 0;
  const int num_kb = f.kb_per_slice ? min(f.kb_per_slice, f.num_kb - kb0) : f.num_kb;
  constexpr int NCTA = SHARE ? 2 : 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmB);
    for (int s = 0; s < L::SA; ++s) { mbar_init(&fullA[s], (SHARE ? 64 : 128) / HX_FUSED_IR); mbar_init(&emptyA[s], NCTA); }
    for (int s = 0; s < L::SB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  if (SHARE) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int units = f.units;

  if (warp == 0) {
    // ---- B producer: C^T planes, one plane of 256 columns x 64 K per stage, in the order the MMA warp releases them:
    //      per K-block and unit the hi plane (two products), then the lo plane (one product)
    if (lane == 0) {
      int it = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        for (int u = 0; u < units; ++u) {
#pragma unroll
          for (int pl = 0; pl < 2; ++pl, ++it) {
            const int s = it % L::SB;
            mbar_wait_parked(&emptyB[s], ((uint32_t)(it / L::SB) & 1u) ^ 1u);
            mbar_expect_tx(&fullB[s], (uint32_t)L::B_STAGE);
            tma_load_2d(stB + s * L::B_STAGE, &tmB, &fullB[s], (kb0 + kb) * 64, pl * f.rowsB_pad + n0 + u * 256);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: per K-block and unit  hi*hi, lo*hi (B hi plane), hi*lo (B lo plane) into the single accumulator -----------
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
      int it = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int sa = kb % L::SA;
        mbar_wait_parked(&fullA[sa], (uint32_t)(kb / L::SA) & 1u);
        fence_proxy_async();       // the A stage was written through the generic proxy (st.shared / st.async); the MMA reads it through the async proxy
        tc_fence_after();
        const uint32_t a_hi = smem_u32(stA + sa * L::A_STAGE), a_lo = a_hi + 1024u;
        const uint64_t ad_hi = umma_desc_k128_sbo(a_hi, L::A_GROUP), ad_lo = umma_desc_k128_sbo(a_lo, L::A_GROUP);
        for (int u = 0; u < units; ++u) {
          const uint32_t dcol = tmem_base + (uint32_t)(u * 256);
#pragma unroll
          for (int pl = 0; pl < 2; ++pl, ++it) {
            const int s = it % L::SB;
            mbar_wait_parked(&fullB[s], (uint32_t)(it / L::SB) & 1u);
            tc_fence_after();
            const uint64_t bd = umma_desc_k128(smem_u32(stB + s * L::B_STAGE));
            if (f.diag != 2 && FMT == 1) {
              constexpr uint32_t dims = ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
              constexpr uint32_t idesc16 = (1u << 4) | dims;                    // kind::f16, A = B = F16 (format 0)
              constexpr uint32_t idesc_vr = (1u << 4) | (1u << 10) | dims;      // kind::f8f6f4, A = E4M3 (0), B = E5M2 (1)
              constexpr uint32_t idesc_rv = (1u << 4) | (1u << 7) | dims;       // kind::f8f6f4, A = E5M2 (1), B = E4M3 (0)
              if (pl == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(dcol, ad_hi + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc16, (kb | k) != 0 ? 1u : 0u);
              } else {
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_f8(dcol, ad_lo + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_vr, 1u);
#pragma unroll
                for (int k = 2; k < 4; ++k) umma_f8(dcol, ad_lo + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_rv, 1u);
              }
            } else if (f.diag != 2) {
              if (pl == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(dcol, ad_hi + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(dcol, ad_lo + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(dcol, ad_hi + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
              }
            }
            umma_commit(&emptyB[s]);
          }
        }
        if (SHARE) umma_commit_mcast(&emptyA[sa], 3);
        else umma_commit(&emptyA[sa]);
      }
      umma_commit(tmem_full);
    }
  } else {
    // ---- RNG warps: draw, split, store into the sharing CTAs' A stages -----------------------------------------------------
    // Work item = IR rows x 64 K of one K-block (batches of 4 rows: 32 lanes x 8 consecutive elements); the RG items of a K-block and the K-blocks
    // themselves are dealt round-robin over the NW warps, so NW need not divide the tile and neighbouring warps work on
    // neighbouring K-blocks (at most NW / RG + 1 apart: inside the three-stage A ring).
    Key key = f.key;
    if (f.key_ptr) { key.k0 = f.key_ptr[0]; key.k1 = f.key_ptr[1]; }
    const uint32_t ks2 = key.k0 ^ key.k1 ^ 0x1BD11BDAu;
    constexpr int ROWS_CTA = SHARE ? 64 : 128;                    // rows of the A tile this CTA draws
    constexpr int IR = HX_FUSED_IR;                               // rows per work item
    constexpr int NB = IR / 4;                                    // batches of 4 rows x 64 K (32 lanes x 8 elements) per item
    constexpr int RG = ROWS_CTA / IR;                             // items per K-block
    // A warp's consecutive items are NW / RG K-blocks apart.  It may not skip a whole cycle of the A ring: the parity wait on
    // `emptyA` only tells the last two phases of a stage apart, so a warp that touches a stage for the first time two uses late
    // would pass while the stage is still waiting to be consumed (seen on B200 with 8-row items, 28 warps and a 3-stage ring:
    // overwritten stages and surplus arrivals on fullA -> launch failure).
    static_assert((NW + RG - 1) / RG <= L::SA, "RNG warps would skip a full cycle of the A ring: fewer warps, smaller items or a deeper ring");
    const int wr = warp - 2;                                      // 0 .. NW - 1
    const int chunk = lane & 7, rsub = lane >> 3;
    const uint32_t stA_local = smem_u32(stA);
    const uint32_t stA_peer = SHARE ? mapa_shared(stA_local, rank ^ 1u) : 0u;
    const uint32_t fullA_peer = SHARE ? mapa_shared(smem_u32(fullA), rank ^ 1u) : 0u;
    const uint64_t row_stride = (uint64_t)(uint32_t)f.n;
    const uint64_t e00 = (uint64_t)(uint32_t)(f.row0 + m0 + (int)rank * ROWS_CTA + rsub) * row_stride + (uint64_t)(kb0 * 64 + chunk * 8);
    const int nitems = num_kb * RG;
    // first element of batch b of item g (for this lane)
    auto elem = [&](int g, int b) -> uint64_t {
      const int kb = g / RG, rg = g - kb * RG;
      return e00 + (uint64_t)(IR * rg + 4 * b) * row_stride + (uint64_t)(kb * 64);
    };
    // software pipeline: the hash of the NEXT batch (alu pipe: SHF / LOP3) is issued in the same straight-line block as the
    // erf_inv polynomial and the operand split of the current one (fma pipe), so that each warp keeps both pipes busy by itself
    // All RNG warps run at the same rate, so without a stagger the NW / RG K-blocks of a round are completed at the same moment:
    // the MMA warp then works through a burst while the warps of the next round wait for free stages (ncu: a third of their
    // samples parked on emptyA although the draw is the slower side).  Starting the warps of the p-th K-block p * stagger late
    // would spread the completions evenly.  Measured on B200 (hx_ctx_set_tuning(24, ns), scripts/fused_knob.py): 0 ... 3200 ns per
    // position change nothing (4.644 +- 0.003 ms at c5) -- the first blocked wait re-aligns the warps -- so the default is 0.
    if (f.stagger_ns > 0 && wr >= RG) __nanosleep((unsigned)(f.stagger_ns * (wr / RG)));
    uint32_t bcur[8];
    {
      const uint64_t e = elem(wr, 0);
      bits8_fast(key, ks2, e, bcur);
      bits8_fix_wrap(key, ks2, e, bcur);
    }
    // NW a multiple of RG (28 RNG warps are not, 24 are for a CTA pair): a warp keeps its row group and advances NW / RG
    // K-blocks per item, so the element index of the next batch is the current one plus a constant
    constexpr bool kAligned = (NW % RG) == 0;
    const uint64_t rs4 = 4ull * row_stride;
    uint64_t e_item = elem(wr, 0);
    for (int g = wr; g < nitems; g += NW) {
      const int kb = g / RG, rg = g - kb * RG;        // (keeping rg loop-invariant by hand for the aligned case costs registers: slower)
      const int sa = kb % L::SA;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int r = (int)rank * ROWS_CTA + IR * rg + 4 * b + rsub;         // row inside the 128-row tile
        uint4 hi, lo;
        if (f.diag == 1) {
          hi = make_uint4(0u, 0u, 0u, 0u); lo = hi;
        } else {
          // (the batch after the last one is hashed too and dropped: no branch inside the straight-line block)
          uint32_t bnext[8];
          uint64_t en;
          if (kAligned) {
            en = (b + 1 < NB) ? e_item + (uint64_t)(b + 1) * rs4 : e_item + (uint64_t)((NW / RG) * 64);
            if (b + 1 == NB) e_item = en;
          } else {
            en = (b + 1 < NB) ? elem(g, b + 1) : elem(g + NW, 0);
          }
          float x[8], u[8], w[8];
          bits8_fast(key, ks2, en, bnext);
          const bool tail = normals8_central(bcur, x, u, w);
          if (tail) normals8_fix_tail(x, u, w);
          bits8_fix_wrap(key, ks2, en, bnext);
          if (FMT == 1) {
            uint2 v8, r8;
            split8_f16e(x, hi, v8, r8);
            lo = make_uint4(v8.x, v8.y, r8.x, r8.y);
          } else {
            split8_bf16(x, hi, lo);
          }
#pragma unroll
          for (int t = 0; t < 8; ++t) bcur[t] = bnext[t];
        }
        if (b == 0) {
          if (f.relaxed_wait) mbar_wait_relaxed(&emptyA[sa], ((uint32_t)(kb / L::SA) & 1u) ^ 1u);
          else mbar_wait_parked(&emptyA[sa], ((uint32_t)(kb / L::SA) & 1u) ^ 1u);
        }
        // row r of plane p lives at (r >> 3) * 2 KB + p * 1 KB + (r & 7) * 128, 16-byte chunks XOR-swizzled with the row (128B swizzle)
        const uint32_t so = (uint32_t)(sa * L::A_STAGE) + (uint32_t)((r >> 3) * L::A_GROUP + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
        st_shared_v4(stA_local + so, hi);
        if (FMT == 1) {
          // fp8 plane row = [64 x e4m3(x) | 64 x e5m2(res)]: this lane's 8 bytes of each at byte chunk * 8 / 64 + chunk * 8
          const uint32_t rowb = stA_local + (uint32_t)(sa * L::A_STAGE + (r >> 3) * L::A_GROUP + 1024 + (r & 7) * 128) + (uint32_t)((chunk & 1) << 3);
          st_shared_v2(rowb + ((uint32_t)((chunk >> 1) ^ (r & 7)) << 4), make_uint2(lo.x, lo.y));
          st_shared_v2(rowb + ((uint32_t)((4 + (chunk >> 1)) ^ (r & 7)) << 4), make_uint2(lo.z, lo.w));
        } else {
          st_shared_v4(stA_local + so + 1024u, lo);
        }
      }
      fence_proxy_async();         // generic-proxy stores -> visible to the async proxy (the tensor core's operand reads, the bulk copy)
      __syncwarp();
      // one arrival per item.  In a pair the item's IR rows x 128 bytes per plane (contiguous in the swizzled layout) are also
      // copied into the peer's stage by the TMA engine, and the arrival announces the bytes the peer copies into THIS CTA
      if (lane == 0) {
        if (SHARE) {
          if (IR == 8) {      // the item is one 8-row group: both planes in one 2 KB copy
            const uint32_t io = (uint32_t)(sa * L::A_STAGE) + (uint32_t)((((int)rank * ROWS_CTA + IR * rg) >> 3) * L::A_GROUP);
            bulk_copy_to_peer(stA_peer + io, stA_local + io, 2048u, fullA_peer + 8u * sa);
          } else {            // half a group: 4 rows x 128 bytes of each plane
            const int r0 = (int)rank * ROWS_CTA + IR * rg;
            const uint32_t io = (uint32_t)(sa * L::A_STAGE) + (uint32_t)((r0 >> 3) * L::A_GROUP + (r0 & 7) * 128);
            bulk_copy_to_peer(stA_peer + io, stA_local + io, IR * 128u, fullA_peer + 8u * sa);
            bulk_copy_to_peer(stA_peer + io + 1024u, stA_local + io + 1024u, IR * 128u, fullA_peer + 8u * sa);
          }
          mbar_expect_tx(&fullA[sa], 2u * IR * 128u);
        } else {
          mbar_arrive(&fullA[sa]);
        }
      }
    }
    // ---- epilogue: TMEM -> registers -> S0 (one thread per output row) ---------------------------------------------------------
    mbar_wait_backoff(tmem_full, 0, 256);
    tc_fence_after();
    const int q = warp & 3, grp = (warp - 2) >> 2;               // TMEM lane quadrant of this warp; column-chunk group
    const int row = m0 + 32 * q + lane;
    const bool vec = (f.d & 3) == 0 && (f.ldo & 3) == 0;
    float* out = f.out + (size_t)blockIdx.y * f.slice_stride;
    const float inv_sc = (FMT == 1) ? f.scale_ptr[1] : 1.0f;
#pragma unroll 1
    for (int cc = grp; cc < units * 8; cc += (NW + 3) / 4) {
      const int col0 = n0 + cc * 32;
      if (col0 >= f.d) break;             // warp-uniform
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(cc * 32), v);
      if (FMT == 1) {
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] *= inv_sc;
      }
      if (row < f.rows) {
        const int nval = min(32, f.d - col0);
        float* o = out + (size_t)row * f.ldo + col0;
        if (!f.atomic_out) {
          store32(o, v, nval, vec && nval == 32);
        } else if (vec && nval == 32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]),
                         "f"(v[4 * j + 3])
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nval) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(o + j), "f"(v[j]) : "memory");
        }
      }
    }
    tc_fence_before();
  }
  __syncwarp();
  tc_fence_before();
  if (SHARE) cluster_sync_all(); else __syncthreads();      // the peer's stages and barriers stay alive until both CTAs are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc
}  // namespace hx
