// hx_rng.cuh — counter-based Threefry-2x32-20 and the jax.random value maps, device side.
//
// Bit-exact with jax.random for bits / split / uniform; normals use the same Giles polynomials
// as XLA's erf_inv (f32: ErfInv32, f64: ErfInv64).  f32 normals come in two forms selected by the RNG flags word
// (HX_RNGF_EXACT, set under HX_PREC_EXACT): exact = w from a correctly rounded log1p (<= 1 ulp from the oracle's
// np.log1p form, measured: tests/test_gpu_rng.py), fast = w from the SFU lg2 (<= 3 ulp, ~84 % bit-equal; log1p
// differs per backend even inside JAX).  Reference call sites: moves/hamiltonian/hmc_walk.py:36,
// moves/hamiltonian/hmc_side.py:37-54, samplers/sampler_utils.py:114,
// samplers/hamiltonian_ensemble.py:90,228,330.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace hx {

struct Key { uint32_t k0, k1; };

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(x, x, r);
#else
  return (x << r) | (x >> (32 - r));
#endif
}

// Threefry-2x32, 20 rounds; key schedule ks = {k0, k1, k0^k1^0x1BD11BDA}.
__host__ __device__ __forceinline__ void threefry2x32(Key k, uint32_t x0, uint32_t x1, uint32_t& y0,
                                                      uint32_t& y1) {
  const uint32_t ks0 = k.k0, ks1 = k.k1, ks2 = k.k0 ^ k.k1 ^ 0x1BD11BDAu;
  x0 += ks0; x1 += ks1;
#define HX_R(r) x0 += x1; x1 = rotl32(x1, r); x1 ^= x0;
  HX_R(13) HX_R(15) HX_R(26) HX_R(6)
  x0 += ks1; x1 += ks2 + 1u;
  HX_R(17) HX_R(29) HX_R(16) HX_R(24)
  x0 += ks2; x1 += ks0 + 2u;
  HX_R(13) HX_R(15) HX_R(26) HX_R(6)
  x0 += ks0; x1 += ks1 + 3u;
  HX_R(17) HX_R(29) HX_R(16) HX_R(24)
  x0 += ks1; x1 += ks2 + 4u;
  HX_R(13) HX_R(15) HX_R(26) HX_R(6)
  x0 += ks2; x1 += ks0 + 5u;
#undef HX_R
  y0 = x0; y1 = x1;
}

// Element e of `random_bits(key, 32, (m,))`.
//   partitionable: y0^y1 of counter (hi32(e), lo32(e))
//   legacy       : counters iota(m) padded to even m', halves paired: element e<h is y0 of
//                  (e, e+h), element e>=h is y1 of (e-h, e); the pad counter is 0.
template <bool LEGACY>
__host__ __device__ __forceinline__ uint32_t bits32(Key k, uint64_t e, uint64_t m) {
  uint32_t y0, y1;
  if (!LEGACY) {
    threefry2x32(k, (uint32_t)(e >> 32), (uint32_t)e, y0, y1);
    return y0 ^ y1;
  } else {
    const uint64_t h = (m + 1) >> 1;
    if (e < h) {
      const uint64_t c1 = e + h;
      threefry2x32(k, (uint32_t)e, (c1 < m) ? (uint32_t)c1 : 0u, y0, y1);
      return y0;
    }
    threefry2x32(k, (uint32_t)(e - h), (uint32_t)e, y0, y1);
    return y1;
  }
}

// Element e of `random_bits(key, 64, (m,))`.
//   partitionable: y0<<32 | y1 of counter e;  legacy: words[e]<<32 | words[m+e] of 2m words,
//   i.e. (y0, y1) of the single hash (e, m+e).
template <bool LEGACY>
__host__ __device__ __forceinline__ uint64_t bits64(Key k, uint64_t e, uint64_t m) {
  uint32_t y0, y1;
  if (!LEGACY) threefry2x32(k, (uint32_t)(e >> 32), (uint32_t)e, y0, y1);
  else threefry2x32(k, (uint32_t)e, (uint32_t)(m + e), y0, y1);
  return ((uint64_t)y0 << 32) | (uint64_t)y1;
}

// key i of split(key, num): partitionable = (y0,y1) of counter (0,i); legacy = flat words 2i, 2i+1
// of the hash over iota(2*num).
template <bool LEGACY>
__host__ __device__ __forceinline__ Key split_key(Key k, uint64_t i, uint64_t num) {
  Key o;
  if (!LEGACY) {
    threefry2x32(k, (uint32_t)(i >> 32), (uint32_t)i, o.k0, o.k1);
  } else {
    o.k0 = bits32<true>(k, 2 * i, 2 * num);
    o.k1 = bits32<true>(k, 2 * i + 1, 2 * num);
  }
  return o;
}

// ---- value maps -------------------------------------------------------------------------------
// jax.random._uniform: f = bitcast((bits >> (nbits - nmant)) | one) - 1 ; max(lo, f*(hi-lo)+lo).
__device__ __forceinline__ float unit_float(uint32_t b) { return __uint_as_float((b >> 9) | 0x3F800000u) - 1.0f; }
__device__ __forceinline__ double unit_float(uint64_t b) {
  return __longlong_as_double((long long)((b >> 12) | 0x3FF0000000000000ull)) - 1.0;
}

__device__ __forceinline__ float uniform_from_bits(uint32_t b, float lo, float span) {
  return fmaxf(lo, __fadd_rn(__fmul_rn(unit_float(b), span), lo));
}
__device__ __forceinline__ double uniform_from_bits(uint64_t b, double lo, double span) {
  return fmax(lo, __dadd_rn(__dmul_rn(unit_float(b), span), lo));
}

// XLA erf_inv, f32 (Giles single precision).
// w = -log1p(-x^2) is evaluated as -ln2 * lg2(1 - x2) with the SFU (MUFU.LG2, absolute error 2^-22.6 in log2), where
// x2 = rn(x*x) is rounded FIRST exactly like XLA's log1p(-x*x) argument (near |x| = 1 that rounding is the dominant error of
// the reference's own value, so it has to be reproduced, not improved on): the absolute error of w stays below ~2e-7, i.e. <= 1 ulp in the result (dp/dw ~ 0.3),
// the same order as the backend-to-backend log1p differences inside JAX itself; libdevice's log1pf costs ~22 FMA-pipe
// instructions per draw, which is 15 % of the whole draw (the walk move needs n^2 of them per half-move).
__device__ __forceinline__ float fast_lg2(float t) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}
__device__ __forceinline__ float fast_sqrt(float t) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}
__device__ __forceinline__ float erfinv_xla(float x) {
  float w = -0.693147182f * fast_lg2(__fsub_rn(1.0f, __fmul_rn(x, x)));
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, w, 3.43273939e-07f);
    p = fmaf(p, w, -3.5233877e-06f);
    p = fmaf(p, w, -4.39150654e-06f);
    p = fmaf(p, w, 0.00021858087f);
    p = fmaf(p, w, -0.00125372503f);
    p = fmaf(p, w, -0.00417768164f);
    p = fmaf(p, w, 0.246640727f);
    p = fmaf(p, w, 1.50140941f);
  } else {
    w = fast_sqrt(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, w, 0.000100950558f);
    p = fmaf(p, w, 0.00134934322f);
    p = fmaf(p, w, -0.00367342844f);
    p = fmaf(p, w, 0.00573950773f);
    p = fmaf(p, w, -0.0076224613f);
    p = fmaf(p, w, 0.00943887047f);
    p = fmaf(p, w, 1.00167406f);
    p = fmaf(p, w, 2.83297682f);
  }
  return p * x;
}

// Exact form of the f32 erf_inv (HX_RNGF_EXACT): w = -log1p(-x^2) rounded once from the f64 log1p (correctly rounded in all
// but ~1e-8 of the cases) and an IEEE sqrt, everything else as above.  Out of line: the f64 log1p would otherwise raise the
// register count of the fused SIMT move kernels for a branch that only HX_PREC_EXACT takes.
static __device__ __noinline__ float erfinv_xla_exact(float x) {
  float w = -(float)log1p(-(double)__fmul_rn(x, x));
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, w, 3.43273939e-07f);
    p = fmaf(p, w, -3.5233877e-06f);
    p = fmaf(p, w, -4.39150654e-06f);
    p = fmaf(p, w, 0.00021858087f);
    p = fmaf(p, w, -0.00125372503f);
    p = fmaf(p, w, -0.00417768164f);
    p = fmaf(p, w, 0.246640727f);
    p = fmaf(p, w, 1.50140941f);
  } else {
    w = __fsqrt_rn(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, w, 0.000100950558f);
    p = fmaf(p, w, 0.00134934322f);
    p = fmaf(p, w, -0.00367342844f);
    p = fmaf(p, w, 0.00573950773f);
    p = fmaf(p, w, -0.0076224613f);
    p = fmaf(p, w, 0.00943887047f);
    p = fmaf(p, w, 1.00167406f);
    p = fmaf(p, w, 2.83297682f);
  }
  return p * x;
}

// XLA erf_inv, f64 (Giles double precision, three branches on w = -log1p(-x^2)).
__device__ __forceinline__ double erfinv_xla(double x) {
  double w = -log1p(-x * x);
  double p;
  if (w < 6.25) {
    w = w - 3.125;
    p = -3.6444120640178196996e-21;
    p = fma(p, w, -1.685059138182016589e-19);
    p = fma(p, w, 1.2858480715256400167e-18);
    p = fma(p, w, 1.115787767802518096e-17);
    p = fma(p, w, -1.333171662854620906e-16);
    p = fma(p, w, 2.0972767875968561637e-17);
    p = fma(p, w, 6.6376381343583238325e-15);
    p = fma(p, w, -4.0545662729752068639e-14);
    p = fma(p, w, -8.1519341976054721522e-14);
    p = fma(p, w, 2.6335093153082322977e-12);
    p = fma(p, w, -1.2975133253453532498e-11);
    p = fma(p, w, -5.4154120542946279317e-11);
    p = fma(p, w, 1.051212273321532285e-09);
    p = fma(p, w, -4.1126339803469836976e-09);
    p = fma(p, w, -2.9070369957882005086e-08);
    p = fma(p, w, 4.2347877827932403518e-07);
    p = fma(p, w, -1.3654692000834678645e-06);
    p = fma(p, w, -1.3882523362786468719e-05);
    p = fma(p, w, 0.0001867342080340571352);
    p = fma(p, w, -0.00074070253416626697512);
    p = fma(p, w, -0.0060336708714301490533);
    p = fma(p, w, 0.24015818242558961693);
    p = fma(p, w, 1.6536545626831027356);
  } else if (w < 16.0) {
    w = sqrt(w) - 3.25;
    p = 2.2137376921775787049e-09;
    p = fma(p, w, 9.0756561938885390979e-08);
    p = fma(p, w, -2.7517406297064545428e-07);
    p = fma(p, w, 1.8239629214389227755e-08);
    p = fma(p, w, 1.5027403968909827627e-06);
    p = fma(p, w, -4.013867526981545969e-06);
    p = fma(p, w, 2.9234449089955446044e-06);
    p = fma(p, w, 1.2475304481671778723e-05);
    p = fma(p, w, -4.7318229009055733981e-05);
    p = fma(p, w, 6.8284851459573175448e-05);
    p = fma(p, w, 2.4031110387097893999e-05);
    p = fma(p, w, -0.0003550375203628474796);
    p = fma(p, w, 0.00095328937973738049703);
    p = fma(p, w, -0.0016882755560235047313);
    p = fma(p, w, 0.0024914420961078508066);
    p = fma(p, w, -0.0037512085075692412107);
    p = fma(p, w, 0.005370914553590063617);
    p = fma(p, w, 1.0052589676941592334);
    p = fma(p, w, 3.0838856104922207635);
  } else {
    w = sqrt(w) - 5.0;
    p = -2.7109920616438573243e-11;
    p = fma(p, w, -2.5556418169965252055e-10);
    p = fma(p, w, 1.5076572693500548083e-09);
    p = fma(p, w, -3.7894654401267369937e-09);
    p = fma(p, w, 7.6157012080783393804e-09);
    p = fma(p, w, -1.4960026627149240478e-08);
    p = fma(p, w, 2.9147953450901080826e-08);
    p = fma(p, w, -6.7711997758452339498e-08);
    p = fma(p, w, 2.2900482228026654717e-07);
    p = fma(p, w, -9.9298272942317002539e-07);
    p = fma(p, w, 4.5260625972231537039e-06);
    p = fma(p, w, -1.9681778105531670567e-05);
    p = fma(p, w, 7.5995277030017761139e-05);
    p = fma(p, w, -0.00021503011930044477347);
    p = fma(p, w, -0.00013871931833623122026);
    p = fma(p, w, 1.0103004648645343977);
    p = fma(p, w, 4.8499064014085844221);
  }
  return p * x;
}

template <typename T> struct RngTraits;
template <> struct RngTraits<float> {
  using bits_t = uint32_t;
  // nextafter(-1, 0) and (1 - lo) rounded in f32 (= 2.0f exactly after rounding)
  __device__ static __forceinline__ float normal_lo() { return -0.99999994f; }
  __device__ static __forceinline__ float normal_span() { return 2.0f; }
  __device__ static __forceinline__ float sqrt2() { return 1.41421354f; }
};
template <> struct RngTraits<double> {
  using bits_t = uint64_t;
  __device__ static __forceinline__ double normal_lo() { return -0.99999999999999989; }
  __device__ static __forceinline__ double normal_span() { return 2.0; }
  __device__ static __forceinline__ double sqrt2() { return 1.4142135623730951; }
};

template <typename T, bool LEGACY>
__device__ __forceinline__ typename RngTraits<T>::bits_t draw_bits(Key k, uint64_t e, uint64_t m);
template <> __device__ __forceinline__ uint32_t draw_bits<float, false>(Key k, uint64_t e, uint64_t m) { return bits32<false>(k, e, m); }
template <> __device__ __forceinline__ uint32_t draw_bits<float, true>(Key k, uint64_t e, uint64_t m) { return bits32<true>(k, e, m); }
template <> __device__ __forceinline__ uint64_t draw_bits<double, false>(Key k, uint64_t e, uint64_t m) { return bits64<false>(k, e, m); }
template <> __device__ __forceinline__ uint64_t draw_bits<double, true>(Key k, uint64_t e, uint64_t m) { return bits64<true>(k, e, m); }

// RNG flags word carried by the kernels' `legacy` argument: bit 0 = legacy counter layout, bit 1 = exact f32 normals
enum : int { HX_RNGF_LEGACY = 1, HX_RNGF_EXACT = 2 };

__device__ __forceinline__ float erfinv_sel(float u, bool exact) { return exact ? erfinv_xla_exact(u) : erfinv_xla(u); }
__device__ __forceinline__ double erfinv_sel(double u, bool) { return erfinv_xla(u); }

// element e of jax.random.normal(key, shape with m elements, dtype T)
template <typename T, bool LEGACY>
__device__ __forceinline__ T normal_at(Key k, uint64_t e, uint64_t m, bool exact = false) {
  const T u = uniform_from_bits(draw_bits<T, LEGACY>(k, e, m), RngTraits<T>::normal_lo(), RngTraits<T>::normal_span());
  return RngTraits<T>::sqrt2() * erfinv_sel(u, exact);
}

// element e of jax.random.uniform(key, (m,), minval=lo, maxval=lo+span) with span rounded in T
template <typename T, bool LEGACY>
__device__ __forceinline__ T uniform_at(Key k, uint64_t e, uint64_t m, T lo, T span) {
  return uniform_from_bits(draw_bits<T, LEGACY>(k, e, m), lo, span);
}

// runtime-selected counter layout and normal form (`flags` = HX_RNGF_* bits; the branches are uniform across the grid)
template <typename T>
__device__ __forceinline__ T normal_rt(Key k, uint64_t e, uint64_t m, int flags) {
  const bool exact = (flags & HX_RNGF_EXACT) != 0;
  return (flags & HX_RNGF_LEGACY) ? normal_at<T, true>(k, e, m, exact) : normal_at<T, false>(k, e, m, exact);
}
template <typename T>
__device__ __forceinline__ T uniform_rt(Key k, uint64_t e, uint64_t m, T lo, T span, int flags) {
  return (flags & HX_RNGF_LEGACY) ? uniform_at<T, true>(k, e, m, lo, span) : uniform_at<T, false>(k, e, m, lo, span);
}
__host__ __device__ __forceinline__ Key split_key_rt(Key k, uint64_t i, uint64_t num, int flags) {
  return (flags & 1) ? split_key<true>(k, i, num) : split_key<false>(k, i, num);
}

}  // namespace hx
