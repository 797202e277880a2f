// Micro-benchmark (diagnostic): rate of the fused kernel's draw arithmetic (threefry -> normal -> bf16 hi/lo split) alone, without
// any pipeline synchronisation, for several code shapes.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I../../h-emcee_b200/csrc -o normals_micro normals_micro.cu
#include <cstdio>
#include "hx_tc_fused.cuh"
using namespace hx; using namespace hx::tc;

__device__ __forceinline__ float central_poly(float w) {
  const float ww = w - 2.5f;
  float p = 2.81022636e-08f;
  p = fmaf(p, ww, 3.43273939e-07f); p = fmaf(p, ww, -3.5233877e-06f); p = fmaf(p, ww, -4.39150654e-06f);
  p = fmaf(p, ww, 0.00021858087f); p = fmaf(p, ww, -0.00125372503f); p = fmaf(p, ww, -0.00417768164f);
  p = fmaf(p, ww, 0.246640727f); p = fmaf(p, ww, 1.50140941f);
  return p;
}
__device__ __forceinline__ float tail_poly(float w) {
  const float ww = fast_sqrt(w) - 3.0f;
  float p = -0.000200214257f;
  p = fmaf(p, ww, 0.000100950558f); p = fmaf(p, ww, 0.00134934322f); p = fmaf(p, ww, -0.00367342844f);
  p = fmaf(p, ww, 0.00573950773f); p = fmaf(p, ww, -0.0076224613f); p = fmaf(p, ww, 0.00943887047f);
  p = fmaf(p, ww, 1.00167406f); p = fmaf(p, ww, 2.83297682f);
  return p;
}
__device__ __forceinline__ void uw(uint32_t b, float& u, float& w) {
  const uint32_t m = __umulhi(b, c_shr9mul) + 0x3F800000u;
  const float f = __uint_as_float(m) - 1.0f;
  u = fmaxf(-0.99999994f, fmaf(f, 2.0f, -0.99999994f));
  w = -0.693147182f * fast_lg2(__fsub_rn(1.0f, __fmul_rn(u, u)));
}
// V: 0 hash then float; 1 software-pipelined blocks; 2 sliced interleave (2 elements per slice); 3 = 1 + branch-free tail;
//    4 = 2 + branch-free tail; 5 = 1 + integer pre-filter on the random bits + vote-guarded fix; 6 = 2 + (5)'s tail handling
template <int V> __global__ void __launch_bounds__(1024, 1) k_draw(Key key, int items, uint32_t* out) {
  const uint32_t ks2 = key.k0 ^ key.k1 ^ 0x1BD11BDAu;
  uint64_t e = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8ull;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 8ull;
  uint32_t acc = 0;
  uint32_t bcur[8];
  if (V >= 1) bits8_fast(key, ks2, e, bcur);
  for (int it = 0; it < items; ++it) {
    float x[8], u[8], w[8];
    uint4 hi, lo;
    uint32_t bnext[8];
    if (V == 0) {
      bits8_fast(key, ks2, e, bcur);
      const bool tail = normals8_central(bcur, x, u, w);
      if (tail) normals8_fix_tail(x, u, w);
    } else if (V == 1) {
      bits8_fast(key, ks2, e + stride, bnext);
      const bool tail = normals8_central(bcur, x, u, w);
      if (tail) normals8_fix_tail(x, u, w);
    } else if (V == 3) {
      bits8_fast(key, ks2, e + stride, bnext);
#pragma unroll
      for (int t = 0; t < 8; ++t) { uw(bcur[t], u[t], w[t]); const float pc = central_poly(w[t]), pt = tail_poly(w[t]); x[t] = 1.41421354f * ((w[t] < 5.0f ? pc : pt) * u[t]); }
    } else if (V == 2 || V == 4 || V == 6) {
      const uint64_t en = e + stride;
      const uint32_t a0 = (uint32_t)(en >> 32) + key.k0, a1 = (uint32_t)en + key.k1;
      uint32_t flags = 0;
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        bnext[2 * sl] = tf_bits(a0, a1 + 2 * sl, key.k0, key.k1, ks2);
        bnext[2 * sl + 1] = tf_bits(a0, a1 + 2 * sl + 1, key.k0, key.k1, ks2);
#pragma unroll
        for (int t = 2 * sl; t < 2 * sl + 2; ++t) {
          uw(bcur[t], u[t], w[t]);
          if (V == 4) { const float pc = central_poly(w[t]), pt = tail_poly(w[t]); x[t] = 1.41421354f * ((w[t] < 5.0f ? pc : pt) * u[t]); }
          else { x[t] = 1.41421354f * (central_poly(w[t]) * u[t]); }
          if (V == 6) flags |= 0x7FD000u - (bcur[t] >> 9);        // placeholder pre-filter arithmetic (one subtract + OR per element)
        }
      }
      if (V == 2) { bool tail = false;
#pragma unroll
        for (int t = 0; t < 8; ++t) tail = tail || !(w[t] < 5.0f);
        if (tail) normals8_fix_tail(x, u, w); }
      if (V == 6) { if (__any_sync(0xffffffffu, (int)flags < 0)) normals8_fix_tail(x, u, w); }
    } else if (V == 5) {
      bits8_fast(key, ks2, e + stride, bnext);
      uint32_t flags = 0;
#pragma unroll
      for (int t = 0; t < 8; ++t) { uw(bcur[t], u[t], w[t]); x[t] = 1.41421354f * (central_poly(w[t]) * u[t]); flags |= 0x7FD000u - (bcur[t] >> 9); }
      if (__any_sync(0xffffffffu, (int)flags < 0)) normals8_fix_tail(x, u, w);
    }
    if (V >= 1) {
#pragma unroll
      for (int t = 0; t < 8; ++t) bcur[t] = bnext[t];
    }
    split8_bf16(x, hi, lo);
    acc ^= hi.x ^ hi.y ^ hi.z ^ hi.w ^ lo.x ^ lo.y ^ lo.z ^ lo.w;
    e += stride;
  }
  if (acc == 0x12345678u) out[0] = acc;
}
template <int V> static void run(int sms, int clk_khz, uint32_t* out) {
  const int items = 1000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps : {8, 16, 28}) {
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0); k_draw<V><<<sms, warps * 32>>>(Key{1u, 2u}, items, out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    const double normals = (double)sms * warps * 32 * 8 * items;
    printf("variant %d, %2d warps/SM: %7.3f ms, %6.1f clk per warp-normal per SMSP, 1.07e9 normals in %.2f ms\n", V, warps, ms,
           ms * 1e-3 * clk_khz * 1e3 * sms * 4 / (normals / 32), ms * 1.073741824e9 / normals);
  }
}
int main() {
  uint32_t* out; cudaMalloc(&out, 4);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  run<0>(sms, clk_khz, out); run<1>(sms, clk_khz, out); run<2>(sms, clk_khz, out); run<3>(sms, clk_khz, out);
  run<4>(sms, clk_khz, out); run<5>(sms, clk_khz, out); run<6>(sms, clk_khz, out);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
