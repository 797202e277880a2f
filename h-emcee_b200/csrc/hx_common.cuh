// hx_common.cuh — shared device helpers: vector loads, warp reductions, target descriptor view.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace hx {

constexpr int kWarp = 32;

template <typename T> __device__ __forceinline__ T inf_v();
template <> __device__ __forceinline__ float inf_v<float>() { return CUDART_INF_F; }
template <> __device__ __forceinline__ double inf_v<double>() { return CUDART_INF; }

// 4-wide aligned loads/stores (16 B for float, 2 x 16 B for double)
__device__ __forceinline__ void ld4(const float* p, float (&o)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
__device__ __forceinline__ void ld4(const double* p, double (&o)[4]) {
  const double2 a = reinterpret_cast<const double2*>(p)[0];
  const double2 b = reinterpret_cast<const double2*>(p)[1];
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
  reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
  reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}
__device__ __forceinline__ void ld4_ro(const float* p, float (&o)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
__device__ __forceinline__ void ld4_ro(const double* p, double (&o)[4]) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float exp_t(float a) { return expf(a); }
__device__ __forceinline__ double exp_t(double a) { return exp(a); }
__device__ __forceinline__ float log_t(float a) { return logf(a); }
__device__ __forceinline__ double log_t(double a) { return log(a); }
__device__ __forceinline__ float sqrt_t(float a) { return sqrtf(a); }
__device__ __forceinline__ double sqrt_t(double a) { return sqrt(a); }
__device__ __forceinline__ float min_t(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double min_t(double a, double b) { return fmin(a, b); }

// Device view of hx_target (include/hx.h) in the run dtype.
template <typename T>
struct TargetDev {
  int kind, dim, ld, has_lower0;
  const T* mean;  // [dim] or nullptr
  const T* prec;  // [dim, ld] (GAUSS_FULL)
  T a, b, lower0;
};

enum : int { kGaussIso = 0, kGaussFull = 1, kRosenbrock = 2, kFunnel = 3 };

}  // namespace hx
