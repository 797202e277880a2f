// Micro-benchmark (diagnostic, not part of libhx): throughput of threefry2x32-20 variants that differ in how the rotations
// are issued, and of the individual integer instructions involved.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ uint32_t c_mul[32];

template <int V> __device__ __forceinline__ uint32_t rot(uint32_t x, int r) {
  if (V == 0) return __funnelshift_l(x, x, r);                                   // SHF.L.W
  if (V == 1) return (x * c_mul[r]) | (x >> (32 - r));                            // IMAD (lo) + SHF.R + (LOP3 merged)
  if (V == 2) { const uint64_t w = (uint64_t)x * (uint64_t)c_mul[r]; return (uint32_t)w | (uint32_t)(w >> 32); }   // IMAD.WIDE
  if (V == 3) return (x * c_mul[r]) | __umulhi(x, c_mul[r]);                      // IMAD + IMAD.HI
  if (V == 4) { if (r == 16) return __byte_perm(x, x, 0x1032); if (r == 24) return __byte_perm(x, x, 0x0321); return __funnelshift_l(x, x, r); }
  if (V == 5) return (x << r) | (x >> (32 - r));                                  // compiler's choice
  return 0;
}
template <int V> __device__ __forceinline__ uint32_t tf(uint32_t x0, uint32_t x1, uint32_t k0, uint32_t k1, uint32_t k2) {
#define R(r) x0 += x1; x1 = rot<V>(x1, r) ^ x0;
  R(13) R(15) R(26) R(6) x0 += k1; x1 += k2 + 1u;
  R(17) R(29) R(16) R(24) x0 += k2; x1 += k0 + 2u;
  R(13) R(15) R(26) R(6) x0 += k0; x1 += k1 + 3u;
  R(17) R(29) R(16) R(24) x0 += k1; x1 += k2 + 4u;
  R(13) R(15) R(26) R(6) x0 += k2; x1 += k0 + 5u;
#undef R
  return x0 ^ x1;
}
template <int V> __global__ void __launch_bounds__(256) k_hash(uint32_t k0, uint32_t k1, int iters, uint32_t* out) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  uint32_t acc = 0, base = (blockIdx.x * blockDim.x + threadIdx.x) * 8u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) acc ^= tf<V>(k0, base + t + k1, k0, k1, k2);
    base += gridDim.x * blockDim.x * 8u;
  }
  if (acc == 0x12345678u) out[0] = acc;
}
// single-instruction throughput: 8 independent chains per thread
template <int OP> __global__ void __launch_bounds__(256) k_op(uint32_t a, uint32_t b, int iters, uint32_t* out) {
  uint32_t x[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) x[t] = threadIdx.x + t + a;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        if (OP == 0) x[t] = __funnelshift_l(x[t], x[t], 13);                       // SHF.L.W
        if (OP == 1) x[t] = x[t] ^ b;                                              // LOP3
        if (OP == 2) x[t] = x[t] + b;                                              // IADD3 / IMAD
        if (OP == 3) x[t] = x[t] * c_mul[13];                                      // IMAD
        if (OP == 4) { const uint64_t w = (uint64_t)x[t] * (uint64_t)c_mul[13]; x[t] = (uint32_t)(w >> 32) ^ (uint32_t)w; }   // IMAD.WIDE + LOP3
        if (OP == 5) x[t] = __umulhi(x[t], c_mul[13]);                             // IMAD.HI
        if (OP == 6) x[t] = __byte_perm(x[t], b, 0x1032);                          // PRMT
        if (OP == 7) x[t] = x[t] >> 3;                                             // SHF.R
        if (OP == 8) x[t] = __uint_as_float(x[t]) * 1.0001f + 0.5f > 0.f ? x[t] + 1 : x[t];   // mixed
        if (OP == 9) x[t] = __float_as_uint(fmaf(__uint_as_float(x[t]), 1.0001f, 0.5f));       // FFMA
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) acc ^= x[t];
  if (acc == 0x12345678u) out[0] = acc;
}
template <typename F> static float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  uint32_t mul[32]; for (int i = 0; i < 32; ++i) mul[i] = 1u << i;
  cudaMemcpyToSymbol(c_mul, mul, sizeof(mul));
  uint32_t* out; cudaMalloc(&out, 4);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int grid = sms * 4, iters = 2000;
  const double hashes = (double)grid * 256 * 8 * iters;
  printf("SMs %d, clock %d kHz\n", sms, clk_khz);
#define HB(V, name) { float ms = timeit([&] { k_hash<V><<<grid, 256>>>(1u, 2u, iters, out); }); \
    printf("hash %-28s %8.3f ms  %.2f Ghash/s  %.1f clk/warp-hash/SMSP (at %.3f GHz)\n", name, ms, hashes / ms * 1e-6, ms * 1e-3 * clk_khz * 1e3 * sms * 4 / (hashes / 32), clk_khz * 1e-6); }
  HB(0, "SHF.L.W") HB(1, "IMAD + SHF.R") HB(2, "IMAD.WIDE") HB(3, "IMAD + IMAD.HI") HB(4, "PRMT for 16/24") HB(5, "shl|shr")
  const double ops = (double)grid * 256 * 64 * 500;
#define OB(OP, name) { float ms = timeit([&] { k_op<OP><<<grid, 256>>>(3u, 5u, 500, out); }); \
    printf("op   %-28s %8.3f ms  %.2f clk/warp-instr/SMSP\n", name, ms, ms * 1e-3 * clk_khz * 1e3 * sms * 4 / (ops / 32)); }
  OB(0, "SHF.L.W") OB(1, "LOP3") OB(2, "IADD") OB(3, "IMAD") OB(4, "IMAD.WIDE+LOP3") OB(5, "IMAD.HI") OB(6, "PRMT") OB(7, "SHF.R") OB(9, "FFMA")
  return 0;
}
