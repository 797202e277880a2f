// Micro-benchmark (diagnostic): do the alu pipe (SHF / LOP3 / IADD3) and the fma pipe (FFMA / IMAD) overlap on sm_100a?
// Each kernel runs 8 independent chains per thread of pattern P; reports clk per warp-instruction per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define SHF(x)   asm volatile("shf.l.wrap.b32 %0, %0, %0, 13;" : "+r"(x))
#define LOP(x,b) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(b))
#define ADD(x,b) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b))
#define FMA(y,a,b) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y) : "f"(a), "f"(b))
#define FMAI(y)  asm volatile("fma.rn.f32 %0, %0, 0f3F800347, 0f3F000000;" : "+f"(y))
#define MAD(x,a,b) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b))
template <int P> __global__ void __launch_bounds__(256) k_mix(uint32_t a, uint32_t b, float fa, float fb, int iters, uint32_t* out) {
  uint32_t x[8]; float y[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) { x[t] = threadIdx.x + t + a; y[t] = (float)(threadIdx.x + t) * fa; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        if (P == 0) { SHF(x[t]); SHF(x[t]); }
        if (P == 1) { FMA(y[t], fa, fb); FMA(y[t], fa, fb); }
        if (P == 2) { SHF(x[t]); FMA(y[t], fa, fb); }
        if (P == 3) { LOP(x[t], b); LOP(x[t], a); }
        if (P == 4) { LOP(x[t], b); FMA(y[t], fa, fb); }
        if (P == 5) { ADD(x[t], b); ADD(x[t], a); }
        if (P == 6) { SHF(x[t]); ADD(x[t], b); }
        if (P == 7) { MAD(x[t], a, b); MAD(x[t], b, a); }
        if (P == 8) { SHF(x[t]); MAD(x[t], a, b); }
        if (P == 9) { FMAI(y[t]); FMAI(y[t]); }
        if (P == 10) { SHF(x[t]); FMAI(y[t]); }
        if (P == 11) { ADD(x[t], b); SHF(x[t]); LOP(x[t], a); }      // one threefry round's instruction classes
        if (P == 12) { SHF(x[t]); LOP(x[t], a); FMA(y[t], fa, fb); FMA(y[t], fb, fa); }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) acc ^= x[t] ^ __float_as_uint(y[t]);
  if (acc == 0x12345678u) out[0] = acc;
}
int main() {
  uint32_t* out; cudaMalloc(&out, 4);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int grid = sms * 4, iters = 1000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"SHF SHF", "FFMA FFMA (3 reg)", "SHF FFMA", "LOP LOP", "LOP FFMA", "ADD ADD", "SHF ADD", "IMAD IMAD", "SHF IMAD", "FFMA FFMA (imm)", "SHF FFMA(imm)", "ADD SHF LOP (round)", "SHF LOP FFMA FFMA"};
  const int per[] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 4};
#define RUN(P) { float ms = 0; for (int r = 0; r < 2; ++r) { cudaEventRecord(e0); k_mix<P><<<grid, 256>>>(3u, 5u, 1.0001f, 0.5f, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); } \
    const double ins = (double)grid * 256 / 32 * 8 * 4 * per[P] * iters; \
    printf("%-24s %7.3f ms  %.2f clk per warp-instruction per SMSP (%.2f per pattern)\n", names[P], ms, ms * 1e-3 * clk_khz * 1e3 * sms * 4 / ins, ms * 1e-3 * clk_khz * 1e3 * sms * 4 / ins * per[P]); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
