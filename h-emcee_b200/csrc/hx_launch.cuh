// hx_launch.cuh — tile-configuration dispatch for the fused half-move kernels.  The heavy template
// instantiations live in hx_walk_*.cu / hx_side_*.cu / hx_logprob.cu so they compile in parallel.
#pragma once
#include "../../include/hx.h"
#include "hx_err.h"
#include "hx_move.cuh"

namespace hx {

// (BM rows per CTA, TM rows per thread); threads = (BM/TM) * CGP, CG = ldd/4 column groups of 4.
template <typename T> struct Cfg;
template <> struct Cfg<float> {
  static constexpr int BMa = 32, TMa = 32;   // CG in [128, 256]
  static constexpr int BMb = 32, TMb = 16;   // CG in [32, 128)
  static constexpr int BMw = 8, TMw = 8;     // CG in (256, 1024]
};
template <> struct Cfg<double> {
  static constexpr int BMa = 16, TMa = 16;
  static constexpr int BMb = 16, TMb = 8;
  static constexpr int BMw = 4, TMw = 4;
};

// f.run<BM, TM, MAXT>() launches one instantiation; MAXT is the __launch_bounds__ of that variant.
template <typename T, class F>
static int dispatch_tile(int ldd, int rows, int sm_count, F&& f) {
  const int CG = ldd / 4;
  if (CG > 1024) return hx_fail(HX_E_UNSUPPORTED, "dim %d too large (max 4096)", ldd);
  const bool small = (rows + 31) / 32 < sm_count;   // few walkers: smaller row tiles -> more CTAs
  if (CG > 256) return f.template run<Cfg<T>::BMw, Cfg<T>::TMw, 1024>();
  if (CG >= 32) {
    if (small) return f.template run<8, 8, 256>();
    if (CG >= 128) return f.template run<Cfg<T>::BMa, Cfg<T>::TMa, 256>();
    return f.template run<Cfg<T>::BMb, Cfg<T>::TMb, 256>();
  }
  // narrow targets (dim < 128) are RNG/latency-bound: prefer the 8-row tile unless the 32-row tile already gives >= 16
  // resident warps per SM (4 warps per CTA)
  if (small || (long long)((rows + 31) / 32) * 4 < 16ll * sm_count) return f.template run<8, 1, 256>();
  return f.template run<32, 4, 256>();
}

// narrow-target momentum projection (hx_move.cuh k_project_narrow); returns HX_E_UNSUPPORTED if ldd > 64
template <typename T> int launch_project_narrow(Key key, const uint32_t* key_ptr, int legacy, const T* C, int n, int ldd, int d, int row0,
                                               int rows, T* S0, int sm_count, cudaStream_t s);
template <typename T> int launch_walk(const WalkParams<T>& P, int sm_count, cudaStream_t s);
template <typename T> int launch_side(const SideParams<T>& P, int sm_count, cudaStream_t s);
template <typename T> int launch_log_prob(const LogProbParams<T>& P, int sm_count, cudaStream_t s);

}  // namespace hx
