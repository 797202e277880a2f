// hx_walk_f32.cu — instantiations of the fused walk half-move kernel for float.
#include "hx_launch.cuh"
namespace hx {
template <typename T>
struct WalkLaunch {
  const WalkParams<T>& P;
  cudaStream_t s;
  template <int BM, int TM, int MAXT>
  int run() {
    k_walk_half<T, BM, TM, MAXT><<<(P.rows + BM - 1) / BM, tile_threads<BM, TM>(P.ldd), 0, s>>>(P);
    HX_LAUNCH_CHECK("k_walk_half");
    return 0;
  }
};
template <>
int launch_walk<float>(const WalkParams<float>& P, int sm_count, cudaStream_t s) {
  WalkLaunch<float> W{P, s};
  return dispatch_tile<float>(P.ldd, P.rows, sm_count, W);
}
}  // namespace hx
