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

template <int DV, int WPW>
static int project_narrow_launch(Key key, const uint32_t* key_ptr, int legacy, const float* C, int n, int ldd, int d, int row0, int rows,
                                 float* S0, cudaStream_t s) {
  const size_t smem = (size_t)64 * ((DV | 1) * 4) * sizeof(float);
  k_project_narrow<float, DV, WPW><<<(rows + 8 * WPW - 1) / (8 * WPW), 256, smem, s>>>(key, key_ptr, legacy, C, n, ldd, d, row0, rows, S0);
  HX_LAUNCH_CHECK("k_project_narrow");
  return 0;
}
template <int WPW>
static int project_narrow_dv(int dv, Key key, const uint32_t* key_ptr, int legacy, const float* C, int n, int ldd, int d, int row0, int rows,
                             float* S0, cudaStream_t s) {
#define HX_PN(DV) return project_narrow_launch<DV, WPW>(key, key_ptr, legacy, C, n, ldd, d, row0, rows, S0, s)
  if (dv <= 1) HX_PN(1);
  if (dv <= 2) HX_PN(2);
  if (dv <= 4) HX_PN(4);
  if (dv <= 7) HX_PN(7);
  if (dv <= 10) HX_PN(10);
  if (dv <= 13) HX_PN(13);
  HX_PN(16);
#undef HX_PN
}
template <>
int launch_project_narrow<float>(Key key, const uint32_t* key_ptr, int legacy, const float* C, int n, int ldd, int d, int row0, int rows,
                               float* S0, int sm_count, cudaStream_t s) {
  if (ldd > 64) return hx_fail(HX_E_UNSUPPORTED, "narrow projection needs dim <= 64");
  const int dv = ldd / 4;
  // more walkers per warp = more reuse of the shared complement tile and more independent draws per lane, as long as the grid
  // still fills the SMs and the WPW x dim partial sums fit in registers.  (The summation order per walker does not depend on it.)
  int wpw = 4;
  if (dv > 7 && wpw > 2) wpw = 2;
  if (false) wpw = 1;
  while (wpw > 1 && (rows + 8 * wpw - 1) / (8 * wpw) < 2 * sm_count) wpw >>= 1;
  if (wpw == 4) return project_narrow_dv<4>(dv, key, key_ptr, legacy, C, n, ldd, d, row0, rows, S0, s);
  if (wpw == 2) return project_narrow_dv<2>(dv, key, key_ptr, legacy, C, n, ldd, d, row0, rows, S0, s);
  return project_narrow_dv<1>(dv, key, key_ptr, legacy, C, n, ldd, d, row0, rows, S0, s);
}
}  // namespace hx
