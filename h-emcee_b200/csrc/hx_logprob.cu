// hx_logprob.cu — instantiations of the log-density / gradient kernel.
#include "hx_launch.cuh"
namespace hx {
template <typename T>
struct LogProbLaunch {
  const LogProbParams<T>& P;
  cudaStream_t s;
  template <int BM, int TM, int MAXT>
  int run() {
    k_log_prob<T, BM, TM, MAXT><<<(P.rows + BM - 1) / BM, tile_threads<BM, TM>(P.ldd), 0, s>>>(P);
    HX_LAUNCH_CHECK("k_log_prob");
    return 0;
  }
};
template <>
int launch_log_prob<float>(const LogProbParams<float>& P, int sm_count, cudaStream_t s) {
  LogProbLaunch<float> W{P, s};
  return dispatch_tile<float>(P.ldd, P.rows, sm_count, W);
}
template <>
int launch_log_prob<double>(const LogProbParams<double>& P, int sm_count, cudaStream_t s) {
  LogProbLaunch<double> W{P, s};
  return dispatch_tile<double>(P.ldd, P.rows, sm_count, W);
}
}  // namespace hx
