// hx_err.h — error plumbing shared by the translation units of libhx.so.
#pragma once
#include <cuda_runtime.h>

// Records a message for hx_last_error() (thread-local) and returns `code`.
int hx_fail(int code, const char* fmt, ...);

#define HX_CUDA(expr)                                                                             \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess) return hx_fail((int)e_, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)
#define HX_LAUNCH_CHECK(name)                                                                        \
  do {                                                                                               \
    cudaError_t e_ = cudaGetLastError();                                                             \
    if (e_ != cudaSuccess) return hx_fail((int)e_, "launch %s failed: %s", name, cudaGetErrorString(e_)); \
  } while (0)
