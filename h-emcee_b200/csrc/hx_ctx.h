// hx_ctx.h — the per-device context (private scratch workspace + measurement hooks) shared by the
// translation units of libhx.so.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include <cuda_runtime.h>
#include "hx_err.h"

enum Slot {
  WS_C = 0, WS_M, WS_MEAN, WS_PART, WS_G, WS_S, WS_Q, WS_LA, WS_LP, WS_IDX, WS_GRAMP,
  // tensor-core path
  WS_TC_CT, WS_TC_MP, WS_TC_PP, WS_TC_P0, WS_TC_QC, WS_TC_GP, WS_TC_LPP, WS_TC_DKP, WS_TC_MASK, WS_TC_BO, WS_TC_S0,
  WS_NSLOT
};

struct hx_ctx {
  int device;
  void* buf[WS_NSLOT];
  size_t cap[WS_NSLOT];
  int sm_count;
  // measurement hooks
  int timing;
  std::vector<cudaEvent_t>* ev;   // pairs (start, stop) around the dominant kernel
  size_t ev_used;
  int64_t launches;               // all kernels launched through this ctx
  int last_gram_slices;
  int precision;                  // HX_PREC_*
  int last_path;                  // 0 fused SIMT half-move, 1 tcgen05 projection GEMM
  void* encode_fn;                // cuTensorMapEncodeTiled (resolved lazily)
  // side stream + events used to overlap the momentum draw (ALU) with the projection GEMM (tensor pipe)
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_rng[2], ev_gemm[2];
  int side_ready;
};

// grow-only scratch slot
int ws_get(hx_ctx* ctx, int slot, size_t bytes, void** out);
int timing_begin(hx_ctx* ctx, cudaStream_t s);
int timing_end(hx_ctx* ctx, cudaStream_t s);
