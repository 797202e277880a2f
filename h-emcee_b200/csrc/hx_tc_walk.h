// hx_tc_walk.h — entry points of the tensor-core walk path (hx_tc_walk.cu) used by hx_api.cu.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/hx.h"
#include "hx_ctx.h"
#include "hx_rng.cuh"
#include "hx_dist.cuh"

namespace hx {
namespace tc {

struct FuseF32 {      // batch-level fused accept + chain emit (same fields as Fuse<float> in hx_api.cu)
  const uint32_t* akey_ptr;
  float* lp1w;
  int32_t* accepts;
  float* chain_out;
  float* chain_lp;
  int32_t* nonfinite;
  const DistPush* push;   // row-sharded runs: store the accepted rows into the peers' copies and publish the epoch (nullable)
};

struct NextDraw {     // the walk half-move that follows this one (its momentum draw is started early)
  Key key;
  const uint32_t* key_ptr;
  int row0, rows;
};

// true if this half-move should take the tcgen05 path under the ctx precision policy
bool walk_eligible(const hx_ctx* ctx, int dtype, const hx_target* tgt, int n, int rows, int flags);

// one walk half-move on tensor cores; `mean_dev` = mean_0(X2) [dim] already computed
int walk_half(hx_ctx* ctx, int impl, const hx_target* tgt, const float* mean_dev, const float* X1, const float* X2, int n,
              int row0, int rows, double eps, Key key, const uint32_t* key_ptr, int L, const float* lp1, float* Q,
              float* logacc, float* S, float* lp_out, const FuseF32* fuse, const NextDraw* next, cudaStream_t s);

// ChEES inner product beyond dim 128 (adaptation/chees.py:57-73): part[i * nt_out + t], t < nt_out, are the column-tile partial
// sums of <(pproj_i covinv), Xprop_i - mean_prop> for rows i < rows; covinv[d, d] (fp32, symmetric) = cov(group2)^-1.
// bf16 hi/lo operands, three products; overwrites the leapfrog operand workspaces (the half-move is finished by then).
int chees_inner(hx_ctx* ctx, const float* pproj, int rows, int d, const float* covinv, const float* Xprop, const float* mean_prop,
                float** part_out, int* nt_out, cudaStream_t s);

int microbench(hx_ctx* ctx, int mode, int n, int rows, int d, int reps, double* ms_out, cudaStream_t s);

}  // namespace tc
}  // namespace hx
