// hx_ctx.h — the per-device context (private scratch workspace + measurement hooks) shared by the
// translation units of libhx.so.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include <cuda_runtime.h>
#include "hx_err.h"
#include "../../include/hx.h"

enum Slot {
  WS_C = 0, WS_M, WS_MEAN, WS_PART, WS_G, WS_S, WS_Q, WS_LA, WS_LP, WS_IDX, WS_GRAMP,
  // tensor-core path
  WS_TC_CT, WS_TC_MP, WS_TC_PP, WS_TC_P0, WS_TC_QC, WS_TC_GP, WS_TC_LPP, WS_TC_DKP, WS_TC_MASK, WS_TC_BO, WS_TC_S0,
  // propagator form of the Gaussian leapfrog
  WS_TC_BX, WS_TC_BS, WS_TC_TQ, WS_TC_TS, WS_TC_TW, WS_TC_WP, WS_TC_B1, WS_TC_B2, WS_TC_ZB, WS_TC_ZT, WS_TC_QC2, WS_TC_PP32, WS_TC_S0P, WS_TC_CT32, WS_TC_CT8, WS_TC_SC8, WS_TC_AM, WS_TC_RSC, WS_TC_RSB, WS_TC_CTR,
  // device-resident adaptation
  WS_AD_STATE, WS_AD_MEAN, WS_AD_PART, WS_AD_CHOL, WS_AD_STAT, WS_AD_PROP,
  WS_NSLOT
};

// one momentum draw P0 = normal(key, (n, n))[row0 : row0 + rows] held (or being written) in a draw buffer
struct P0Job {
  int valid, fmt, impl, n, row0, rows;       // fmt: 0 bf16 hi/lo planes, 1 fp16 + e4m3 planes
  uint32_t k0, k1;
  const uint32_t* key_ptr;       // device-resident key (batch loop) or NULL (key by value)
};

// row-sharded multi-GPU state (hx_dist.cuh)
struct HxDist {
  int on, rank, world, n, dim, dtype;
  void* base[8];          // own region + peers' regions mapped through CUDA IPC
  size_t bytes, lp_off, flag_off, counter_off;
  uint32_t epoch;         // half-moves pushed so far
};

// cached CUDA graphs of launch-bound batch loops (small ensembles): one entry per distinct argument set
struct GraphEntry {
  uint64_t key;
  uint64_t ws_gen;           // ctx->ws_generation when the graph was captured (it bakes in workspace addresses)
  cudaGraphExec_t exec;      // nullptr: the argument set was seen once (run eagerly, workspaces primed), capture on the next use
  uint64_t last_use;
};

struct hx_ctx {
  int device;
  void* buf[WS_NSLOT];
  size_t cap[WS_NSLOT];
  int sm_count;
  // measurement hooks
  int timing;
  std::vector<cudaEvent_t>* ev;   // marks on the launch stream; the interval after mark i belongs to phase ev_tag[i]
  std::vector<int>* ev_tag;       // HX_PH_* or -1 (end of a measured section)
  size_t ev_used;
  double phase_ms[8]; int64_t phase_n[8];   // last hx_ctx_timing_read, by phase
  int64_t launches;               // all kernels launched through this ctx
  int nvtx_open;                  // an NVTX phase range is open on this host thread (timing_mark)
  int last_gram_slices;
  int precision;                  // HX_PREC_*
  int last_path;                  // 0 fused SIMT half-move, 1 tcgen05 projection GEMM
  void* encode_fn;                // cuTensorMapEncodeTiled (resolved lazily)
  // side stream + events used to overlap the momentum draw (ALU) with the projection GEMM (tensor pipe)
  cudaStream_t side;
  cudaEvent_t ev_fork;
  int side_ready;
  P0Job job[2];                            // contents of the two draw buffers
  std::vector<cudaEvent_t>* ev_chunk[2];   // per buffer: "chunk c drawn" events
  void* p0_base; size_t p0_elems;          // identity of the draw workspace the jobs refer to
  P0Job hint; int hint_valid;              // hx_walk_prefetch: the walk half-move after the next one
  int leap_form;                           // HX_LEAP_*
  HxDist dist;
  std::vector<GraphEntry>* graphs; cudaStream_t gstream; uint64_t graph_tick; int graphs_off;
  uint64_t ws_generation;                  // bumped whenever ws_get moves a workspace slot: captured graphs from before are stale
  void* key_stage; size_t key_stage_cap;    // fixed-address copy of a batch's keys (graph replays read the keys from here)
  hx_chain_stats stats; int stats_on;      // on-device chain statistics fed by the batch loops after every iteration (hx_stats.cu)
  int chain_every, chain_phase;            // thinned chain emit of the batch loops (hx_ctx_set_chain_thinning); 0/1 = every iteration
  const void* ctl_eps; const int* ctl_L;   // device-resident (step, L) the SIMT move kernels read during adaptive warm-up
  int narrow_off;                          // 1: keep the momentum projection of narrow targets inside the half-move kernel
  int energy_bf16;                         // propagator-form energy products: 0 tf32 hi/lo (default), 1 bf16 hi/lo, 2 fp16 hi/lo (scaled)
  int gram_tf32;                           // 1: Gram from tf32 hi/lo operands instead of bf16 hi/lo
  int proj_slices;                         // split-K slices of the projection GEMM (0: default)
  int rng_ctas_per_sm;                     // resident CTAs/SM of the draw kernel (0: default 4)
  int draw_fork;                           // next half-move's draw is queued after this many projection chunks (0: at the start)
  int proj_pair;                           // 1: momentum projection by the CTA-pair (cta_group::2) GEMM kernel
  int pair_chunked;                        // 1: CTA-pair projection launched one wave at a time even when the draw is complete
  int draw_nostore;                        // diagnostics: momentum draw without its stores (results are garbage)
  int aux_pair;                            // 1: Gram and propagator-apply GEMMs by the CTA-pair kernel
  int proj_one_launch;                     // 1: single-CTA projection as ONE launch over all rows when the draw was queued ahead
  int kick_f16;                            // kicks / propagator apply from fp16 hi/lo operands: 0 auto (stepwise form only), 1 always, 2 never
  // fused draw + projection kernel (hx_tc_fused.cuh)
  int proj_fused;                          // 1 (default): P0 is drawn inside the projection GEMM when eligible; 0: staged draw + GEMM
  int fused_nw;                            // RNG warps per CTA: 8 or 16 (0: default)
  int fused_slices;                        // split-K slices of the fused projection (0: default = 2 from 4096 walkers per group)
  int fused_relaxed;                       // 1: RNG warps wait for a free A stage with test_wait + nanosleep (0: parked try_wait)
  int fused_fmt;                           // operand form of the fused projection: 0 by precision (AUTO: fp16 + fp8, BF16X3: bf16 hi/lo), 1 bf16 hi/lo, 2 fp16 + fp8
  int row_epilogue;                        // 1: row-per-lane GEMM epilogues (diagnostics; default: coalesced through shared memory)
  int fused_stagger_ns;                    // start-up stagger of the fused kernel's RNG warps per K-block position, in ns
  int fused_no_atomic;                     // 1: two split-K slices of the fused projection go through slice buffers + a sum kernel instead of red.add
  int fused_diag;                          // diagnostics: 1 no draw, 2 no MMA (results are garbage)
  int one_acc;                             // staged bf16 hi/lo projection accumulates all three products in one accumulator (bit-compare with the fused kernel)
  std::vector<const void*>* attr_done;     // kernels whose dynamic-shared-memory attribute was set on THIS ctx's device (attributes are per device)
};

namespace hx { int stats_feed(hx_ctx* ctx, int dtype, const hx_chain_stats* st, const void* X, cudaStream_t s); }
// cudaFuncAttributeMaxDynamicSharedMemorySize once per (ctx = device, kernel)
int ensure_smem_attr(hx_ctx* ctx, const void* func, int bytes);
// grow-only scratch slot
int ws_get(hx_ctx* ctx, int slot, size_t bytes, void** out);
// phases of a half-move for the timing hooks (hx_ctx_timing_read / hx_ctx_timing_phases)
enum { HX_PH_MAIN = 0, HX_PH_PREP = 1, HX_PH_BASIS = 2, HX_PH_LEAP = 3, HX_PH_FINISH = 4, HX_PH_WAIT = 5, HX_PH_PUSH = 6, HX_PH_STATS = 7, HX_PH_COUNT = 8 };
int timing_mark(hx_ctx* ctx, cudaStream_t s, int phase);   // phase -1 closes the section
inline int timing_begin(hx_ctx* ctx, cudaStream_t s) { return timing_mark(ctx, s, HX_PH_MAIN); }
inline int timing_end(hx_ctx* ctx, cudaStream_t s) { return timing_mark(ctx, s, -1); }
