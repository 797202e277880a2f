// hx_api.cu — C-ABI of libhx.so (include/hx.h): argument validation, private workspace, kernel
// dispatch.  No torch types, no exceptions across the boundary.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <new>
#include <vector>
#include <nvtx3/nvToolsExt.h>

#include "../../include/hx.h"
#include "hx_common.cuh"
#include "hx_rng.cuh"
#include "hx_prep.cuh"
#include "hx_launch.cuh"
#include "hx_ctx.h"
#include "hx_tc_walk.h"
#include "hx_dist.cuh"
#include "hx_adapt.cuh"
#include "hx_spd.cuh"
#include "hx_vanilla.cuh"
#include <type_traits>

using namespace hx;

// ---- error plumbing ---------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int hx_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define fail hx_fail

// ---- ctx / workspace ---------------------------------------------------------------------------------
int ws_get(hx_ctx* ctx, int slot, size_t bytes, void** out) {
  if (bytes == 0) bytes = 16;
  if (ctx->cap[slot] < bytes) {
    if (ctx->buf[slot]) HX_CUDA(cudaFree(ctx->buf[slot]));
    ctx->buf[slot] = nullptr;
    ctx->cap[slot] = 0;
    size_t want = bytes + bytes / 4;
    HX_CUDA(cudaMalloc(&ctx->buf[slot], want));
    ctx->cap[slot] = want;
    ctx->ws_generation += 1;        // cached CUDA graphs hold the old address (run_maybe_graphed re-captures)
  }
  *out = ctx->buf[slot];
  return 0;
}

int ensure_smem_attr(hx_ctx* ctx, const void* func, int bytes) {
  if (!ctx->attr_done) ctx->attr_done = new (std::nothrow) std::vector<const void*>();
  if (ctx->attr_done)
    for (const void* f : *ctx->attr_done) if (f == func) return 0;
  HX_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (ctx->attr_done) ctx->attr_done->push_back(func);
  return 0;
}

// Phase boundary of a half-move on the launch stream: an NVTX range per phase (nsys / ncu --nvtx attribute the launches enqueued
// inside it; header-only NVTX3, a no-op unless a tool is attached) and, when hx_ctx_set_timing is on, a CUDA event.
static const char* const kPhaseName[HX_PH_COUNT] = {"hx:project(draw+GEMM)", "hx:prep(mean,C^T,Gram,MP)", "hx:propagator-basis", "hx:leapfrog+energies",
                                                    "hx:finish(accept,emit)", "hx:wait-peers", "hx:push-rows", "hx:stats"};
int timing_mark(hx_ctx* ctx, cudaStream_t s, int phase) {
  if (ctx->nvtx_open) { nvtxRangePop(); ctx->nvtx_open = 0; }
  if (phase >= 0 && phase < HX_PH_COUNT) { nvtxRangePushA(kPhaseName[phase]); ctx->nvtx_open = 1; }
  if (!ctx->timing) return 0;
  if (!ctx->ev) { ctx->ev = new (std::nothrow) std::vector<cudaEvent_t>(); ctx->ev_tag = new (std::nothrow) std::vector<int>(); }
  if (!ctx->ev || !ctx->ev_tag) return hx_fail(HX_E_NULL, "out of host memory");
  if (ctx->ev->size() <= ctx->ev_used) {
    cudaEvent_t e;
    HX_CUDA(cudaEventCreate(&e));
    ctx->ev->push_back(e);
    ctx->ev_tag->push_back(-1);
  }
  HX_CUDA(cudaEventRecord((*ctx->ev)[ctx->ev_used], s));
  (*ctx->ev_tag)[ctx->ev_used] = phase;
  ctx->ev_used += 1;
  return 0;
}

extern "C" HX_API int hx_ctx_set_precision(hx_ctx* ctx, int mode) {
  if (!ctx) return hx_fail(HX_E_NULL, "hx_ctx_set_precision: NULL ctx");
  if (mode < HX_PREC_AUTO || mode > HX_PREC_F16F8) return hx_fail(HX_E_DTYPE, "hx_ctx_set_precision: unknown mode %d", mode);
  ctx->precision = mode;
  return 0;
}
extern "C" HX_API int hx_ctx_set_leapfrog_form(hx_ctx* ctx, int form) {
  if (!ctx) return hx_fail(HX_E_NULL, "hx_ctx_set_leapfrog_form: NULL ctx");
  if (form < HX_LEAP_AUTO || form > HX_LEAP_PROPAGATOR) return hx_fail(HX_E_DTYPE, "hx_ctx_set_leapfrog_form: unknown form %d", form);
  ctx->leap_form = form;
  return 0;
}
extern "C" HX_API int hx_walk_prefetch(hx_ctx* ctx, int impl, int32_t n, int32_t row0, int32_t rows, const uint32_t key[2]) {
  if (!ctx || !key) return hx_fail(HX_E_NULL, "hx_walk_prefetch: NULL ctx/key");
  if (impl != HX_RNG_PARTITIONABLE && impl != HX_RNG_LEGACY) return hx_fail(HX_E_DTYPE, "hx_walk_prefetch: unknown rng impl %d", impl);
  if (n < 2 || row0 < 0 || rows < 1 || row0 + rows > n) return hx_fail(HX_E_SHAPE, "hx_walk_prefetch: bad row shard");
  ctx->hint.valid = 1; ctx->hint.impl = impl; ctx->hint.n = n; ctx->hint.row0 = row0; ctx->hint.rows = rows;
  ctx->hint.k0 = key[0]; ctx->hint.k1 = key[1]; ctx->hint.key_ptr = nullptr;
  ctx->hint_valid = 1;
  return 0;
}
extern "C" HX_API int hx_ctx_set_tuning(hx_ctx* ctx, int what, int value) {
  if (!ctx) return hx_fail(HX_E_NULL, "hx_ctx_set_tuning: NULL ctx");
  if (what == 0) ctx->rng_ctas_per_sm = value;
  else if (what == 1) ctx->proj_slices = value;
  else if (what == 2) ctx->gram_tf32 = value;
  else if (what == 3) ctx->energy_bf16 = value;
  else if (what == 4) ctx->graphs_off = value ? 0 : 1;
  else if (what == 5) ctx->narrow_off = value ? 0 : 1;
  else if (what == 6) ctx->draw_fork = value;
  else if (what == 8) ctx->proj_pair = value ? 1 : 0;
  else if (what == 9) ctx->pair_chunked = value ? 1 : 0;
  else if (what == 10) ctx->draw_nostore = value ? 1 : 0;
  else if (what == 12) ctx->aux_pair = value;
  else if (what == 13) ctx->proj_one_launch = value ? 1 : 0;
  else if (what == 14) ctx->kick_f16 = value;
  else if (what == 15) ctx->proj_fused = value ? 1 : 0;
  else if (what == 16) ctx->fused_nw = value;
  else if (what == 17) ctx->fused_slices = value;
  else if (what == 18) ctx->one_acc = value ? 1 : 0;
  else if (what == 19) ctx->fused_diag = value;
  else if (what == 20) ctx->fused_relaxed = value ? 1 : 0;
  else if (what == 21) ctx->fused_fmt = value;
  else if (what == 22) ctx->fused_no_atomic = value ? 1 : 0;
  else if (what == 23) ctx->row_epilogue = value;
  else if (what == 24) ctx->fused_stagger_ns = value;
  else return hx_fail(HX_E_DTYPE, "hx_ctx_set_tuning: unknown knob %d", what);
  return 0;
}
extern "C" HX_API int hx_tc_microbench(hx_ctx* ctx, int mode, int32_t n, int32_t rows, int32_t dim, int32_t reps, double* ms_out, void* stream) {
  if (!ctx || !ms_out) return hx_fail(HX_E_NULL, "hx_tc_microbench: NULL argument");
  if (mode < 0 || mode > 3 || n < 64 || rows < 1 || dim < 16 || reps < 1) return hx_fail(HX_E_SHAPE, "hx_tc_microbench: bad argument");
  HX_CUDA(cudaSetDevice(ctx->device));
  return tc::microbench(ctx, mode, n, rows, dim, reps, ms_out, (cudaStream_t)stream);
}
extern "C" HX_API int hx_ctx_last_path(const hx_ctx* ctx) { return ctx ? ctx->last_path : -1; }

extern "C" HX_API int hx_ctx_timing_enable(hx_ctx* ctx, int enable) {
  if (!ctx) return hx_fail(HX_E_NULL, "hx_ctx_timing_enable: NULL ctx");
  ctx->timing = enable ? 1 : 0;
  ctx->ev_used = 0;
  ctx->launches = 0;
  return 0;
}

static int timing_collect(hx_ctx* ctx, double* phase_ms, int64_t* phase_n) {
  for (int p = 0; p < HX_PH_COUNT; ++p) { phase_ms[p] = 0.0; phase_n[p] = 0; }
  for (size_t i = 0; i + 1 < ctx->ev_used; ++i) {
    const int tag = (*ctx->ev_tag)[i];
    if (tag < 0 || tag >= HX_PH_COUNT) continue;
    HX_CUDA(cudaEventSynchronize((*ctx->ev)[i + 1]));
    float t = 0.f;
    HX_CUDA(cudaEventElapsedTime(&t, (*ctx->ev)[i], (*ctx->ev)[i + 1]));
    phase_ms[tag] += t;
    phase_n[tag] += 1;
  }
  return 0;
}

extern "C" HX_API int hx_ctx_timing_read(hx_ctx* ctx, double* kernel_ms, int64_t* kernel_launches, int64_t* all_launches) {
  if (!ctx) return hx_fail(HX_E_NULL, "hx_ctx_timing_read: NULL ctx");
  double pm[HX_PH_COUNT];
  int64_t pn[HX_PH_COUNT];
  if (int e = timing_collect(ctx, pm, pn)) return e;
  for (int p = 0; p < HX_PH_COUNT; ++p) { ctx->phase_ms[p] = pm[p]; ctx->phase_n[p] = pn[p]; }
  if (kernel_ms) *kernel_ms = pm[HX_PH_MAIN];
  if (kernel_launches) *kernel_launches = pn[HX_PH_MAIN];
  if (all_launches) *all_launches = ctx->launches;
  ctx->ev_used = 0;
  ctx->launches = 0;
  return 0;
}

extern "C" HX_API int hx_ctx_timing_phases(hx_ctx* ctx, double* phase_ms, int64_t* phase_count) {
  if (!ctx || !phase_ms || !phase_count) return hx_fail(HX_E_NULL, "hx_ctx_timing_phases: NULL argument");
  for (int p = 0; p < HX_PH_COUNT; ++p) { phase_ms[p] = ctx->phase_ms[p]; phase_count[p] = ctx->phase_n[p]; }
  return 0;
}

extern "C" HX_API int hx_abi_version(void) { return HX_ABI_VERSION; }
extern "C" HX_API const char* hx_last_error(void) { return g_err; }

extern "C" HX_API int hx_ctx_create(hx_ctx** out, int device) {
  if (!out) return fail(HX_E_NULL, "hx_ctx_create: out is NULL");
  int ndev = 0;
  HX_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(HX_E_SHAPE, "hx_ctx_create: device %d out of range (%d)", device, ndev);
  HX_CUDA(cudaSetDevice(device));
  hx_ctx* c = new (std::nothrow) hx_ctx();
  if (!c) return fail(HX_E_NULL, "hx_ctx_create: out of host memory");
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->aux_pair = 1;
  c->energy_bf16 = 2;
  c->proj_fused = 1;
  c->row_epilogue = (1 << 2) | (1 << 3);      // EPI_DOT, EPI_PROP: row-per-lane epilogue (measured: coalescing costs 0.4 ms / nothing there); STORE, KICKB: coalesced

  cudaDeviceProp prop;
  HX_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) {
    delete c;
    return fail(HX_E_UNSUPPORTED, "libhx is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  }
  *out = c;
  return 0;
}

extern "C" HX_API int hx_dist_destroy(hx_ctx* ctx);
extern "C" HX_API int hx_ctx_destroy(hx_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  hx_dist_destroy(ctx);
  for (int i = 0; i < WS_NSLOT; ++i)
    if (ctx->buf[i]) cudaFree(ctx->buf[i]);
  if (ctx->ev) {
    for (cudaEvent_t e : *ctx->ev) cudaEventDestroy(e);
    delete ctx->ev;
    delete ctx->ev_tag;
  }
  if (ctx->graphs) {
    for (GraphEntry& g : *ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    delete ctx->graphs;
  }
  delete ctx->attr_done;
  if (ctx->gstream) cudaStreamDestroy(ctx->gstream);
  if (ctx->key_stage) cudaFree(ctx->key_stage);
  {
  }
  if (ctx->side_ready) {
    cudaStreamDestroy(ctx->side);
    cudaEventDestroy(ctx->ev_fork);
    for (int b = 0; b < 2; ++b) {
      if (!ctx->ev_chunk[b]) continue;
      for (cudaEvent_t e : *ctx->ev_chunk[b]) cudaEventDestroy(e);
      delete ctx->ev_chunk[b];
    }
  }
  delete ctx;
  return 0;
}

extern "C" HX_API size_t hx_ctx_workspace_bytes(const hx_ctx* ctx) {
  size_t s = 0;
  if (ctx) for (int i = 0; i < WS_NSLOT; ++i) s += ctx->cap[i];
  return s;
}

// ---- helpers ---------------------------------------------------------------------------------------
static inline int grid_for(uint64_t n, int block, int cap = 148 * 16) {
  uint64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > (uint64_t)cap) g = cap;
  return (int)g;
}
static inline int round4(int d) { return (d + 3) & ~3; }
// RNG flags word of the kernels (hx_rng.cuh HX_RNGF_*): counter layout + the exact f32 normal form under HX_PREC_EXACT
static inline int rng_flags(const hx_ctx* ctx, int impl) {
  return (impl == HX_RNG_LEGACY ? HX_RNGF_LEGACY : 0) | (ctx->precision == HX_PREC_EXACT ? HX_RNGF_EXACT : 0);
}

static inline int check_dtype_impl(int dtype, int impl, const char* who) {
  if (dtype != HX_F32 && dtype != HX_F64) return fail(HX_E_DTYPE, "%s: unknown dtype %d", who, dtype);
  if (impl != HX_RNG_PARTITIONABLE && impl != HX_RNG_LEGACY) return fail(HX_E_DTYPE, "%s: unknown rng impl %d", who, impl);
  return 0;
}

template <typename T>
static int make_target(const hx_target* t, TargetDev<T>* o, const char* who) {
  if (!t) return fail(HX_E_NULL, "%s: target is NULL", who);
  if (t->dim < 1) return fail(HX_E_TARGET, "%s: target dim %d < 1", who, t->dim);
  if (t->kind < HX_TARGET_GAUSS_ISO || t->kind > HX_TARGET_FUNNEL) return fail(HX_E_TARGET, "%s: unknown target kind %d", who, t->kind);
  if (t->kind == HX_TARGET_GAUSS_FULL) {
    if (!t->precision) return fail(HX_E_TARGET, "%s: GAUSS_FULL needs a precision matrix", who);
    if (t->ld < t->dim || (t->ld & 3)) return fail(HX_E_TARGET, "%s: precision ld %d must be a multiple of 4 and >= dim %d", who, t->ld, t->dim);
    if (((uintptr_t)t->precision) & 15) return fail(HX_E_ALIGN, "%s: precision must be 16-byte aligned", who);
  }
  if (t->kind == HX_TARGET_ROSENBROCK && t->dim < 2) return fail(HX_E_TARGET, "%s: rosenbrock needs dim >= 2", who);
  if (t->kind == HX_TARGET_FUNNEL && !(t->a > 0)) return fail(HX_E_TARGET, "%s: funnel sigma must be > 0", who);
  o->kind = t->kind; o->dim = t->dim; o->ld = t->ld; o->has_lower0 = t->has_lower0;
  o->mean = (const T*)t->mean; o->prec = (const T*)t->precision;
  o->a = (T)t->a; o->b = (T)t->b; o->lower0 = (T)t->lower0;
  return 0;
}

// ---- (1) RNG entry points ----------------------------------------------------------------------------
extern "C" HX_API int hx_split(const uint32_t key[2], int impl, int64_t num, uint32_t* out_dev, void* stream) {
  if (!key || !out_dev) return fail(HX_E_NULL, "hx_split: NULL argument");
  if (num < 0) return fail(HX_E_SHAPE, "hx_split: num %lld < 0", (long long)num);
  if (int e = check_dtype_impl(HX_F32, impl, "hx_split")) return e;
  if (impl == HX_RNG_LEGACY && (uint64_t)num * 2 > 0xFFFFFFFFull) return fail(HX_E_SHAPE, "hx_split: legacy split supports 2*num < 2^32");
  if (num == 0) return 0;
  Key k{key[0], key[1]};
  cudaStream_t s = (cudaStream_t)stream;
  if (impl == HX_RNG_LEGACY) k_split<true><<<grid_for(num, 256), 256, 0, s>>>(k, (uint64_t)num, out_dev);
  else k_split<false><<<grid_for(num, 256), 256, 0, s>>>(k, (uint64_t)num, out_dev);
  HX_LAUNCH_CHECK("k_split");
  return 0;
}

extern "C" HX_API int hx_random_bits(const uint32_t key[2], int impl, int bit_width, int64_t count, void* out_dev, void* stream) {
  if (!key || !out_dev) return fail(HX_E_NULL, "hx_random_bits: NULL argument");
  if (count < 0) return fail(HX_E_SHAPE, "hx_random_bits: count < 0");
  if (bit_width != 32 && bit_width != 64) return fail(HX_E_DTYPE, "hx_random_bits: bit_width must be 32 or 64");
  if (int e = check_dtype_impl(HX_F32, impl, "hx_random_bits")) return e;
  if (impl == HX_RNG_LEGACY && (uint64_t)count * (bit_width / 32) > 0xFFFFFFFFull)
    return fail(HX_E_SHAPE, "hx_random_bits: legacy layout supports < 2^32 words");
  if (count == 0) return 0;
  Key k{key[0], key[1]};
  cudaStream_t s = (cudaStream_t)stream;
  const int g = grid_for(count, 256);
  if (bit_width == 32) {
    if (impl == HX_RNG_LEGACY) k_bits32<true><<<g, 256, 0, s>>>(k, (uint64_t)count, (uint32_t*)out_dev);
    else k_bits32<false><<<g, 256, 0, s>>>(k, (uint64_t)count, (uint32_t*)out_dev);
  } else {
    if (impl == HX_RNG_LEGACY) k_bits64<true><<<g, 256, 0, s>>>(k, (uint64_t)count, (uint64_t*)out_dev);
    else k_bits64<false><<<g, 256, 0, s>>>(k, (uint64_t)count, (uint64_t*)out_dev);
  }
  HX_LAUNCH_CHECK("k_bits");
  return 0;
}

extern "C" HX_API int hx_uniform(const uint32_t key[2], int impl, int dtype, double minval, double maxval, int64_t count,
                          void* out_dev, void* stream) {
  if (!key || !out_dev) return fail(HX_E_NULL, "hx_uniform: NULL argument");
  if (count < 0) return fail(HX_E_SHAPE, "hx_uniform: count < 0");
  if (int e = check_dtype_impl(dtype, impl, "hx_uniform")) return e;
  if (count == 0) return 0;
  Key k{key[0], key[1]};
  cudaStream_t s = (cudaStream_t)stream;
  const int g = grid_for(count, 256);
  if (dtype == HX_F32) {
    const float lo = (float)minval, hi = (float)maxval;
    const float span = hi - lo;   // rounded in f32 like jax.random.uniform
    if (impl == HX_RNG_LEGACY) k_uniform<float, true><<<g, 256, 0, s>>>(k, (uint64_t)count, lo, span, (float*)out_dev);
    else k_uniform<float, false><<<g, 256, 0, s>>>(k, (uint64_t)count, lo, span, (float*)out_dev);
  } else {
    const double span = maxval - minval;
    if (impl == HX_RNG_LEGACY) k_uniform<double, true><<<g, 256, 0, s>>>(k, (uint64_t)count, minval, span, (double*)out_dev);
    else k_uniform<double, false><<<g, 256, 0, s>>>(k, (uint64_t)count, minval, span, (double*)out_dev);
  }
  HX_LAUNCH_CHECK("k_uniform");
  return 0;
}

extern "C" HX_API int hx_normal(const uint32_t key[2], int impl, int dtype, int64_t count, void* out_dev, void* stream) {
  if (!key || !out_dev) return fail(HX_E_NULL, "hx_normal: NULL argument");
  const bool exact = (impl & HX_RNG_EXACT_NORMALS) != 0;
  impl &= ~HX_RNG_EXACT_NORMALS;
  if (count < 0) return fail(HX_E_SHAPE, "hx_normal: count < 0");
  if (int e = check_dtype_impl(dtype, impl, "hx_normal")) return e;
  if (count == 0) return 0;
  Key k{key[0], key[1]};
  cudaStream_t s = (cudaStream_t)stream;
  const int g = grid_for(count, 256);
  if (dtype == HX_F32) {
    if (impl == HX_RNG_LEGACY) k_normal<float, true><<<g, 256, 0, s>>>(k, (uint64_t)count, (float*)out_dev, exact);
    else k_normal<float, false><<<g, 256, 0, s>>>(k, (uint64_t)count, (float*)out_dev, exact);
  } else {
    if (impl == HX_RNG_LEGACY) k_normal<double, true><<<g, 256, 0, s>>>(k, (uint64_t)count, (double*)out_dev, false);
    else k_normal<double, false><<<g, 256, 0, s>>>(k, (uint64_t)count, (double*)out_dev, false);
  }
  HX_LAUNCH_CHECK("k_normal");
  return 0;
}

static int shuffle_rounds(int n) {
  // jax.random._shuffle: ceil(3 * ln(max(1, n)) / ln(2^32 - 1))
  const double r = std::ceil(3.0 * std::log((double)(n > 1 ? n : 1)) / std::log(4294967295.0));
  return (int)r;
}

static int launch_choice2(const uint32_t* keys_dev, Key parent, const uint32_t* parent_ptr, uint64_t nsplit,
                          int64_t first, int64_t nkeys, int impl, int n, int32_t* out_dev, cudaStream_t s) {
  const int rounds = shuffle_rounds(n);
  if (rounds > 2) return fail(HX_E_UNSUPPORTED, "choice: n=%d needs %d shuffle rounds (max 2 supported)", n, rounds);
  const int g = (int)((nkeys < 148 * 8) ? nkeys : 148 * 8);
  if (impl == HX_RNG_LEGACY) k_choice2<true><<<g, 256, 0, s>>>(keys_dev, parent, parent_ptr, nsplit, first, nkeys, n, rounds, out_dev);
  else k_choice2<false><<<g, 256, 0, s>>>(keys_dev, parent, parent_ptr, nsplit, first, nkeys, n, rounds, out_dev);
  HX_LAUNCH_CHECK("k_choice2");
  return 0;
}

extern "C" HX_API int hx_choice2(hx_ctx* ctx, const uint32_t* keys_dev, int64_t nkeys, int impl, int32_t n, int32_t* out_dev,
                          void* stream) {
  (void)ctx;
  if (!keys_dev || !out_dev) return fail(HX_E_NULL, "hx_choice2: NULL argument");
  if (nkeys < 0 || n < 2) return fail(HX_E_SHAPE, "hx_choice2: need nkeys >= 0 and n >= 2 (got %lld, %d)", (long long)nkeys, n);
  if (int e = check_dtype_impl(HX_F32, impl, "hx_choice2")) return e;
  if (nkeys == 0) return 0;
  return launch_choice2(keys_dev, Key{0, 0}, nullptr, 0, 0, nkeys, impl, n, out_dev, (cudaStream_t)stream);
}

// ---- prep: mean, C, M ---------------------------------------------------------------------------------------
template <typename T>
__global__ void k_pad_copy(const T* __restrict__ src, int n, int d, int ldd, T* __restrict__ dst) {
  const size_t tot = (size_t)n * ldd;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ldd), c = (int)(i - (size_t)r * ldd);
    dst[i] = (c < d) ? src[(size_t)r * d + c] : T(0);
  }
}

// C = (X2 - mean_0(X2)) / sqrt(n), zero padded to ldd   (moves/hamiltonian/hmc_walk.py:35)
template <typename T>
static int center_complement(hx_ctx* ctx, const T* X2, int n, int d, int ldd, T** C_out, cudaStream_t s) {
  void *pC, *pMean, *pPart;
  const int nchunks = (n + kColChunk - 1) / kColChunk;
  if (int e = ws_get(ctx, WS_C, (size_t)n * ldd * sizeof(T), &pC)) return e;
  if (int e = ws_get(ctx, WS_MEAN, (size_t)d * sizeof(T), &pMean)) return e;
  if (int e = ws_get(ctx, WS_PART, (size_t)nchunks * d * sizeof(T), &pPart)) return e;
  dim3 g1((d + 127) / 128, nchunks);
  k_colsum_partial<T><<<g1, 128, 0, s>>>(X2, n, d, (T*)pPart);
  HX_LAUNCH_CHECK("k_colsum_partial");
  k_mean_final<T><<<(d + 31) / 32, 256, 0, s>>>((const T*)pPart, nchunks, n, d, (T*)pMean);
  HX_LAUNCH_CHECK("k_mean_final");
  k_center<T><<<grid_for((uint64_t)n * ldd, 256), 256, 0, s>>>(X2, (const T*)pMean, n, d, ldd, (T*)pC);
  HX_LAUNCH_CHECK("k_center");
  *C_out = (T*)pC;
  return 0;
}

// mean_0(X2) only (tensor-core path builds its own transposed planes)
template <typename T>
static int complement_mean(hx_ctx* ctx, const T* X2, int n, int d, T** mean_out, cudaStream_t s) {
  void *pMean, *pPart;
  const int nchunks = (n + kColChunk - 1) / kColChunk;
  if (int e = ws_get(ctx, WS_MEAN, (size_t)d * sizeof(T), &pMean)) return e;
  if (int e = ws_get(ctx, WS_PART, (size_t)nchunks * d * sizeof(T), &pPart)) return e;
  dim3 g1((d + 127) / 128, nchunks);
  k_colsum_partial<T><<<g1, 128, 0, s>>>(X2, n, d, (T*)pPart);
  HX_LAUNCH_CHECK("k_colsum_partial");
  k_mean_final<T><<<(d + 31) / 32, 256, 0, s>>>((const T*)pPart, nchunks, n, d, (T*)pMean);
  HX_LAUNCH_CHECK("k_mean_final");
  *mean_out = (T*)pMean;
  return 0;
}

// M = C^T C with a fixed-order split over rows (deterministic)
template <typename T>
static int gram_matrix(hx_ctx* ctx, const T* C, int n, int ldd, T** M_out, cudaStream_t s) {
  void *pM, *pGram;
  if (int e = ws_get(ctx, WS_M, (size_t)ldd * ldd * sizeof(T), &pM)) return e;
  const int tiles = (ldd + 63) / 64;
  int slices = (2 * ctx->sm_count) / (tiles * tiles);
  const int max_slices = (n + 255) / 256;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  int rows_per_slice = (n + slices - 1) / slices;
  rows_per_slice = ((rows_per_slice + 15) / 16) * 16;
  slices = (n + rows_per_slice - 1) / rows_per_slice;
  dim3 gg(tiles, tiles, slices);
  ctx->last_gram_slices = slices;
  if (slices == 1) {
    k_gram<T><<<gg, 256, 0, s>>>(C, n, ldd, rows_per_slice, (T*)pM);
    HX_LAUNCH_CHECK("k_gram");
  } else {
    if (int e = ws_get(ctx, WS_GRAMP, (size_t)slices * ldd * ldd * sizeof(T), &pGram)) return e;
    k_gram<T><<<gg, 256, 0, s>>>(C, n, ldd, rows_per_slice, (T*)pGram);
    HX_LAUNCH_CHECK("k_gram");
    k_sum_slices<T><<<grid_for((uint64_t)ldd * ldd, 256), 256, 0, s>>>((const T*)pGram, slices, (size_t)ldd * ldd, (T*)pM);
    HX_LAUNCH_CHECK("k_sum_slices");
  }
  *M_out = (T*)pM;
  return 0;
}

template <typename T>
struct Fuse {       // batch-level fused accept + chain emit
  const uint32_t* akey_ptr;
  T* lp1w;
  int32_t* accepts;
  T* chain_out;
  T* chain_lp;
  int32_t* nonfinite;
  const DistPush* push;   // sharded runs, tensor-core path: the finish kernel pushes the rows itself
};

static int check_move_args(const char* who, const void* X1, const void* X2, int n, int row0, int rows, int L,
                           const void* lp1, const void* Q, const void* logacc, const void* lp_out, double eps) {
  if (!X1 || !X2 || !lp1 || !Q || !logacc || !lp_out) return fail(HX_E_NULL, "%s: NULL state pointer", who);
  if (n < 2) return fail(HX_E_SHAPE, "%s: group size n=%d must be >= 2", who, n);
  if (row0 < 0 || rows < 1 || row0 + rows > n) return fail(HX_E_SHAPE, "%s: bad row shard [%d, %d) of %d", who, row0, row0 + rows, n);
  if (L < 1) return fail(HX_E_SHAPE, "%s: L=%d must be >= 1", who, L);
  if (!(eps == eps)) return fail(HX_E_SHAPE, "%s: step_size is NaN", who);
  return 0;
}

// ---- walk move ---------------------------------------------------------------------------------------------------
template <typename T>
static int walk_move_t(hx_ctx* ctx, int impl, const hx_target* tgt, const T* X1, const T* X2, int n, int row0,
                       int rows, double eps, Key key, const uint32_t* key_ptr, int L, const T* lp1, T* Q, T* logacc,
                       T* pproj, T* lp_out, int flags, const Fuse<T>* fuse, const tc::NextDraw* next, cudaStream_t s) {
  struct { WalkParams<T> P; } W;
  memset(&W.P, 0, sizeof(W.P));
  if (int e = make_target<T>(tgt, &W.P.tg, "hx_walk_move")) return e;
  const int d = tgt->dim, ldd = round4(d);
  if constexpr (std::is_same<T, float>::value) {
    if (!ctx->ctl_eps && tc::walk_eligible(ctx, HX_F32, tgt, n, rows, flags)) {
      float* mean;
      if (int e = complement_mean<float>(ctx, X2, n, d, &mean, s)) return e;
      if (!pproj) {
        void* pS;
        if (int e = ws_get(ctx, WS_S, (size_t)rows * d * sizeof(float), &pS)) return e;
        pproj = (float*)pS;
      }
      tc::FuseF32 f32;
      if (fuse) {
        f32.akey_ptr = fuse->akey_ptr; f32.lp1w = fuse->lp1w; f32.accepts = fuse->accepts;
        f32.chain_out = fuse->chain_out; f32.chain_lp = fuse->chain_lp; f32.nonfinite = fuse->nonfinite;
        f32.push = fuse->push;
      }
      ctx->launches += 2;
      ctx->last_path = 1;
      return tc::walk_half(ctx, impl, tgt, mean, X1, X2, n, row0, rows, eps, key, key_ptr, L, lp1, Q, logacc, pproj,
                           lp_out, fuse ? &f32 : nullptr, next, s);
    }
  }
  ctx->last_path = 0;
  T *C, *M;
  if ((size_t)n * ldd <= (size_t)kPrepSmall && ldd <= 64) {
    void *pC, *pMean, *pM;
    if (int e = ws_get(ctx, WS_C, (size_t)n * ldd * sizeof(T), &pC)) return e;
    if (int e = ws_get(ctx, WS_MEAN, (size_t)d * sizeof(T), &pMean)) return e;
    if (int e = ws_get(ctx, WS_M, (size_t)ldd * ldd * sizeof(T), &pM)) return e;
    k_prep_small<T><<<1, 256, 0, s>>>(X2, n, d, ldd, (T*)pMean, (T*)pC, (T*)pM);
    HX_LAUNCH_CHECK("k_prep_small");
    C = (T*)pC; M = (T*)pM;
    ctx->last_gram_slices = 1;
  } else {
    if (int e = center_complement<T>(ctx, X2, n, d, ldd, &C, s)) return e;
    if (int e = gram_matrix<T>(ctx, C, n, ldd, &M, s)) return e;
  }
  void* pG;
  if (int e = ws_get(ctx, WS_G, (size_t)rows * d * sizeof(T), &pG)) return e;
  if (!pproj) {
    void* pS;
    if (int e = ws_get(ctx, WS_S, (size_t)rows * d * sizeof(T), &pS)) return e;
    pproj = (T*)pS;
  }
  W.P.X1 = X1; W.P.C = C; W.P.M = M; W.P.lp1 = lp1; W.P.Q = Q; W.P.S = pproj; W.P.G = (T*)pG;
  W.P.logacc = logacc; W.P.lp_out = lp_out; W.P.dK_out = nullptr;
  W.P.n = n; W.P.d = d; W.P.ldd = ldd; W.P.row0 = row0; W.P.rows = rows; W.P.L = L;
  W.P.flags = (flags & HX_MOVE_FROZEN_GRAD) ? kFlagFrozen : 0;
  // narrow targets with many complement walkers: the RNG-bound projection runs in its own kernel (lanes split the complement
  // index, no barrier per draw) and the half-move kernel starts from the given S0.  The choice depends on (n, dim) only, so
  // row shards reproduce the single-GPU chain.
  if (ldd <= 40 && n >= 1024 && !ctx->narrow_off) {       // beyond ~40 columns the per-lane partial sums crowd out occupancy
    if (int e = launch_project_narrow<T>(key, key_ptr, rng_flags(ctx, impl), C, n, ldd, d, row0, rows, pproj, ctx->sm_count, s))
      return e;
    W.P.flags |= kFlagGivenMomentum;
    ctx->launches += 1;
  }
  W.P.eps = (T)eps; W.P.key = key; W.P.key_ptr = key_ptr;
  W.P.eps_ptr = (const T*)ctx->ctl_eps; W.P.L_ptr = ctx->ctl_L;
  if (fuse) {
    W.P.fuse_accept = 1; W.P.akey_ptr = fuse->akey_ptr; W.P.X1w = const_cast<T*>(X1); W.P.lp1w = fuse->lp1w;
    W.P.accepts = fuse->accepts; W.P.chain_out = fuse->chain_out; W.P.chain_lp = fuse->chain_lp;
    W.P.nonfinite = fuse->nonfinite;
  }
  W.P.legacy = rng_flags(ctx, impl);
  ctx->launches += 5 + (ctx->last_gram_slices > 1 ? 1 : 0);   // colsum, mean, centre, gram(+reduce), half-move
  if (int e = timing_begin(ctx, s)) return e;
  if (int e = launch_walk<T>(W.P, ctx->sm_count, s)) return e;
  return timing_end(ctx, s);
}

extern "C" HX_API int hx_walk_move(hx_ctx* ctx, int dtype, int impl, const hx_target* tgt, const void* X1, const void* X2,
                            int32_t n, int32_t row0, int32_t rows, double step_size, const uint32_t key[2], int32_t L,
                            const void* lp1, void* Q, void* logacc, void* pproj, void* lp_out, int flags, void* stream) {
  if (!ctx || !key) return fail(HX_E_NULL, "hx_walk_move: NULL ctx/key");
  if (int e = check_dtype_impl(dtype, impl, "hx_walk_move")) return e;
  if (int e = check_move_args("hx_walk_move", X1, X2, n, row0, rows, L, lp1, Q, logacc, lp_out, step_size)) return e;
  HX_CUDA(cudaSetDevice(ctx->device));
  const Key k{key[0], key[1]};
  cudaStream_t s = (cudaStream_t)stream;
  tc::NextDraw nd;
  const tc::NextDraw* next = nullptr;
  if (ctx->hint_valid) {           // hx_walk_prefetch: start the following half-move's momentum draw under this one
    ctx->hint_valid = 0;
    if (ctx->hint.impl == impl && ctx->hint.n == n) {
      nd.key = Key{ctx->hint.k0, ctx->hint.k1}; nd.key_ptr = nullptr; nd.row0 = ctx->hint.row0; nd.rows = ctx->hint.rows;
      next = &nd;
    }
  }
  if (dtype == HX_F32)
    return walk_move_t<float>(ctx, impl, tgt, (const float*)X1, (const float*)X2, n, row0, rows, step_size, k, nullptr, L,
                              (const float*)lp1, (float*)Q, (float*)logacc, (float*)pproj, (float*)lp_out, flags, nullptr, next, s);
  return walk_move_t<double>(ctx, impl, tgt, (const double*)X1, (const double*)X2, n, row0, rows, step_size, k, nullptr, L,
                             (const double*)lp1, (double*)Q, (double*)logacc, (double*)pproj, (double*)lp_out, flags, nullptr, nullptr, s);
}

// ---- side move ---------------------------------------------------------------------------------------------------
template <typename T>
static int side_move_t(hx_ctx* ctx, int impl, const hx_target* tgt, const T* X1, const T* X2, int n, int row0,
                       int rows, double eps, Key key, const uint32_t* key_ptr, int L, const T* lp1, T* Q, T* logacc,
                       T* pproj, T* lp_out, int flags, const Fuse<T>* fuse, cudaStream_t s) {
  struct { SideParams<T> P; } W;
  memset(&W.P, 0, sizeof(W.P));
  if (int e = make_target<T>(tgt, &W.P.tg, "hx_side_move")) return e;
  const int d = tgt->dim, ldd = round4(d);
  void *pG, *pIdx;
  if (int e = ws_get(ctx, WS_G, (size_t)rows * d * sizeof(T), &pG)) return e;
  if (int e = ws_get(ctx, WS_IDX, (size_t)rows * 2 * sizeof(int32_t), &pIdx)) return e;
  if (!pproj) {
    void* pS;
    if (int e = ws_get(ctx, WS_S, (size_t)rows * d * sizeof(T), &pS)) return e;
    pproj = (T*)pS;
  }
  // keys[0:n] of split(key, n + 2) drive the per-walker choices (hmc_side.py:37-47)
  if (int e = launch_choice2(nullptr, key, key_ptr, (uint64_t)n + 2, row0, rows, impl, n, (int32_t*)pIdx, s)) return e;
  W.P.X1 = X1; W.P.X2 = X2; W.P.idx2 = (const int32_t*)pIdx; W.P.lp1 = lp1; W.P.Q = Q; W.P.D = pproj; W.P.G = (T*)pG;
  W.P.p_io = nullptr; W.P.logacc = logacc; W.P.lp_out = lp_out;
  W.P.n = n; W.P.d = d; W.P.ldd = ldd; W.P.row0 = row0; W.P.rows = rows; W.P.L = L;
  W.P.flags = (flags & HX_MOVE_FROZEN_GRAD) ? kFlagFrozen : 0;
  W.P.eps = (T)eps; W.P.key = key; W.P.key_ptr = key_ptr;
  W.P.eps_ptr = (const T*)ctx->ctl_eps; W.P.L_ptr = ctx->ctl_L;
  if (fuse) {
    W.P.fuse_accept = 1; W.P.akey_ptr = fuse->akey_ptr; W.P.X1w = const_cast<T*>(X1); W.P.lp1w = fuse->lp1w;
    W.P.accepts = fuse->accepts; W.P.chain_out = fuse->chain_out; W.P.chain_lp = fuse->chain_lp;
    W.P.nonfinite = fuse->nonfinite;
  }
  W.P.legacy = rng_flags(ctx, impl);
  ctx->launches += 2;   // choice, half-move
  if (int e = timing_begin(ctx, s)) return e;
  if (int e = launch_side<T>(W.P, ctx->sm_count, s)) return e;
  return timing_end(ctx, s);
}

extern "C" HX_API int hx_side_move(hx_ctx* ctx, int dtype, int impl, const hx_target* tgt, const void* X1, const void* X2,
                            int32_t n, int32_t row0, int32_t rows, double step_size, const uint32_t key[2], int32_t L,
                            const void* lp1, void* Q, void* logacc, void* pproj, void* lp_out, int flags, void* stream) {
  if (!ctx || !key) return fail(HX_E_NULL, "hx_side_move: NULL ctx/key");
  if (int e = check_dtype_impl(dtype, impl, "hx_side_move")) return e;
  if (int e = check_move_args("hx_side_move", X1, X2, n, row0, rows, L, lp1, Q, logacc, lp_out, step_size)) return e;
  if (impl == HX_RNG_LEGACY && (uint64_t)(n + 2) * 2 > 0xFFFFFFFFull) return fail(HX_E_SHAPE, "hx_side_move: n too large for legacy split");
  HX_CUDA(cudaSetDevice(ctx->device));
  const Key k{key[0], key[1]};
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == HX_F32)
    return side_move_t<float>(ctx, impl, tgt, (const float*)X1, (const float*)X2, n, row0, rows, step_size, k, nullptr, L,
                              (const float*)lp1, (float*)Q, (float*)logacc, (float*)pproj, (float*)lp_out, flags, nullptr, s);
  return side_move_t<double>(ctx, impl, tgt, (const double*)X1, (const double*)X2, n, row0, rows, step_size, k, nullptr, L,
                             (const double*)lp1, (double*)Q, (double*)logacc, (double*)pproj, (double*)lp_out, flags, nullptr, s);
}

// ---- integrator-only entries --------------------------------------------------------------------------------------
template <typename T>
static int leapfrog_walk_t(hx_ctx* ctx, const hx_target* tgt, T* q, T* S, const T* centered, int n2, int rows,
                           double eps, int L, T* dK_out, cudaStream_t s) {
  struct { WalkParams<T> P; } W;
  memset(&W.P, 0, sizeof(W.P));
  if (int e = make_target<T>(tgt, &W.P.tg, "hx_leapfrog_walk")) return e;
  const int d = tgt->dim, ldd = round4(d);
  void *pC, *pG, *pLa, *pLp;
  if (int e = ws_get(ctx, WS_C, (size_t)n2 * ldd * sizeof(T), &pC)) return e;
  k_pad_copy<T><<<grid_for((uint64_t)n2 * ldd, 256), 256, 0, s>>>(centered, n2, d, ldd, (T*)pC);
  HX_LAUNCH_CHECK("k_pad_copy");
  T* M;
  if (int e = gram_matrix<T>(ctx, (const T*)pC, n2, ldd, &M, s)) return e;
  if (int e = ws_get(ctx, WS_G, (size_t)rows * d * sizeof(T), &pG)) return e;
  if (int e = ws_get(ctx, WS_LA, (size_t)rows * sizeof(T), &pLa)) return e;
  if (int e = ws_get(ctx, WS_LP, (size_t)2 * rows * sizeof(T), &pLp)) return e;
  HX_CUDA(cudaMemsetAsync(pLp, 0, (size_t)2 * rows * sizeof(T), s));
  W.P.X1 = q; W.P.C = (const T*)pC; W.P.M = M; W.P.lp1 = (const T*)pLp; W.P.Q = q; W.P.S = S; W.P.G = (T*)pG;
  W.P.logacc = (T*)pLa; W.P.lp_out = (T*)pLp + rows; W.P.dK_out = dK_out;
  W.P.n = n2; W.P.d = d; W.P.ldd = ldd; W.P.row0 = 0; W.P.rows = rows; W.P.L = L;
  W.P.flags = kFlagGivenMomentum | kFlagInPlaceQ;
  W.P.eps = (T)eps;
  return launch_walk<T>(W.P, ctx->sm_count, s);
}

extern "C" HX_API int hx_leapfrog_walk(hx_ctx* ctx, int dtype, const hx_target* tgt, void* q, void* S, const void* centered,
                                int32_t n2, int32_t rows, double step_size, int32_t L, void* dK_out, void* stream) {
  if (!ctx || !q || !S || !centered) return fail(HX_E_NULL, "hx_leapfrog_walk: NULL argument");
  if (int e = check_dtype_impl(dtype, HX_RNG_PARTITIONABLE, "hx_leapfrog_walk")) return e;
  if (n2 < 1 || rows < 1 || L < 1) return fail(HX_E_SHAPE, "hx_leapfrog_walk: need n2, rows, L >= 1");
  HX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == HX_F32) return leapfrog_walk_t<float>(ctx, tgt, (float*)q, (float*)S, (const float*)centered, n2, rows, step_size, L, (float*)dK_out, s);
  return leapfrog_walk_t<double>(ctx, tgt, (double*)q, (double*)S, (const double*)centered, n2, rows, step_size, L, (double*)dK_out, s);
}

template <typename T>
static int leapfrog_side_t(hx_ctx* ctx, const hx_target* tgt, T* q, T* p, const T* diff, int rows, double eps, int L,
                           cudaStream_t s) {
  struct { SideParams<T> P; } W;
  memset(&W.P, 0, sizeof(W.P));
  if (int e = make_target<T>(tgt, &W.P.tg, "hx_leapfrog_side")) return e;
  const int d = tgt->dim, ldd = round4(d);
  void *pG, *pD, *pLa, *pLp;
  if (int e = ws_get(ctx, WS_G, (size_t)rows * d * sizeof(T), &pG)) return e;
  if (int e = ws_get(ctx, WS_S, (size_t)rows * d * sizeof(T), &pD)) return e;
  if (int e = ws_get(ctx, WS_LA, (size_t)rows * sizeof(T), &pLa)) return e;
  if (int e = ws_get(ctx, WS_LP, (size_t)2 * rows * sizeof(T), &pLp)) return e;
  HX_CUDA(cudaMemcpyAsync(pD, diff, (size_t)rows * d * sizeof(T), cudaMemcpyDeviceToDevice, s));
  HX_CUDA(cudaMemsetAsync(pLp, 0, (size_t)2 * rows * sizeof(T), s));
  W.P.X1 = q; W.P.X2 = nullptr; W.P.idx2 = nullptr; W.P.lp1 = (const T*)pLp; W.P.Q = q; W.P.D = (T*)pD; W.P.G = (T*)pG;
  W.P.p_io = p; W.P.logacc = (T*)pLa; W.P.lp_out = (T*)pLp + rows;
  W.P.n = rows; W.P.d = d; W.P.ldd = ldd; W.P.row0 = 0; W.P.rows = rows; W.P.L = L;
  W.P.flags = kFlagGivenMomentum | kFlagInPlaceQ;
  W.P.eps = (T)eps;
  return launch_side<T>(W.P, ctx->sm_count, s);
}

extern "C" HX_API int hx_leapfrog_side(hx_ctx* ctx, int dtype, const hx_target* tgt, void* q, void* p, const void* diff,
                                int32_t rows, double step_size, int32_t L, void* stream) {
  if (!ctx || !q || !p || !diff) return fail(HX_E_NULL, "hx_leapfrog_side: NULL argument");
  if (int e = check_dtype_impl(dtype, HX_RNG_PARTITIONABLE, "hx_leapfrog_side")) return e;
  if (rows < 1 || L < 1) return fail(HX_E_SHAPE, "hx_leapfrog_side: need rows, L >= 1");
  HX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == HX_F32) return leapfrog_side_t<float>(ctx, tgt, (float*)q, (float*)p, (const float*)diff, rows, step_size, L, s);
  return leapfrog_side_t<double>(ctx, tgt, (double*)q, (double*)p, (const double*)diff, rows, step_size, L, s);
}

// ---- log-density ----------------------------------------------------------------------------------------------------
template <typename T>
static int log_prob_t(hx_ctx* ctx, const hx_target* tgt, const T* X, int64_t rows, T* lp, T* grad, cudaStream_t s) {
  struct { LogProbParams<T> P; } W;
  memset(&W.P, 0, sizeof(W.P));
  if (int e = make_target<T>(tgt, &W.P.tg, "hx_log_prob")) return e;
  const int d = tgt->dim, ldd = round4(d);
  T* G = grad;
  if (!G) {
    void* pG;
    if (int e = ws_get(ctx, WS_G, (size_t)rows * d * sizeof(T), &pG)) return e;
    G = (T*)pG;
  }
  W.P.X = X; W.P.G = G; W.P.lp = lp; W.P.rows = (int)rows; W.P.ldd = ldd; W.P.negate_grad = grad ? 1 : 0;
  return launch_log_prob<T>(W.P, ctx->sm_count, s);
}

extern "C" HX_API int hx_log_prob(hx_ctx* ctx, int dtype, const hx_target* tgt, const void* X, int64_t rows, void* lp,
                           void* grad, void* stream) {
  if (!ctx || !X || !lp) return fail(HX_E_NULL, "hx_log_prob: NULL argument");
  if (int e = check_dtype_impl(dtype, HX_RNG_PARTITIONABLE, "hx_log_prob")) return e;
  if (rows < 0 || rows > 0x7FFFFFFF) return fail(HX_E_SHAPE, "hx_log_prob: bad rows");
  if (rows == 0) return 0;
  HX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == HX_F32) return log_prob_t<float>(ctx, tgt, (const float*)X, rows, (float*)lp, (float*)grad, s);
  return log_prob_t<double>(ctx, tgt, (const double*)X, rows, (double*)lp, (double*)grad, s);
}

// ---- accept ---------------------------------------------------------------------------------------------------------
extern "C" HX_API int hx_accept(hx_ctx* ctx, int dtype, int impl, void* cur, const void* prop, void* lp_cur,
                         const void* lp_prop, const void* logacc, const uint32_t key[2], int32_t n, int32_t row0,
                         int32_t rows, int32_t dim, int32_t* accepts, void* stream) {
  if (!ctx || !cur || !prop || !lp_cur || !lp_prop || !logacc || !key || !accepts) return fail(HX_E_NULL, "hx_accept: NULL argument");
  if (int e = check_dtype_impl(dtype, impl, "hx_accept")) return e;
  if (n < 1 || row0 < 0 || rows < 1 || row0 + rows > n || dim < 1) return fail(HX_E_SHAPE, "hx_accept: bad shape");
  HX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  const Key k{key[0], key[1]};
  const int grid = (rows + 7) / 8;
  const bool leg = impl == HX_RNG_LEGACY;
  if (dtype == HX_F32) {
    if (leg) k_accept<float, true><<<grid, 256, 0, s>>>((float*)cur, (const float*)prop, (float*)lp_cur, (const float*)lp_prop, (const float*)logacc, k, n, row0, rows, dim, accepts);
    else k_accept<float, false><<<grid, 256, 0, s>>>((float*)cur, (const float*)prop, (float*)lp_cur, (const float*)lp_prop, (const float*)logacc, k, n, row0, rows, dim, accepts);
  } else {
    if (leg) k_accept<double, true><<<grid, 256, 0, s>>>((double*)cur, (const double*)prop, (double*)lp_cur, (const double*)lp_prop, (const double*)logacc, k, n, row0, rows, dim, accepts);
    else k_accept<double, false><<<grid, 256, 0, s>>>((double*)cur, (const double*)prop, (double*)lp_cur, (const double*)lp_prop, (const double*)logacc, k, n, row0, rows, dim, accepts);
  }
  HX_LAUNCH_CHECK("k_accept");
  return 0;
}

// ---- CUDA graphs for launch-bound batch loops -----------------------------------------------------------------------------------
// Small ensembles spend their time in launch latency (a 32-walker iteration is ~4 launches of a few microseconds each).  A batch
// call whose argument set was seen before is captured once into a CUDA graph (on a private stream; the caller's stream may be
// the legacy default stream, which cannot be captured) and replayed afterwards; the per-batch keys are first copied to a fixed
// staging address so the replays are argument-identical.  First use of an argument set runs eagerly and primes the workspaces
// (no allocation may happen during capture).  Not used while the timing hooks are on, nor on the tensor-core path (side streams).
static uint64_t fnv1a(const void* p, size_t nbytes, uint64_t h = 1469598103934665603ull) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < nbytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

static int stage_keys(hx_ctx* ctx, const uint32_t* keys, size_t nwords, const uint32_t** staged, cudaStream_t s) {
  const size_t bytes = nwords * sizeof(uint32_t);
  if (ctx->key_stage_cap < bytes) {
    if (ctx->key_stage) HX_CUDA(cudaFree(ctx->key_stage));
    ctx->key_stage = nullptr; ctx->key_stage_cap = 0;
    HX_CUDA(cudaMalloc(&ctx->key_stage, bytes * 2));
    ctx->key_stage_cap = bytes * 2;
    if (ctx->graphs) {                       // cached graphs read the old staging address
      for (GraphEntry& g : *ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
      ctx->graphs->clear();
    }
  }
  HX_CUDA(cudaMemcpyAsync(ctx->key_stage, keys, bytes, cudaMemcpyDeviceToDevice, s));
  *staged = (const uint32_t*)ctx->key_stage;
  return 0;
}

template <class F>
static int run_maybe_graphed(hx_ctx* ctx, cudaStream_t s, uint64_t key, F&& enqueue) {
  if (ctx->timing || ctx->graphs_off) return enqueue(s);
  if (!ctx->graphs) ctx->graphs = new (std::nothrow) std::vector<GraphEntry>();
  if (!ctx->graphs) return enqueue(s);
  ctx->graph_tick += 1;
  GraphEntry* ent = nullptr;
  for (GraphEntry& g : *ctx->graphs) if (g.key == key) ent = &g;
  if (!ent) {
    if (ctx->graphs->size() >= 16) {         // evict the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < ctx->graphs->size(); ++i) if ((*ctx->graphs)[i].last_use < (*ctx->graphs)[lru].last_use) lru = i;
      if ((*ctx->graphs)[lru].exec) cudaGraphExecDestroy((*ctx->graphs)[lru].exec);
      ctx->graphs->erase(ctx->graphs->begin() + lru);
    }
    ctx->graphs->push_back(GraphEntry{key, 0, nullptr, ctx->graph_tick});
    return enqueue(s);                       // first use: eager (primes workspaces and function attributes)
  }
  ent->last_use = ctx->graph_tick;
  if (ent->exec && ent->ws_gen != ctx->ws_generation) {
    // a workspace slot was reallocated since the capture (a larger problem ran on this ctx): the graph's kernel arguments
    // point into freed memory.  Drop it, run eagerly now (re-primes the slots for this size), capture again on the next use.
    cudaGraphExecDestroy(ent->exec);
    ent->exec = nullptr;
    return enqueue(s);
  }
  if (!ent->exec) {
    if (!ctx->gstream) HX_CUDA(cudaStreamCreateWithFlags(&ctx->gstream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(ctx->gstream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); ctx->graphs_off = 1; return enqueue(s); }
    const int e = enqueue(ctx->gstream);
    const cudaError_t ce = cudaStreamEndCapture(ctx->gstream, &graph);
    if (e || ce != cudaSuccess || !graph) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      ctx->graphs_off = 1;                   // fall back to eager launches for good
      return e ? e : enqueue(s);
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ci = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ci != cudaSuccess || !exec) { cudaGetLastError(); ctx->graphs_off = 1; return enqueue(s); }
    ent->exec = exec;
    ent->ws_gen = ctx->ws_generation;
  }
  HX_CUDA(cudaGraphLaunch(ent->exec, s));
  return 0;
}

// thinned chain emit: iteration t of a batch call is stored iff (t + phase) % every == 0, into consecutive slots
static inline int64_t chain_slot(const hx_ctx* ctx, int64_t t) {
  const int every = ctx->chain_every > 1 ? ctx->chain_every : 1;
  if (every == 1) return t;
  const int64_t g = t + ctx->chain_phase;
  return (g % every == 0) ? (g / every - (ctx->chain_phase + every - 1) / every) : -1;
}

extern "C" HX_API int hx_ctx_set_chain_thinning(hx_ctx* ctx, int32_t every, int32_t phase) {
  if (!ctx) return fail(HX_E_NULL, "hx_ctx_set_chain_thinning: NULL ctx");
  if (every < 1 || phase < 0 || phase >= every) return fail(HX_E_SHAPE, "hx_ctx_set_chain_thinning: need every >= 1 and 0 <= phase < every");
  ctx->chain_every = every;
  ctx->chain_phase = phase;
  return 0;
}

// ---- batch level: T iterations of the main phase ----------------------------------------------------------------------
template <typename T>
static int run_iterations_body(hx_ctx* ctx, int impl, int move_kind, int key_order, const hx_target* tgt, T* X, T* lp, int n, double eps,
                            int L, const uint32_t* keys, int64_t Tn, T* chain, T* chain_lp, int32_t* accepts,
                            int32_t* nonfinite, cudaStream_t s) {
  const int d = tgt->dim;
  void *pQ, *pLa, *pLp;
  if (int e = ws_get(ctx, WS_Q, (size_t)n * d * sizeof(T), &pQ)) return e;
  if (int e = ws_get(ctx, WS_LA, (size_t)n * sizeof(T), &pLa)) return e;
  if (int e = ws_get(ctx, WS_LP, (size_t)n * sizeof(T), &pLp)) return e;
  const size_t W2 = (size_t)2 * n;
  for (int64_t t = 0; t < Tn; ++t) {
    const uint32_t* kt = keys + (size_t)t * 8;
    for (int half = 0; half < 2; ++half) {
      // main-phase key order (samplers/hamiltonian_ensemble.py:290,295,300,305):
      //   move1 = keys[0], accept1 = keys[2], move2 = keys[1], accept2 = keys[3]
      // warm-up order (:171,184,191,204): move1 = keys[0], accept1 = keys[1], move2 = keys[2], accept2 = keys[3]
      const int im = key_order == 0 ? (half == 0 ? 0 : 1) : (half == 0 ? 0 : 2);
      const int ia = key_order == 0 ? (half == 0 ? 2 : 3) : (half == 0 ? 1 : 3);
      const uint32_t* kmove = kt + 2 * im;
      const uint32_t* kacc = kt + 2 * ia;
      T* Xa = X + (size_t)half * n * d;          // group being updated
      T* Xb = X + (size_t)(1 - half) * n * d;    // complement (group 2 sees the updated group 1)
      T* lpa = lp + (size_t)half * n;
      Fuse<T> f;
      f.akey_ptr = kacc; f.lp1w = lpa; f.accepts = accepts + (size_t)half * n;
      const int64_t slot = chain_slot(ctx, t);     // -1: this iteration is not stored (thinned emit)
      f.chain_out = (chain && slot >= 0) ? chain + ((size_t)slot * W2 + (size_t)half * n) * d : nullptr;
      f.chain_lp = (chain_lp && slot >= 0) ? chain_lp + (size_t)slot * W2 + (size_t)half * n : nullptr;
      f.nonfinite = nonfinite;
      f.push = nullptr;
      int e;
      // the half-move after this one (its momentum draw only needs the key): move2 of this iteration or move1 of the next
      tc::NextDraw nd;
      const tc::NextDraw* next = nullptr;
      if (half == 0 || t + 1 < Tn) {
        const uint32_t* kn = half == 0 ? kt + 2 * (key_order == 0 ? 1 : 2) : kt + 8;
        nd.key = Key{0, 0}; nd.key_ptr = kn; nd.row0 = 0; nd.rows = n;
        next = &nd;
      }
      if (move_kind == 0)
        e = walk_move_t<T>(ctx, impl, tgt, Xa, Xb, n, 0, n, eps, Key{0, 0}, kmove, L, lpa, (T*)pQ, (T*)pLa, nullptr, (T*)pLp, 0, &f, next, s);
      else
        e = side_move_t<T>(ctx, impl, tgt, Xa, Xb, n, 0, n, eps, Key{0, 0}, kmove, L, lpa, (T*)pQ, (T*)pLa, nullptr, (T*)pLp, 0, &f, s);
      if (e) return e;
    }
    if (ctx->stats_on)
      if (int e = stats_feed(ctx, sizeof(T) == 4 ? HX_F32 : HX_F64, &ctx->stats, X, s)) return e;
  }
  return 0;
}

template <typename T>
static int run_iterations_t(hx_ctx* ctx, int impl, int move_kind, int key_order, const hx_target* tgt, T* X, T* lp, int n, double eps,
                            int L, const uint32_t* keys, int64_t Tn, T* chain, T* chain_lp, int32_t* accepts,
                            int32_t* nonfinite, cudaStream_t s) {
  bool tc_path = false;
  if constexpr (std::is_same<T, float>::value) tc_path = move_kind == 0 && tc::walk_eligible(ctx, HX_F32, tgt, n, n, 0);
  // launch-bound sizes only: a graph node per launch is not worth it once the kernels run for hundreds of microseconds
  const bool small = (size_t)n * tgt->dim <= 65536 && Tn >= 1 && Tn <= 4096;
  if (tc_path || !small || ctx->timing || ctx->graphs_off)
    return run_iterations_body<T>(ctx, impl, move_kind, key_order, tgt, X, lp, n, eps, L, keys, Tn, chain, chain_lp, accepts, nonfinite, s);
  const uint32_t* staged = nullptr;
  if (int e = stage_keys(ctx, keys, (size_t)Tn * 8, &staged, s)) return e;
  struct { int impl, move_kind, key_order, n, L, every, phase, dt, narrow_off, leap_form, stats_on; double eps; int64_t Tn; const void *X, *lp, *chain, *clp, *acc, *nf; hx_target tg; hx_chain_stats st; } K;
  memset(&K, 0, sizeof(K));
  K.stats_on = ctx->stats_on; if (ctx->stats_on) K.st = ctx->stats;
  K.impl = impl; K.move_kind = move_kind; K.key_order = key_order; K.n = n; K.L = L; K.every = ctx->chain_every; K.phase = ctx->chain_phase;
  K.dt = (int)sizeof(T); K.narrow_off = ctx->narrow_off; K.leap_form = ctx->leap_form; K.eps = eps; K.Tn = Tn; K.X = X; K.lp = lp; K.chain = chain; K.clp = chain_lp; K.acc = accepts; K.nf = nonfinite;
  K.tg = *tgt;
  const uint64_t key = fnv1a(&K, sizeof(K), 0x9e3779b97f4a7c15ull ^ (uint64_t)ctx->precision);
  return run_maybe_graphed(ctx, s, key, [&](cudaStream_t cs) {
    return run_iterations_body<T>(ctx, impl, move_kind, key_order, tgt, X, lp, n, eps, L, staged, Tn, chain, chain_lp, accepts, nonfinite, cs);
  });
}

extern "C" HX_API int hx_run_iterations(hx_ctx* ctx, int dtype, int impl, int move_kind, int key_order, const hx_target* tgt, void* X,
                                 void* lp, int32_t n, double step_size, int32_t L, const uint32_t* keys_dev, int64_t T,
                                 void* chain_out, void* lp_out, int32_t* accepts, int32_t* nonfinite, void* stream) {
  if (!ctx || !tgt || !X || !lp || !keys_dev || !accepts || !nonfinite) return fail(HX_E_NULL, "hx_run_iterations: NULL argument");
  if (int e = check_dtype_impl(dtype, impl, "hx_run_iterations")) return e;
  if (move_kind != 0 && move_kind != 1) return fail(HX_E_DTYPE, "hx_run_iterations: move_kind must be 0 (walk) or 1 (side)");
  if (key_order != 0 && key_order != 1) return fail(HX_E_DTYPE, "hx_run_iterations: key_order must be 0 (main) or 1 (warm-up)");
  if (n < 2 || L < 1 || T < 0) return fail(HX_E_SHAPE, "hx_run_iterations: need n >= 2, L >= 1, T >= 0");
  HX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == HX_F32)
    return run_iterations_t<float>(ctx, impl, move_kind, key_order, tgt, (float*)X, (float*)lp, n, step_size, L, keys_dev, T,
                                   (float*)chain_out, (float*)lp_out, accepts, nonfinite, s);
  return run_iterations_t<double>(ctx, impl, move_kind, key_order, tgt, (double*)X, (double*)lp, n, step_size, L, keys_dev, T,
                                  (double*)chain_out, (double*)lp_out, accepts, nonfinite, s);
}

// ---- (6) derivative-free ensemble moves (hx_vanilla.cuh) -----------------------------------------------------------------
template <typename T>
static int vanilla_half_t(hx_ctx* ctx, int impl, int kind, const hx_target* tgt, T* X1, const T* X2, int n, int row0, int rows,
                          double stretch, Key key, const uint32_t* key_ptr, T* lp1, T* Q, T* logacc, T* lp_out, int do_accept,
                          int32_t* accepts, T* chain_out, T* chain_lp, int32_t* nonfinite, cudaStream_t s) {
  const int d = tgt->dim;
  const int legacy = rng_flags(ctx, impl);
  void *pIdx = nullptr, *pExtra, *pShift = nullptr;
  if (int e = ws_get(ctx, WS_LA, (size_t)2 * rows * sizeof(T), &pExtra)) return e;      // [extra | (unused)]
  if (kind == kVanillaStretch || kind == kVanillaSide) {
    if (int e = ws_get(ctx, WS_IDX, (size_t)rows * 2 * sizeof(int32_t), &pIdx)) return e;
    // keys[0:n] of split(key, n + 2) drive the per-walker choices (stretch.py:35-46, side.py:30-39)
    if (int e = launch_choice2(nullptr, key, key_ptr, (uint64_t)n + 2, row0, rows, impl, n, (int32_t*)pIdx, s)) return e;
  } else {
    T* mean;
    if (int e = complement_mean<T>(ctx, X2, n, d, &mean, s)) return e;
    if (int e = ws_get(ctx, WS_AD_MEAN, (size_t)2 * d * sizeof(T), &pShift)) return e;
    k_vanilla_walk_shift<T><<<(d + 255) / 256, 256, 0, s>>>(X2, mean, n, d, key, key_ptr, legacy, (T*)pShift);
    HX_LAUNCH_CHECK("k_vanilla_walk_shift");
  }
  k_vanilla_propose<T><<<(rows + 7) / 8, 256, 0, s>>>(kind, X1, X2, (const int32_t*)pIdx, key, key_ptr, legacy, n, row0, rows, d,
                                                      (T)stretch, (const T*)pShift, Q, (T*)pExtra);
  HX_LAUNCH_CHECK("k_vanilla_propose");
  if (int e = log_prob_t<T>(ctx, tgt, Q, rows, lp_out, nullptr, s)) return e;
  k_vanilla_finish<T><<<(rows + 7) / 8, 256, 0, s>>>(X1, Q, lp1, lp_out, (const T*)pExtra, logacc, do_accept, key, key_ptr, legacy, n,
                                                     row0, rows, d, accepts, chain_out, chain_lp, nonfinite);
  HX_LAUNCH_CHECK("k_vanilla_finish");
  ctx->launches += 5;
  return 0;
}

extern "C" HX_API int hx_vanilla_move(hx_ctx* ctx, int dtype, int impl, int kind, const hx_target* tgt, const void* X1, const void* X2,
                               int32_t n, int32_t row0, int32_t rows, double stretch, const uint32_t key[2], const void* lp1,
                               void* Q, void* logacc, void* lp_out, void* stream) {
  if (!ctx || !tgt || !key || !X1 || !X2 || !lp1 || !Q || !logacc || !lp_out) return fail(HX_E_NULL, "hx_vanilla_move: NULL argument");
  if (int e = check_dtype_impl(dtype, impl, "hx_vanilla_move")) return e;
  if (kind < kVanillaStretch || kind > kVanillaWalk) return fail(HX_E_DTYPE, "hx_vanilla_move: kind must be 0 (stretch), 1 (side) or 2 (walk)");
  if (n < 2 || row0 < 0 || rows < 1 || row0 + rows > n) return fail(HX_E_SHAPE, "hx_vanilla_move: bad row shard [%d, %d) of %d", row0, row0 + rows, n);
  if (kind == kVanillaStretch && !(stretch >= 1.0)) return fail(HX_E_SHAPE, "hx_vanilla_move: stretch must be >= 1");
  HX_CUDA(cudaSetDevice(ctx->device));
  const Key k{key[0], key[1]};
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == HX_F32)
    return vanilla_half_t<float>(ctx, impl, kind, tgt, (float*)const_cast<void*>(X1), (const float*)X2, n, row0, rows, stretch, k, nullptr,
                                 (float*)const_cast<void*>(lp1), (float*)Q, (float*)logacc, (float*)lp_out, 0, nullptr, nullptr, nullptr,
                                 nullptr, s);
  return vanilla_half_t<double>(ctx, impl, kind, tgt, (double*)const_cast<void*>(X1), (const double*)X2, n, row0, rows, stretch, k, nullptr,
                                (double*)const_cast<void*>(lp1), (double*)Q, (double*)logacc, (double*)lp_out, 0, nullptr, nullptr,
                                nullptr, nullptr, s);
}

template <typename T>
static int run_vanilla_t(hx_ctx* ctx, int impl, int kind, const hx_target* tgt, T* X, T* lp, int n, double stretch, const uint32_t* keys,
                         int64_t Tn, T* chain, T* chain_lp, int32_t* accepts, int32_t* nonfinite, cudaStream_t s) {
  const int d = tgt->dim;
  void *pQ, *pLa, *pLp;
  if (int e = ws_get(ctx, WS_Q, (size_t)n * d * sizeof(T), &pQ)) return e;
  if (int e = ws_get(ctx, WS_AD_STAT, (size_t)n * sizeof(T), &pLa)) return e;
  if (int e = ws_get(ctx, WS_LP, (size_t)n * sizeof(T), &pLp)) return e;
  const size_t W2 = (size_t)2 * n;
  for (int64_t t = 0; t < Tn; ++t) {
    for (int half = 0; half < 2; ++half) {
      const uint32_t* k = keys + (size_t)t * 4 + 2 * half;      // keys[t, half]: move AND accept (ensemble.py:100-109)
      T* Xa = X + (size_t)half * n * d;
      T* Xb = X + (size_t)(1 - half) * n * d;
      if (int e = vanilla_half_t<T>(ctx, impl, kind, tgt, Xa, Xb, n, 0, n, stretch, Key{0, 0}, k, lp + (size_t)half * n, (T*)pQ, (T*)pLa,
                                    (T*)pLp, 1, accepts + (size_t)half * n,
                                    chain ? chain + ((size_t)t * W2 + (size_t)half * n) * d : nullptr,
                                    chain_lp ? chain_lp + (size_t)t * W2 + (size_t)half * n : nullptr, nonfinite, s))
        return e;
    }
    if (ctx->stats_on)
      if (int e = stats_feed(ctx, sizeof(T) == 4 ? HX_F32 : HX_F64, &ctx->stats, X, s)) return e;
  }
  return 0;
}

extern "C" HX_API int hx_run_vanilla_iterations(hx_ctx* ctx, int dtype, int impl, int kind, const hx_target* tgt, void* X, void* lp,
                                         int32_t n, double stretch, const uint32_t* keys_dev, int64_t T, void* chain_out,
                                         void* lp_out, int32_t* accepts, int32_t* nonfinite, void* stream) {
  if (!ctx || !tgt || !X || !lp || !keys_dev || !accepts || !nonfinite) return fail(HX_E_NULL, "hx_run_vanilla_iterations: NULL argument");
  if (int e = check_dtype_impl(dtype, impl, "hx_run_vanilla_iterations")) return e;
  if (kind < kVanillaStretch || kind > kVanillaWalk) return fail(HX_E_DTYPE, "hx_run_vanilla_iterations: kind must be 0, 1 or 2");
  if (n < 2 || T < 0) return fail(HX_E_SHAPE, "hx_run_vanilla_iterations: need n >= 2, T >= 0");
  if (kind == kVanillaStretch && !(stretch >= 1.0)) return fail(HX_E_SHAPE, "hx_run_vanilla_iterations: stretch must be >= 1");
  HX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == HX_F32)
    return run_vanilla_t<float>(ctx, impl, kind, tgt, (float*)X, (float*)lp, n, stretch, keys_dev, T, (float*)chain_out, (float*)lp_out,
                                accepts, nonfinite, s);
  return run_vanilla_t<double>(ctx, impl, kind, tgt, (double*)X, (double*)lp, n, stretch, keys_dev, T, (double*)chain_out, (double*)lp_out,
                               accepts, nonfinite, s);
}

// ---- (4b) warm-up scan with live adapters: adapter state, value() and update() on the device (hx_adapt.cuh) ---------------------
template <typename T>
static int column_mean(hx_ctx* ctx, const T* X, int rows, int d, T* mean_out, cudaStream_t s) {
  void* pPart;
  const int nchunks = (rows + kColChunk - 1) / kColChunk;
  if (int e = ws_get(ctx, WS_AD_PART, (size_t)nchunks * d * sizeof(T), &pPart)) return e;
  dim3 g1((d + 127) / 128, nchunks);
  k_colsum_partial<T><<<g1, 128, 0, s>>>(X, rows, d, (T*)pPart);
  HX_LAUNCH_CHECK("k_colsum_partial");
  k_mean_final<T><<<(d + 31) / 32, 256, 0, s>>>((const T*)pPart, nchunks, rows, d, mean_out);
  HX_LAUNCH_CHECK("k_mean_final");
  return 0;
}

template <typename T>
static int run_warmup_t(hx_ctx* ctx, int impl, int move_kind, const hx_target* tgt, T* X, T* lp, int n, const hx_adapt_params& P,
                        hx_adapt_state* st_host, const uint32_t* keys, int64_t Tn, T* chain, T* chain_lp, int32_t* accepts,
                        int32_t* nonfinite, cudaStream_t s) {
  const int d = tgt->dim, ldd = round4(d);
  // Problems the precision policy sends to the tcgen05 path: the adapter state and its update kernels stay on the device, but
  // the launch sequence of that path depends on (step, L), so the two words are read back once per half-move (no other host
  // work; the Python loop this replaces spent 165 ms per iteration at c5 against 16 ms for the moves themselves)
  bool tcp = false;
  if constexpr (std::is_same<T, float>::value) tcp = move_kind == 0 && tc::walk_eligible(ctx, HX_F32, tgt, n, n, 0);
  if (P.use_chees && d > kAdaptMaxDim && !tcp)
    return fail(HX_E_UNSUPPORTED, "hx_run_warmup: ChEES on the device needs dim <= %d (got %d) unless the tensor-core path applies", kAdaptMaxDim, d);
  void *pQ, *pLa, *pLp, *pS, *pSt, *pMean, *pChol, *pStat;
  if (int e = ws_get(ctx, WS_AD_PROP, (size_t)n * d * sizeof(T), &pQ)) return e;
  if (int e = ws_get(ctx, WS_LA, (size_t)n * sizeof(T), &pLa)) return e;
  if (int e = ws_get(ctx, WS_LP, (size_t)n * sizeof(T), &pLp)) return e;
  if (int e = ws_get(ctx, WS_S, (size_t)n * d * sizeof(T), &pS)) return e;
  if (int e = ws_get(ctx, WS_AD_STATE, sizeof(AdaptDev<T>), &pSt)) return e;
  if (int e = ws_get(ctx, WS_AD_MEAN, (size_t)2 * d * sizeof(T), &pMean)) return e;
  const bool big = P.use_chees && d > kAdaptMaxDim;
  const size_t ldw = (size_t)((d + spd::NB - 1) / spd::NB) * spd::NB;
  if (int e = ws_get(ctx, WS_AD_CHOL, big ? 2 * ldw * ldw * sizeof(double) : (size_t)d * d * sizeof(T), &pChol)) return e;
  if (int e = ws_get(ctx, WS_AD_STAT, (size_t)3 * n * sizeof(T), &pStat)) return e;
  AdaptDev<T>* st = (AdaptDev<T>*)pSt;
  T* mean_prev = (T*)pMean;
  T* mean_prop = (T*)pMean + d;
  T* stat_da = (T*)pStat;
  T* stat_acc = stat_da + n;
  T* stat_accg = stat_acc + n;
  AdaptDev<T> h;
  memset(&h, 0, sizeof(h));
  h.da_it = (int)st_host->da_iteration; h.da_ls = (T)st_host->da_log_step; h.da_lsb = (T)st_host->da_log_step_bar;
  h.da_H = (T)st_host->da_H_bar; h.ch_lT = (T)st_host->ch_log_T; h.ch_lTb = (T)st_host->ch_log_T_bar; h.ch_m1 = (T)st_host->ch_m1;
  h.ch_m2 = (T)st_host->ch_m2; h.ch_it = (T)st_host->ch_iteration; h.ch_halton = (float)st_host->ch_halton;
  HX_CUDA(cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice, s));
  HX_CUDA(cudaStreamSynchronize(s));                    // `h` is a stack object
  k_adapt_value<T><<<1, 32, 0, s>>>(st, P);
  HX_LAUNCH_CHECK("k_adapt_value");
  if (!tcp) {
    ctx->ctl_eps = &st->eps;
    ctx->ctl_L = &st->L;
  }
  const size_t W2 = (size_t)2 * n;
  int err = 0;
  for (int64_t t = 0; t < Tn && !err; ++t) {
    const uint32_t* kt = keys + (size_t)t * 8;
    for (int half = 0; half < 2 && !err; ++half) {
      // warm-up key order (samplers/hamiltonian_ensemble.py:171,184,191,204)
      const uint32_t* kmove = kt + 2 * (half == 0 ? 0 : 2);
      const uint32_t* kacc = kt + 2 * (half == 0 ? 1 : 3);
      T* Xa = X + (size_t)half * n * d;
      T* Xb = X + (size_t)(1 - half) * n * d;
      T* lpa = lp + (size_t)half * n;
      if (tcp) {
        AdaptDev<T> cur;
        cudaError_t ce = cudaMemcpyAsync(&cur, st, sizeof(cur), cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
        if (ce != cudaSuccess) { err = fail((int)ce, "hx_run_warmup: reading (step, L) back failed: %s", cudaGetErrorString(ce)); break; }
        err = walk_move_t<T>(ctx, impl, tgt, Xa, Xb, n, 0, n, (double)cur.eps, Key{0, 0}, kmove, cur.L, lpa, (T*)pQ, (T*)pLa, (T*)pS, (T*)pLp,
                             0, nullptr, nullptr, s);
      } else if (move_kind == 0)
        err = walk_move_t<T>(ctx, impl, tgt, Xa, Xb, n, 0, n, 0.0, Key{0, 0}, kmove, 1, lpa, (T*)pQ, (T*)pLa, (T*)pS, (T*)pLp, 0, nullptr, nullptr, s);
      else
        err = side_move_t<T>(ctx, impl, tgt, Xa, Xb, n, 0, n, 0.0, Key{0, 0}, kmove, 1, lpa, (T*)pQ, (T*)pLa, (T*)pS, (T*)pLp, 0, nullptr, s);
      if (err) break;
      // adapter.update(state, log_accept, position_current = X (pre-accept), position_proposed, momentum_proposed, group2)  :176-182
      const T* Lf = nullptr;
      if (P.use_chees) {
        T* M = (T*)ctx->buf[WS_M];
        if (move_kind != 0) {        // the side move does not form the Gram matrix of the complement
          T* C;
          if ((err = center_complement<T>(ctx, Xb, n, d, ldd, &C, s))) break;
          if ((err = gram_matrix<T>(ctx, C, n, ldd, &M, s))) break;
        }
        if ((err = column_mean<T>(ctx, Xa, n, d, mean_prev, s))) break;
        if ((err = column_mean<T>(ctx, (const T*)pQ, n, d, mean_prop, s))) break;
        if (!big) {
          k_chol<T><<<1, 256, 0, s>>>(M, tcp ? d : ldd, d, (T)n / (T)(n - 1), (T*)pChol, d);      // the tensor-core path keeps M unpadded
          Lf = (const T*)pChol;
        }
      }
      bool rows_done = false;
      if constexpr (std::is_same<T, float>::value) {
        if (big) {
          // cov(group2)^-1 = (n / (n - 1) M)^-1 in fp64 once per half-move, then ONE tensor-core GEMM with a row-dot epilogue
          // gives <cen_prop_i, cov^-1 pproj_i> for all walkers (chees.py:57-73 solves with n right-hand sides instead)
          void* pInv;
          if ((err = ws_get(ctx, WS_TC_BO, (size_t)d * d * sizeof(float), &pInv))) break;
          if ((err = spd::spd_inverse(ctx, (const float*)ctx->buf[WS_M], d, d, (double)n / (double)(n - 1), (double*)pChol, (float*)pInv, d, s))) break;
          float* part = nullptr; int nt = 0;
          if ((err = tc::chees_inner(ctx, (const float*)pS, n, d, (const float*)pInv, (const float*)pQ, mean_prop, &part, &nt, s))) break;
          k_adapt_rows_big<<<(n * 32 + 255) / 256, 256, 0, s>>>((const float*)pLa, Xa, (const float*)pQ, mean_prev, mean_prop, part, nt, n, d,
                                                                st, P, stat_da, stat_acc, stat_accg);
          rows_done = true;
        }
      }
      if (!rows_done)
        k_adapt_rows<T><<<(n + 127) / 128, 128, 0, s>>>((const T*)pLa, Xa, (const T*)pQ, (const T*)pS, mean_prev, mean_prop, Lf, d, n, d,
                                                        st, P, stat_da, stat_acc, stat_accg);
      k_adapt_update<T><<<1, 1024, 0, s>>>(st, P, stat_da, stat_acc, stat_accg, n);
      k_accept_emit<T><<<(n + 7) / 8, 256, 0, s>>>(Xa, (const T*)pQ, lpa, (const T*)pLp, (const T*)pLa, kacc, impl == HX_RNG_LEGACY ? 1 : 0,
                                                   n, 0, n, d, accepts + (size_t)half * n,
                                                   chain ? chain + ((size_t)t * W2 + (size_t)half * n) * d : nullptr,
                                                   chain_lp ? chain_lp + (size_t)t * W2 + (size_t)half * n : nullptr, nonfinite);
      cudaError_t ce = cudaGetLastError();
      if (ce != cudaSuccess) err = fail((int)ce, "hx_run_warmup: launch failed: %s", cudaGetErrorString(ce));
      ctx->launches += 8;
    }
  }
  ctx->ctl_eps = nullptr;
  ctx->ctl_L = nullptr;
  if (err) return err;
  HX_CUDA(cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, s));
  HX_CUDA(cudaStreamSynchronize(s));
  st_host->da_iteration = h.da_it; st_host->da_log_step = (double)h.da_ls; st_host->da_log_step_bar = (double)h.da_lsb;
  st_host->da_H_bar = (double)h.da_H; st_host->ch_log_T = (double)h.ch_lT; st_host->ch_log_T_bar = (double)h.ch_lTb;
  st_host->ch_m1 = (double)h.ch_m1; st_host->ch_m2 = (double)h.ch_m2; st_host->ch_iteration = (double)h.ch_it;
  st_host->ch_halton = (double)h.ch_halton;
  return 0;
}

extern "C" HX_API int hx_spd_inverse(hx_ctx* ctx, const float* M_dev, int32_t ldm, int32_t d, double scale, float* inv_out_dev, int32_t ldo,
                              void* stream) {
  if (!ctx || !M_dev || !inv_out_dev) return fail(HX_E_NULL, "hx_spd_inverse: NULL argument");
  if (d < 1 || ldm < d || ldo < d) return fail(HX_E_SHAPE, "hx_spd_inverse: need d >= 1 and leading dimensions >= d");
  HX_CUDA(cudaSetDevice(ctx->device));
  const size_t ldw = (size_t)((d + spd::NB - 1) / spd::NB) * spd::NB;
  void* work;
  if (int e = ws_get(ctx, WS_AD_CHOL, 2 * ldw * ldw * sizeof(double), &work)) return e;
  return spd::spd_inverse(ctx, M_dev, ldm, d, scale, (double*)work, inv_out_dev, ldo, (cudaStream_t)stream);
}

extern "C" HX_API int hx_run_warmup(hx_ctx* ctx, int dtype, int impl, int move_kind, const hx_target* tgt, void* X, void* lp, int32_t n,
                             const hx_adapt_params* params, hx_adapt_state* state, const uint32_t* keys_dev, int64_t T,
                             void* chain_out, void* lp_out, int32_t* accepts, int32_t* nonfinite, void* stream) {
  if (!ctx || !tgt || !X || !lp || !params || !state || !keys_dev || !accepts || !nonfinite) return fail(HX_E_NULL, "hx_run_warmup: NULL argument");
  if (int e = check_dtype_impl(dtype, impl, "hx_run_warmup")) return e;
  if (move_kind != 0 && move_kind != 1) return fail(HX_E_DTYPE, "hx_run_warmup: move_kind must be 0 (walk) or 1 (side)");
  if (!params->use_da && !params->use_chees) return fail(HX_E_DTYPE, "hx_run_warmup: no adapter enabled (use hx_run_iterations with key_order 1)");
  if (n < 2 || T < 0) return fail(HX_E_SHAPE, "hx_run_warmup: need n >= 2, T >= 0");
  HX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == HX_F32)
    return run_warmup_t<float>(ctx, impl, move_kind, tgt, (float*)X, (float*)lp, n, *params, state, keys_dev, T, (float*)chain_out,
                               (float*)lp_out, accepts, nonfinite, s);
  return run_warmup_t<double>(ctx, impl, move_kind, tgt, (double*)X, (double*)lp, n, *params, state, keys_dev, T, (double*)chain_out,
                              (double*)lp_out, accepts, nonfinite, s);
}

// ---- (5) row-sharded multi-GPU batch loop: exchange over NVLink peer memory (hx_dist.cuh) -------------------------------------
extern "C" HX_API int hx_dist_init(hx_ctx* ctx, int dtype, int32_t rank, int32_t world, int32_t n, int32_t dim, void* handle_out) {
  if (!ctx || !handle_out) return fail(HX_E_NULL, "hx_dist_init: NULL argument");
  if (dtype != HX_F32 && dtype != HX_F64) return fail(HX_E_DTYPE, "hx_dist_init: unknown dtype %d", dtype);
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(HX_E_SHAPE, "hx_dist_init: bad rank %d / world %d (max %d)", rank, world, kMaxRanks);
  if (n < 2 || dim < 1 || n % world) return fail(HX_E_SHAPE, "hx_dist_init: group size %d must be a multiple of the world size %d", n, world);
  const size_t esz = dtype == HX_F32 ? 4 : 8;
  if (((size_t)(n / world) * dim * esz) % 16 || ((size_t)(n / world) * esz) % 16)
    return fail(HX_E_ALIGN, "hx_dist_init: a rank's rows (%d x %d) and log-probs must be multiples of 16 bytes", n / world, dim);
  HX_CUDA(cudaSetDevice(ctx->device));
  HxDist& D = ctx->dist;
  if (D.on) return fail(HX_E_UNSUPPORTED, "hx_dist_init: already initialised (call hx_dist_destroy first)");
  memset(&D, 0, sizeof(D));
  const size_t xb = (size_t)2 * n * dim * esz, lb = (size_t)2 * n * esz;
  D.lp_off = (xb + 255) & ~(size_t)255;
  D.flag_off = (D.lp_off + lb + 255) & ~(size_t)255;
  D.counter_off = D.flag_off + 256;
  D.bytes = D.counter_off + 256;
  void* base = nullptr;
  HX_CUDA(cudaMalloc(&base, D.bytes));
  HX_CUDA(cudaMemset(base, 0, D.bytes));
  HX_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, base);
  if (e != cudaSuccess) { cudaFree(base); return fail((int)e, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, 64);
  D.base[rank] = base;
  D.rank = rank; D.world = world; D.n = n; D.dim = dim; D.dtype = dtype; D.epoch = 0; D.on = 1;
  return 0;
}

extern "C" HX_API int hx_dist_connect(hx_ctx* ctx, const void* handles) {
  if (!ctx || !handles) return fail(HX_E_NULL, "hx_dist_connect: NULL argument");
  HxDist& D = ctx->dist;
  if (!D.on) return fail(HX_E_UNSUPPORTED, "hx_dist_connect: hx_dist_init first");
  HX_CUDA(cudaSetDevice(ctx->device));
  for (int p = 0; p < D.world; ++p) {
    if (p == D.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)p * 64, 64);
    void* ptr = nullptr;
    HX_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    D.base[p] = ptr;
  }
  D.on = 2;
  return 0;
}

extern "C" HX_API int hx_dist_buffers(hx_ctx* ctx, void** X_full, void** lp_full) {
  if (!ctx || !ctx->dist.on) return fail(HX_E_UNSUPPORTED, "hx_dist_buffers: not initialised");
  if (X_full) *X_full = ctx->dist.base[ctx->dist.rank];
  if (lp_full) *lp_full = (char*)ctx->dist.base[ctx->dist.rank] + ctx->dist.lp_off;
  return 0;
}

extern "C" HX_API int hx_dist_destroy(hx_ctx* ctx) {
  if (!ctx || !ctx->dist.on) return 0;
  HxDist& D = ctx->dist;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < D.world; ++p)
    if (p != D.rank && D.base[p]) cudaIpcCloseMemHandle(D.base[p]);
  if (D.base[D.rank]) cudaFree(D.base[D.rank]);
  memset(&D, 0, sizeof(D));
  return 0;
}

template <typename T>
static int run_iterations_sharded_t(hx_ctx* ctx, int impl, int move_kind, int key_order, const hx_target* tgt, double eps, int L,
                                    const uint32_t* keys, int64_t Tn, T* chain, T* chain_lp, int32_t* accepts, int32_t* nonfinite,
                                    cudaStream_t s) {
  HxDist& D = ctx->dist;
  const int n = D.n, d = D.dim, world = D.world, rank = D.rank;
  const int rows = n / world, row0 = rank * rows;
  T* Xf = (T*)D.base[rank];
  T* lpf = (T*)((char*)D.base[rank] + D.lp_off);
  const uint32_t* flags = (const uint32_t*)((char*)D.base[rank] + D.flag_off);
  uint32_t* counter = (uint32_t*)((char*)D.base[rank] + D.counter_off);
  PeerTable pt;
  for (int p = 0; p < kMaxRanks; ++p) pt.base[p] = p < world ? D.base[p] : nullptr;
  void *pQ, *pLa, *pLp;
  if (int e = ws_get(ctx, WS_Q, (size_t)rows * d * sizeof(T), &pQ)) return e;
  if (int e = ws_get(ctx, WS_LA, (size_t)rows * sizeof(T), &pLa)) return e;
  if (int e = ws_get(ctx, WS_LP, (size_t)rows * sizeof(T), &pLp)) return e;
  int clock_khz = 1900000;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, ctx->device);
  const long long timeout = (long long)clock_khz * 1000ll * 20ll;      // ~20 s of SM clocks
  for (int64_t t = 0; t < Tn; ++t) {
    const uint32_t* kt = keys + (size_t)t * 8;
    for (int half = 0; half < 2; ++half) {
      const int im = key_order == 0 ? (half == 0 ? 0 : 1) : (half == 0 ? 0 : 2);
      const int ia = key_order == 0 ? (half == 0 ? 2 : 3) : (half == 0 ? 1 : 3);
      // the complement (and, from the previous iteration, this half) must be complete on this rank
      if (D.epoch > 0) {
        if (int e = timing_mark(ctx, s, HX_PH_WAIT)) return e;
        k_dist_wait<<<1, 32, 0, s>>>(flags, world, D.epoch, timeout, nonfinite);
        HX_LAUNCH_CHECK("k_dist_wait");
      }
      if (int e = timing_mark(ctx, s, HX_PH_STATS)) return e;
      T* Xa = Xf + ((size_t)half * n + row0) * d;
      T* Xb = Xf + (size_t)(1 - half) * n * d;
      T* lpa = lpf + (size_t)half * n + row0;
      Fuse<T> f;
      f.akey_ptr = kt + 2 * ia; f.lp1w = lpa; f.accepts = accepts + (size_t)half * rows;
      const int64_t slot = chain_slot(ctx, t);
      f.chain_out = (chain && slot >= 0) ? chain + ((size_t)slot * 2 * rows + (size_t)half * rows) * d : nullptr;
      f.chain_lp = (chain_lp && slot >= 0) ? chain_lp + (size_t)slot * 2 * rows + (size_t)half * rows : nullptr;
      f.nonfinite = nonfinite;
      // exchange: the rank's updated rows (+ log-probs) are stored into every peer's copy and the epoch is published, either by
      // the tensor-core path's finish kernel itself or by k_dist_push after the SIMT half-move
      D.epoch += 1;
      const size_t x_off = ((size_t)half * n + row0) * d * sizeof(T), cnt = (size_t)rows * d * sizeof(T) / 16;
      const size_t l_off = D.lp_off + ((size_t)half * n + row0) * sizeof(T), lcnt = (size_t)rows * sizeof(T) / 16;
      DistPush dp;
      dp.pt = pt; dp.rank = rank; dp.world = world; dp.x_off = x_off; dp.lp_off = l_off; dp.flag_off = D.flag_off;
      dp.counter = counter; dp.epoch = D.epoch;
      bool fused_push = false;
      if constexpr (std::is_same<T, float>::value) fused_push = move_kind == 0 && tc::walk_eligible(ctx, HX_F32, tgt, n, rows, 0);
      f.push = fused_push ? &dp : nullptr;
      tc::NextDraw nd;
      const tc::NextDraw* next = nullptr;
      if (half == 0 || t + 1 < Tn) {
        nd.key = Key{0, 0}; nd.key_ptr = half == 0 ? kt + 2 * (key_order == 0 ? 1 : 2) : kt + 8; nd.row0 = row0; nd.rows = rows;
        next = &nd;
      }
      int e;
      if (move_kind == 0)
        e = walk_move_t<T>(ctx, impl, tgt, Xa, Xb, n, row0, rows, eps, Key{0, 0}, kt + 2 * im, L, lpa, (T*)pQ, (T*)pLa, nullptr, (T*)pLp, 0, &f, next, s);
      else
        e = side_move_t<T>(ctx, impl, tgt, Xa, Xb, n, row0, rows, eps, Key{0, 0}, kt + 2 * im, L, lpa, (T*)pQ, (T*)pLa, nullptr, (T*)pLp, 0, &f, s);
      if (e) return e;
      if (!fused_push) {
        if (int e2 = timing_mark(ctx, s, HX_PH_PUSH)) return e2;
        int grid = (int)((cnt + 255) / 256);
        if (grid > ctx->sm_count * 2) grid = ctx->sm_count * 2;
        if (grid < 1) grid = 1;
        k_dist_push<<<grid, 256, 0, s>>>(pt, rank, world, x_off, cnt, l_off, lcnt, D.flag_off, counter, D.epoch);
        HX_LAUNCH_CHECK("k_dist_push");
      }
      if (int e2 = timing_mark(ctx, s, -1)) return e2;
      ctx->launches += 2;
    }
  }
  if (D.epoch > 0) {
    k_dist_wait<<<1, 32, 0, s>>>(flags, world, D.epoch, timeout, nonfinite);
    HX_LAUNCH_CHECK("k_dist_wait");
  }
  return 0;
}

extern "C" HX_API int hx_run_iterations_sharded(hx_ctx* ctx, int dtype, int impl, int move_kind, int key_order, const hx_target* tgt,
                                         double step_size, int32_t L, const uint32_t* keys_dev, int64_t T, void* chain_out,
                                         void* lp_out, int32_t* accepts, int32_t* nonfinite, void* stream) {
  if (!ctx || !tgt || !keys_dev || !accepts || !nonfinite) return fail(HX_E_NULL, "hx_run_iterations_sharded: NULL argument");
  if (ctx->dist.on != 2) return fail(HX_E_UNSUPPORTED, "hx_run_iterations_sharded: hx_dist_init + hx_dist_connect first");
  if (int e = check_dtype_impl(dtype, impl, "hx_run_iterations_sharded")) return e;
  if (dtype != ctx->dist.dtype || tgt->dim != ctx->dist.dim) return fail(HX_E_SHAPE, "hx_run_iterations_sharded: dtype/dim differ from hx_dist_init");
  if (move_kind != 0 && move_kind != 1) return fail(HX_E_DTYPE, "hx_run_iterations_sharded: move_kind must be 0 (walk) or 1 (side)");
  if (key_order != 0 && key_order != 1) return fail(HX_E_DTYPE, "hx_run_iterations_sharded: key_order must be 0 or 1");
  if (L < 1 || T < 0) return fail(HX_E_SHAPE, "hx_run_iterations_sharded: need L >= 1, T >= 0");
  HX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == HX_F32)
    return run_iterations_sharded_t<float>(ctx, impl, move_kind, key_order, tgt, step_size, L, keys_dev, T, (float*)chain_out,
                                           (float*)lp_out, accepts, nonfinite, s);
  return run_iterations_sharded_t<double>(ctx, impl, move_kind, key_order, tgt, step_size, L, keys_dev, T, (double*)chain_out,
                                          (double*)lp_out, accepts, nonfinite, s);
}
