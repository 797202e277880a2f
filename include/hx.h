/*
 * hx.h — C-ABI of libhx.so, the B200-native (sm_100a) implementation of h-emcee's
 * ensemble-Hamiltonian hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own
 * (pure JAX); every entry point below replaces one Python-level interface of the
 * reference, cited as `file:line` relative to the reference repository root.  The
 * Python mirror of the reference API (h-emcee_b200/hemcee_b200) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross the boundary.
 *   - every `*_dev` / `void*` state argument is a DEVICE pointer owned by the caller
 *     (e.g. torch CUDA tensor .data_ptr()); the library allocates only private scratch.
 *   - walker state is row-major [rows, dim] exactly like the reference's arrays.
 *   - `dtype`: HX_F32 = JAX default mode, HX_F64 = `jax_enable_x64` (changes the RNG
 *     stream: 64 random bits per draw).
 *   - `impl`: threefry counter layout, HX_RNG_PARTITIONABLE (JAX >= 0.5 default) or
 *     HX_RNG_LEGACY.  f32 normals exist in two forms: exact (w = -log1p(-u^2) correctly rounded, <= 1 ulp from
 *     the NumPy oracle of XLA's erf_inv) and fast (SFU log2, <= 3 ulp, ~84 % of the draws bit-equal; used by the
 *     tensor-core path and by default).  ctx-based calls draw exact normals under HX_PREC_EXACT; the ctx-free
 *     hx_normal takes the request as `impl | HX_RNG_EXACT_NORMALS`.
 *   - keys are raw uint32[2] on the HOST unless the name says `_dev`.
 *   - all launches are asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - return value: 0 ok; <0 invalid argument (HX_E_*); >0 a cudaError_t.  The message
 *     is available from hx_last_error() (thread-local).  Nothing throws.
 *   - a ctx is bound to one device and is not thread-safe (one host thread per ctx,
 *     like the reference's single-threaded driver).
 */
#ifndef HX_H_
#define HX_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HX_ABI_VERSION 1

#if defined(__GNUC__)
#define HX_API __attribute__((visibility("default")))
#else
#define HX_API
#endif

enum { HX_F32 = 0, HX_F64 = 1 };
enum { HX_RNG_PARTITIONABLE = 0, HX_RNG_LEGACY = 1, HX_RNG_EXACT_NORMALS = 16 };

/* error codes (<0) */
enum {
  HX_E_NULL = -1,      /* required pointer is NULL */
  HX_E_SHAPE = -2,     /* bad n / dim / rows / L */
  HX_E_DTYPE = -3,     /* unknown dtype / impl / flag */
  HX_E_TARGET = -4,    /* malformed target descriptor */
  HX_E_UNSUPPORTED = -5,
  HX_E_ALIGN = -6
};

/* Built-in targets.  The reference differentiates an arbitrary callable with
 * jax.vmap(jax.grad(log_prob)) (src/hemcee/samplers/sampler.py:52,55); this library ships
 * analytic log-density + gradient for the reference's own test targets
 * (src/hemcee/tests/distribution.py:30-48,76-97,143-156; funnel per SURVEY §8a row T). */
enum {
  HX_TARGET_GAUSS_ISO = 0,   /* -1/2 |x - mean|^2                      (quickstart.ipynb cell 4) */
  HX_TARGET_GAUSS_FULL = 1,  /* -1/2 (x-mean)^T P (x-mean), P = precision[dim, ld] symmetric    */
  HX_TARGET_ROSENBROCK = 2,  /* -sum_i b (x_{i+1}-x_i^2)^2 + (a-x_i)^2  (distribution.py:155-156) */
  HX_TARGET_FUNNEL = 3       /* x0=nu: -nu^2/(2 a^2) - 1/2 e^{-nu} sum x_i^2 - (dim-1) nu/2, a=sigma */
};

typedef struct hx_target {
  int32_t kind;            /* HX_TARGET_* */
  int32_t dim;
  int32_t ld;              /* leading dimension (elements) of `precision`; multiple of 4, >= dim */
  int32_t has_lower0;      /* if non-zero: log-density is -inf where x[0] < lower0 (hard constraint,
                              reference src/hemcee/tests/test_log_prob_validation.py:95-155) */
  const void* mean;        /* device [dim] in the run dtype, or NULL for zero mean */
  const void* precision;   /* device [dim, ld] in the run dtype (GAUSS_FULL only), zero padded */
  double a;                /* rosenbrock a / funnel sigma */
  double b;                /* rosenbrock b */
  double lower0;
} hx_target;

/* flags for the move entry points */
enum {
  HX_MOVE_FROZEN_GRAD = 1, /* use grad(log_prob) at the initial positions for every kick
                              (src/hemcee/adaptation/reasonable_stepsize.py:34-38) */
};

typedef struct hx_ctx hx_ctx;

HX_API int hx_abi_version(void);
HX_API const char* hx_last_error(void);

/* ctx = per-device scratch workspace (grow-only, private) */
HX_API int hx_ctx_create(hx_ctx** out, int device);
HX_API int hx_ctx_destroy(hx_ctx* ctx);
/* bytes of private scratch currently held (diagnostics) */
HX_API size_t hx_ctx_workspace_bytes(const hx_ctx* ctx);
/* Precision policy of the walk move's dense contractions:
 *   HX_PREC_AUTO   exact SIMT fp32/fp64 kernels, except fp32 full-covariance Gaussian targets with
 *                  dim >= 128 and >= 2048 walkers per group, which use HX_PREC_F16F8
 *   HX_PREC_EXACT  always the exact SIMT kernels (fp32 FMA / fp64)
 *   HX_PREC_BF16X3 tcgen05 tensor cores, operands split into bf16 hi+lo, three products, fp32
 *                  accumulation in TMEM (fp32-class accuracy, ~2^-16 per product)
 *   HX_PREC_BF16   tcgen05, single bf16 product for the momentum projection (statistical parity only)
 *   HX_PREC_TF32X3 like BF16X3 but the leapfrog kicks / energy products use tf32 hi+lo operands
 *                  (~2^-21 per product) — for targets conditioned worse than ~1e6
 *   HX_PREC_F16F8  like BF16X3 but the momentum projection takes fp16(x) for the main product and e4m3(x),
 *                  e4m3((x - fp16(x)) 2^11) for the two correction products (kind::f8f6f4 at twice the rate;
 *                  the complement is scaled by a power of two into the fp16 / e4m3 range): 8 instead of 12
 *                  MMA slots per K-block at ~2^-15 per product (measured |dS0| 1.2e-5 vs 0.7e-5 relative)  */
enum { HX_PREC_AUTO = 0, HX_PREC_EXACT = 1, HX_PREC_BF16X3 = 2, HX_PREC_BF16 = 3, HX_PREC_TF32X3 = 4, HX_PREC_F16F8 = 5 };
HX_API int hx_ctx_set_precision(hx_ctx* ctx, int mode);
/* Form of the walk move's leapfrog on the tensor-core path (full-covariance Gaussian targets):
 *   HX_LEAP_STEPWISE    one kick GEMM per leapfrog step on every walker (hmc_walk.py:85-102 as written)
 *   HX_LEAP_PROPAGATOR  the dynamics of a Gaussian target are linear, so the L-step map of (q - mean, p C) is a
 *                       2dim x 2dim matrix: integrate 2*dim basis trajectories with the same recurrence, then apply the
 *                       map to all walkers in one GEMM (cost independent of L per walker)
 *   HX_LEAP_AUTO        whichever a cost model fitted to B200 timings predicts to be faster (propagator for many rows / long L)                                          */
enum { HX_LEAP_AUTO = 0, HX_LEAP_STEPWISE = 1, HX_LEAP_PROPAGATOR = 2 };
HX_API int hx_ctx_set_leapfrog_form(hx_ctx* ctx, int form);
/* Diagnostics for the roofline discussion: wall time [ms] of `reps` repetitions of (mode 0) the momentum draw
 * of a rows x n shard alone, (1) its projection GEMM alone, (2) both running concurrently on two streams.  */
HX_API int hx_tc_microbench(hx_ctx* ctx, int mode /* 0 draw, 1 GEMM, 2 both, 3 fused draw + GEMM kernel */, int32_t n, int32_t rows, int32_t dim, int32_t reps, double* ms_out /*[3]: total,
                     GEMM stream, draw stream*/, void* stream);
/* tuning knobs: what = 0 resident CTAs per SM of the momentum-draw kernel (default 4); what = 1 split-K slices of the
 * projection GEMM, summed in fp32 in a fixed order (default 1 = accumulate all of K = n inside the tensor core; measured:
 * no accuracy gain from more); what = 2: 1 = Gram matrix from tf32 hi/lo operands instead of bf16 hi/lo (default 0);
 * what = 3: operand kind of the propagator-form energy products: 0 = tf32 hi/lo (2^-21 per product, tf32 rate), 1 = bf16 hi/lo
 * (2^-16: too coarse for kappa = 1e4), 2 (default) = fp16 hi/lo with power-of-two scales per walker row / per matrix (2^-22 at
 * the kind::f16 rate, twice the tf32 rate; dim <= 1024, tf32 beyond); what = 4: 0 disables the CUDA-graph
 * replay of launch-bound batch loops (default 1); what = 5: 0 keeps the momentum projection of narrow targets (dim <= 40, n >= 1024)
 * inside the half-move kernel instead of the lane-split projection kernel (default 1);
 * what = 6: the next half-move's momentum draw is queued after that many projection chunks instead of at the projection's start
 * (default 0; measured: any later start runs into the latency-bound launches and costs more than it saves);
 * what = 8: 1 = momentum projection by the CTA-pair kernel (tcgen05 cta_group::2: M = 256 per pair, each CTA stages half of the
 * B tile; 3.15 instead of 3.98 ms alone at c5, bit-identical output), default 0 because it runs WORSE next to the concurrent
 * momentum draw; what = 9: 1 = CTA-pair projection launched one wave at a time even when the draw is
 * complete (default 0 = one launch); what = 10: 1 = draw kernel without its stores (garbage results; overlap diagnostics only);
 * what = 13: 1 = single-CTA projection as ONE launch over all rows when the draw was queued ahead (default 0: one wave per
 * launch; measured: the GEMM ends earlier but the starved draw then runs into the next phases); what = 14: operand kind of the
 * kicks / propagator apply: 0 (default) = fp16 hi/lo with power-of-two row / matrix scales for the stepwise form (log-accept
 * error vs f64 1.1e-2 -> 7.7e-4 rms at c5 scale, same MMA rate) and bf16 hi/lo for the propagator form, 1 = fp16 for both,
 * 2 = bf16 for both;
 * what = 12: 0 = Gram and propagator-apply GEMMs by the single-CTA kernel instead of the CTA-pair kernel
 * (default 1: bit-identical, ~0.1 ms per half-move faster at c5; these launches do not overlap the draw);
 * what = 15: 1 (default) = the momentum draw P0 = normal(key, (n, n)) (hmc_walk.py:36) is generated INSIDE the projection GEMM
 * (hx_tc_fused.cuh: RNG warps write bf16 hi/lo operand tiles into the shared memory of a CTA pair, P0 never reaches HBM) when
 * the problem is eligible (partitionable RNG layout, dim <= 1024, three-product modes), 0 = staged draw kernel + GEMM;
 * what = 16: RNG warps per CTA of the fused kernel (8, 16, 20, 24, 28; 0 = default: 24 for a CTA pair, 28 otherwise); what = 17: split-K slices of the fused kernel
 * (0 = default: 2 from 4096 walkers per group, a function of n only so that row shards stay bit-identical); what = 18: 1 = the staged bf16 hi/lo
 * projection accumulates its three products in ONE TMEM accumulator like the fused kernel (bit-for-bit comparison of the two);
 * what = 19: fused-kernel diagnostics (1 = no draw, 2 = no MMA; garbage results); what = 20: 1 = RNG warps poll for a free stage
 * with test_wait + nanosleep instead of a parked try_wait;
 * what = 21: operand form of the fused projection: 0 (default) by precision (AUTO: fp16 main product + e4m3 x e5m2 correction
 * products at true scale, one accumulator; BF16X3 / TF32X3: bf16 hi/lo, three products, bit-equal to the staged kernel), 1 = bf16
 * hi/lo, 2 = fp16 + fp8;  what = 22: 1 = the two split-K slices of the fused projection go through slice buffers and a sum kernel
 * instead of red.global.add into the zeroed output (same bits);  what = 23: bit mask over the GEMM epilogue kinds (bit 0 STORE,
 * 1 KICKB, 2 DOT, 3 PROP): set = row-per-lane epilogue, clear = 32 x 32 chunks transposed through shared memory for coalesced
 * global accesses (default 0b1100; launches over many rows keep the row path for KICKB);  what = 24: start-up stagger of the fused
 * kernel's RNG warps in ns per K-block position (default 0: measured to change nothing)                                       */
HX_API int hx_ctx_set_tuning(hx_ctx* ctx, int what, int value);
/* which kernel the timing hooks bracketed last: 0 = fused SIMT half-move, 1 = tcgen05 projection GEMM */
HX_API int hx_ctx_last_path(const hx_ctx* ctx);

/* Measurement hooks (bench.py roofline): when enabled, every launch of the dominant half-move kernel
 * is bracketed by CUDA events on its launch stream.  hx_ctx_timing_read synchronises those events and
 * returns the summed kernel time [ms], the number of launches and the number of ALL kernels this
 * library launched through `ctx` since the last reset; then it resets the counters.               */
HX_API int hx_ctx_timing_enable(hx_ctx* ctx, int enable);
HX_API int hx_ctx_timing_read(hx_ctx* ctx, double* kernel_ms, int64_t* kernel_launches, int64_t* all_launches);
/* per-phase breakdown of the last hx_ctx_timing_read (8 entries each): [0] dominant kernel section (fused SIMT half-move, or
 * the projection GEMM on the tensor-core path), [1] complement statistics (centre, Gram, M P), [2] propagator basis run,
 * [3] leapfrog apply / kicks + energy products, [4] finish (energies, accept, chain emit), [5] waiting for the peers' rows,
 * [6] storing the updated rows into the peers (sharded loop), [7] complement mean; summed ms and interval counts */
HX_API int hx_ctx_timing_phases(hx_ctx* ctx, double* phase_ms, int64_t* phase_count);

/* ---- (1) RNG: bit-exact jax.random subset ------------------------------------------------
 * replaces jax.random.split   (call sites: samplers/hamiltonian_ensemble.py:90,228,330;
 *                              moves/hamiltonian/hmc_side.py:37; adaptation/reasonable_stepsize.py:44)
 *          jax.random.normal  (moves/hamiltonian/hmc_walk.py:36; moves/hamiltonian/hmc_side.py:54)
 *          jax.random.uniform (samplers/sampler_utils.py:114)
 *          jax.random.choice  (moves/hamiltonian/hmc_side.py:44-47)                              */
HX_API int hx_split(const uint32_t key[2], int impl, int64_t num, uint32_t* out_dev /*[num,2]*/, void* stream);
HX_API int hx_random_bits(const uint32_t key[2], int impl, int bit_width /*32|64*/, int64_t count,
                   void* out_dev /*uint32|uint64 [count]*/, void* stream);
HX_API int hx_uniform(const uint32_t key[2], int impl, int dtype, double minval, double maxval, int64_t count,
               void* out_dev, void* stream);
HX_API int hx_normal(const uint32_t key[2], int impl, int dtype, int64_t count, void* out_dev, void* stream);
/* per-walker choice(key_i, arange(n), (2,), replace=False) for keys_dev[nkeys,2] -> int32 out_dev[nkeys,2] */
HX_API int hx_choice2(hx_ctx* ctx, const uint32_t* keys_dev, int64_t nkeys, int impl, int32_t n,
               int32_t* out_dev, void* stream);

/* ---- (2) targets ---------------------------------------------------------------------------
 * replaces the jitted vmap(log_prob) / vmap(grad(log_prob)) of samplers/sampler.py:52,55       */
HX_API int hx_log_prob(hx_ctx* ctx, int dtype, const hx_target* tgt, const void* X_dev, int64_t rows,
                void* lp_out_dev, void* grad_out_dev /*nullable [rows,dim]*/, void* stream);

/* ---- (3) move level: mirrors the reference move protocol ------------------------------------
 *   move(group1, group2, step_size, key, log_prob, grad_log_prob, L, log_prob_group1)
 *     -> (proposed, log_accept_prob, momentum_projected, proposed_log_prob)
 * (moves/hamiltonian/hmc_walk.py:6-59, moves/hamiltonian/hmc_side.py:6-77).
 *
 * `n` is the size of BOTH groups (the walk move needs equal halves, hmc_walk.py:36,91).
 * `X1_dev` holds `rows` walkers of group 1 starting at global row `row0` (row0=0, rows=n on
 * one GPU; a row shard otherwise — the RNG counters use the global row so results do not
 * depend on the sharding).  `X2_dev` is the full complement [n, dim].                          */
HX_API int hx_walk_move(hx_ctx* ctx, int dtype, int impl, const hx_target* tgt,
                 const void* X1_dev, const void* X2_dev, int32_t n, int32_t row0, int32_t rows,
                 double step_size, const uint32_t key[2], int32_t L, const void* lp1_dev,
                 void* Q_out_dev, void* logacc_out_dev, void* pproj_out_dev /*nullable*/,
                 void* lp_out_dev, int flags, void* stream);

/* Hint: the walk half-move AFTER the next hx_walk_move call will use (key, row0, rows) with the same n and impl.
 * The momentum matrix normal(key, (n, n)) (hmc_walk.py:36) depends only on the key, so its draw is started on a
 * side stream under the next call's tensor-core work.  Purely an optimisation: results never depend on it.   */
HX_API int hx_walk_prefetch(hx_ctx* ctx, int impl, int32_t n, int32_t row0, int32_t rows, const uint32_t key[2]);

HX_API int hx_side_move(hx_ctx* ctx, int dtype, int impl, const hx_target* tgt,
                 const void* X1_dev, const void* X2_dev, int32_t n, int32_t row0, int32_t rows,
                 double step_size, const uint32_t key[2], int32_t L, const void* lp1_dev,
                 void* Q_out_dev, void* logacc_out_dev, void* pproj_out_dev /*nullable*/,
                 void* lp_out_dev, int flags, void* stream);

/* integrator-only entries (moves/hamiltonian/hmc_walk.py:61-107, hmc_side.py:80-135), used by the
 * reversibility / affine-invariance tests of the reference (tests/test_leapfrog.py:73-116,
 * tests/affine_invariance/test_leapfrog.py:26-100).  `centered_dev` is the centred complement
 * [n2, dim]; the walk momentum is given and returned in projected form S = p @ centered [rows,dim]
 * (the n x n momentum is never materialised).                                     */
HX_API int hx_leapfrog_walk(hx_ctx* ctx, int dtype, const hx_target* tgt, void* q_inout_dev,
                     void* S_inout_dev, const void* centered_dev, int32_t n2, int32_t rows,
                     double step_size, int32_t L, void* dK_out_dev /*nullable [rows]*/, void* stream);
HX_API int hx_leapfrog_side(hx_ctx* ctx, int dtype, const hx_target* tgt, void* q_inout_dev,
                     void* p_inout_dev /*[rows]*/, const void* diff_dev /*[rows,dim]*/, int32_t rows,
                     double step_size, int32_t L, void* stream);

/* Metropolis accept (samplers/sampler_utils.py:91-124): log(uniform(key,(n,),1e-10,1)) < logacc.
 * Updates cur/lp_cur in place for rows [row0,row0+rows) of a group of n; accepts_inout (int32
 * [rows]) is incremented.  The same key drives positions and log-probs, like the reference's two
 * calls (samplers/hamiltonian_ensemble.py:295-296).                                              */
HX_API int hx_accept(hx_ctx* ctx, int dtype, int impl, void* cur_inout_dev, const void* prop_dev,
              void* lp_cur_inout_dev, const void* lp_prop_dev, const void* logacc_dev,
              const uint32_t key[2], int32_t n, int32_t row0, int32_t rows, int32_t dim,
              int32_t* accepts_inout_dev, void* stream);

/* ---- (4) batch level: replaces the lax.scan of the MAIN phase -------------------------------
 * (samplers/hamiltonian_ensemble.py:284-318 driven by samplers/sampler_utils.py:29-88).
 * Runs T iterations: move(X1|X2) accept, move(X2|X1') accept, with the reference's main-phase
 * key order (keys[t,0] move1, keys[t,2] accept1, keys[t,1] move2, keys[t,3] accept2).
 *   keys_dev          uint32 [T,4,2] = split(key_main, 4T) (hx_split)
 *   X_inout_dev       [2n, dim] (group 1 rows then group 2 rows), lp_inout_dev [2n]
 *   chain_out_dev     [T, 2n, dim] or NULL; lp_out_dev [T, 2n] or NULL
 *   accepts_inout_dev int32 [2n]
 *   nonfinite_flag_dev int32[1], set to 1 if a stored log-prob is NaN/inf (sampler_utils.py:74-75)
 * move_kind: 0 walk, 1 side.
 * key_order: 0 = main phase (above); 1 = warm-up order keys[t,0] move1, keys[t,1] accept1,
 *            keys[t,2] move2, keys[t,3] accept2 (samplers/hamiltonian_ensemble.py:171,184,191,204),
 *            for warm-up runs whose adapter is the no-op one.                                    */
HX_API int hx_run_iterations(hx_ctx* ctx, int dtype, int impl, int move_kind, int key_order, const hx_target* tgt,
                      void* X_inout_dev, void* lp_inout_dev, int32_t n, double step_size, int32_t L,
                      const uint32_t* keys_dev, int64_t T, void* chain_out_dev, void* lp_out_dev,
                      int32_t* accepts_inout_dev, int32_t* nonfinite_flag_dev, void* stream);

/* Thinned chain emit for hx_run_iterations / hx_run_iterations_sharded: iteration t of a call is written
 * to chain_out / lp_out iff (t + phase) % every == 0, into consecutive slots (chain_out then holds
 * ceil((T + phase) / every) - ceil(phase / every) slices).  The reference stores every iteration and thins on read
 * (backend/backend.py:63-66), which is 262 MB per iteration at c5.  every = 1 restores the default.               */
HX_API int hx_ctx_set_chain_thinning(hx_ctx* ctx, int32_t every, int32_t phase);

/* ---- (4a) on-device chain statistics ------------------------------------------------------------
 * The reference estimates moments and the integrated autocorrelation time from the stored chain
 * (backend/backend.py:174-211 keeps every iteration, autocorr.py:15-34,44-118 FFT estimator).  These accumulators give the
 * same numbers without storing the chain: per iteration `hx_stats_feed` adds the ensemble X[nwalkers, dim] to
 *   count[1], sum[dim], sumsq[dim]          over ALL walkers (fixed-order partial sums: bit-reproducible)
 * and, for the tracked series s = j * n_dims + i  (walker row j * walker_stride, parameter dims_dev[i] or i), to
 *   lagsum[k, s] = sum_{t >= k} x_t x_{t-k}  (k < max_lag),  total[s] = sum_t x_t,
 *   first[t, s] = x_t (t < max_lag),  ring[t mod max_lag, s] = x_t (the last max_lag values)
 * from which sum_t (x_t - m)(x_{t+k} - m) of autocorr.py:30-31 follows exactly for k < max_lag (csrc/hx_stats.cu).
 * All pointers are DEVICE buffers owned and ZEROED by the caller; `ring`/`first` have the run dtype, the rest is f64;
 * partial_dev is scratch of 2 * hx_stats_row_blocks() * dim doubles.  hx_ctx_set_chain_stats(ctx, stats) makes the
 * single-GPU batch loops (hx_run_iterations, hx_run_vanilla_iterations) feed `stats` after EVERY iteration (thinning of the
 * chain emit does not apply); NULL switches it off.  The struct is copied.                                            */
typedef struct hx_chain_stats {
  int32_t nwalkers, dim;
  int32_t n_track, walker_stride, n_dims, max_lag;
  const int32_t* dims_dev;      /* int32 [n_dims] or NULL = parameters 0 .. n_dims-1 */
  int64_t* count_dev;
  double *sum_dev, *sumsq_dev, *partial_dev;
  double* lagsum_dev;           /* [max_lag, n_track * n_dims] */
  void *ring_dev, *first_dev;   /* [max_lag, n_track * n_dims], run dtype */
  double* total_dev;            /* [n_track * n_dims] */
} hx_chain_stats;
HX_API int32_t hx_stats_row_blocks(void);
HX_API int hx_stats_feed(hx_ctx* ctx, int dtype, const hx_chain_stats* stats, const void* X_dev, void* stream);
HX_API int hx_ctx_set_chain_stats(hx_ctx* ctx, const hx_chain_stats* stats);

/* ---- (4b) batch level: the WARM-UP scan with live adapters (device-resident adaptation) ---------------------
 * Replaces the warm-up scan body of samplers/hamiltonian_ensemble.py:165-217 for the dual-averaging step-size adapter
 * (adaptation/dual_averaging.py:16-96), the ChEES trajectory-length adapter (adaptation/chees.py:18-219) and their
 * composite (adaptation/adapter.py:44-78): per half-move (step, L) = adapter.value(state) is read by the move kernel from
 * device memory, adapter.update runs on the device after the move and before the accept, keys are used in the warm-up
 * order (keys[t,0] move1, [t,1] accept1, [t,2] move2, [t,3] accept2).  `state_inout` is a HOST struct (doubles hold the
 * run dtype's values exactly); the call synchronises the stream once at the end to return it.
 * Exact SIMT kernels: (step, L) never leave the device; ChEES needs dim <= 128 there (HX_E_UNSUPPORTED otherwise: the
 * caller keeps its host-driven loop).  Problems the precision policy sends to the tensor-core path (fp32 full-covariance
 * Gaussians, walk move): the launch sequence depends on (step, L), which are read back once per half-move (16 bytes, the
 * only host work); ChEES at any dim <= 1024: cov(group2)^-1 = (n/(n-1) M)^-1 by hx_spd_inverse on the Gram matrix the move
 * already forms, then one tensor-core GEMM with a row-dot epilogue for <cen_prop_i, cov^-1 pproj_i> of all walkers
 * (chees.py:57-73 solves an LU system with n right-hand sides).                                                      */
typedef struct hx_adapt_params {
  int32_t use_da, use_chees, agg_harmonic, pass_L;
  double da_target_accept, da_t0, da_gamma, da_kappa, da_init_step;
  double ch_T_min, ch_T_max, ch_T_interp, ch_jitter, ch_lr, ch_beta1, ch_beta2, ch_reg;
  double pass_step;
} hx_adapt_params;
typedef struct hx_adapt_state {
  int64_t da_iteration;
  double da_log_step, da_log_step_bar, da_H_bar;
  double ch_log_T, ch_log_T_bar, ch_m1, ch_m2, ch_iteration, ch_halton;
} hx_adapt_state;
/* inv_out[d, ldo] (fp32) = (scale * M)^-1 for a symmetric positive definite M[d, ldm] (fp32, lower triangle read): blocked
 * Cholesky, blocked inverse of the factor and X^T X in fp64 (csrc/hx_spd.cuh).  A non-positive pivot yields NaNs.         */
HX_API int hx_spd_inverse(hx_ctx* ctx, const float* M_dev, int32_t ldm, int32_t d, double scale, float* inv_out_dev,
                   int32_t ldo, void* stream);
HX_API int hx_run_warmup(hx_ctx* ctx, int dtype, int impl, int move_kind, const hx_target* tgt, void* X_inout_dev,
                  void* lp_inout_dev, int32_t n, const hx_adapt_params* params, hx_adapt_state* state_inout,
                  const uint32_t* keys_dev, int64_t T, void* chain_out_dev, void* lp_out_dev,
                  int32_t* accepts_inout_dev, int32_t* nonfinite_flag_dev, void* stream);

/* ---- (6) derivative-free ensemble moves and their scan --------------------------------------
 * kind 0 = stretch_move (moves/vanilla/stretch.py:5-76), 1 = side_move (moves/vanilla/side.py:5-56), 2 = walk_move
 * (moves/vanilla/walk.py:5-42).  hx_vanilla_move mirrors move(group1, group2, key, log_prob, log_prob_group1, stretch)
 * -> (proposed, log_accept_prob, proposed_log_prob); hx_run_vanilla_iterations replaces the scan of
 * samplers/ensemble.py:97-129 (keys_dev uint32 [T,2,2] = split(key, 2T); keys[t,h] drives move AND accept of half h).  */
HX_API int hx_vanilla_move(hx_ctx* ctx, int dtype, int impl, int kind, const hx_target* tgt, const void* X1_dev,
                    const void* X2_dev, int32_t n, int32_t row0, int32_t rows, double stretch, const uint32_t key[2],
                    const void* lp1_dev, void* Q_out_dev, void* logacc_out_dev, void* lp_out_dev, void* stream);
HX_API int hx_run_vanilla_iterations(hx_ctx* ctx, int dtype, int impl, int kind, const hx_target* tgt, void* X_inout_dev,
                              void* lp_inout_dev, int32_t n, double stretch, const uint32_t* keys_dev, int64_t T,
                              void* chain_out_dev, void* lp_out_dev, int32_t* accepts_inout_dev,
                              int32_t* nonfinite_flag_dev, void* stream);

/* ---- (5) row-sharded multi-GPU batch loop (the reference is single-device) ---------------------------
 * One process per GPU.  Rows of BOTH halves are split over `world` ranks (rank r owns rows [r n/world, (r+1) n/world) of each
 * half); every rank keeps a full copy [2n, dim] + log-probs [2n] in a library-owned region that its peers map with CUDA
 * IPC.  After a half-move the rank's updated rows are stored directly into every peer's copy over NVLink and an epoch
 * flag is published (csrc/hx_dist.cuh): this replaces the per-half-move NCCL all-gather, no staging copy, no collective
 * library on the data path.  RNG counters use global rows, so the chain equals the single-GPU chain bit for bit.
 *   hx_dist_init     allocate the region, return its 64-byte IPC handle (exchange the handles with any host-side
 *                    all-gather, e.g. torch.distributed)
 *   hx_dist_connect  map the peers' regions (`handles` = world x 64 bytes, indexed by rank)
 *   hx_dist_buffers  device pointers of the local full copy: load the initial ensemble / log-probs there (every rank the
 *                    full arrays), synchronise all ranks, then call hx_run_iterations_sharded; read the final state there
 *   hx_run_iterations_sharded  like hx_run_iterations for this rank's rows; chain_out [T, 2 n/world, dim] and lp_out
 *                    [T, 2 n/world] (rows of half 1, then of half 2), accepts int32 [2 n/world]; bit 1 of the
 *                    non-finite flag reports an exchange timeout (a peer stopped)                                   */
HX_API int hx_dist_init(hx_ctx* ctx, int dtype, int32_t rank, int32_t world, int32_t n, int32_t dim, void* handle_out /*64 B*/);
HX_API int hx_dist_connect(hx_ctx* ctx, const void* handles /*world x 64 B*/);
HX_API int hx_dist_buffers(hx_ctx* ctx, void** X_full_dev, void** lp_full_dev);
HX_API int hx_dist_destroy(hx_ctx* ctx);
HX_API int hx_run_iterations_sharded(hx_ctx* ctx, int dtype, int impl, int move_kind, int key_order, const hx_target* tgt,
                              double step_size, int32_t L, const uint32_t* keys_dev, int64_t T, void* chain_out_dev,
                              void* lp_out_dev, int32_t* accepts_inout_dev, int32_t* nonfinite_flag_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HX_H_ */
