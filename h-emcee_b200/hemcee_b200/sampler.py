"""`HamiltonianEnsembleSampler` with the reference's emcee-style API, driven by libhx kernels.

Mirrors reference src/hemcee/samplers/hamiltonian_ensemble.py:12-346 and the BaseSampler of
src/hemcee/samplers/sampler.py:12-102 (constructor, `run_mcmc`, `get_chain`, `get_logprob`,
`get_acceptance_prob`, `get_autocorr`, `diagnostics_warmup/main`, `adapter`, `backend`).

Differences a user switching over sees:
  * `log_prob` is a built-in target (`hemcee_b200.targets`) instead of an arbitrary JAX callable;
  * keys are uint32[2] arrays (`hemcee_b200.random.PRNGKey`); dtype / RNG layout follow
    `hemcee_b200.config` (`enable_x64`, `rng_impl`) like the JAX flags they stand for;
  * the chain is returned as a NumPy array on the host.

The main phase replaces the reference's `lax.scan` (hamiltonian_ensemble.py:284-318 through
sampler_utils.py:29-88) with `hx_run_iterations`: per half-move one fused kernel does momentum draw,
projection, leapfrog, energy and Metropolis accept, and writes the new state straight into the
on-device chain buffer, which is copied to pinned host memory per batch on a side stream.

With `group=` (a torch.distributed process group) the walkers of both halves are row-sharded over
the ranks.  Default `exchange="peer"`: every rank keeps a full copy of the ensemble that its peers map
through CUDA IPC; after a half-move the finish kernel stores the rank's updated rows straight into the
peers' copies over NVLink and publishes an epoch flag (csrc/hx_dist.cuh; no collective library on the data
path).  `exchange="nccl"` all-gathers the updated rows instead (used for warm-up with live adapters).

Also here, beyond the reference: `store="thinned" | "none"` (only the returned iterations / no chain leave
the device), `online_stats=` (on-device moments and lagged products: `sampler.stats.integrated_time()`
without a stored chain, hemcee_b200/stats.py), device-resident adaptation (`hx_run_warmup`).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, _rt, adaptation, autocorr
from . import random as hrandom
from .backend import Backend
from .moves import accept_proposal, hmc_side_move, hmc_walk_move, walk_prefetch
from .targets import as_target


_PIN_LIMIT = 16 << 30      # bytes of page-locked host memory one phase may hold for its chain


def calculate_batch_size(total_chains, dim, total_iterations, user_batch_size=None):
    """reference src/hemcee/samplers/sampler_utils.py:10-26."""
    if user_batch_size is not None:
        return user_batch_size
    return min(max(50, 10000 // (total_chains * dim)), total_iterations)


_NONFINITE_MSG = "Log probability returned NaN or inf value, please check your log probability function"


class BaseSampler:
    """reference src/hemcee/samplers/sampler.py:12-102."""

    def __init__(self, total_chains, dim, log_prob, move, backend=None):
        if total_chains < 4:
            raise ValueError("`total_chains` must be at least 4 to form meaningful ensemble groups")
        self.total_chains = int(total_chains)
        self.dim = int(dim)
        self.move = move
        self.backend = Backend() if backend is None else backend
        self.backend.reset(total_chains, dim)
        self.target = as_target(log_prob)
        if self.target.dim != self.dim:
            raise ValueError(f"target has dim {self.target.dim}, sampler dim is {self.dim}")
        # the reference builds jit(vmap(log_prob)) / jit(vmap(grad(log_prob))) here (sampler.py:52,55)
        self.log_prob = self.target.log_prob
        self.grad_log_prob = self.target.grad_log_prob
        self.diagnostics_warmup = None
        self.diagnostics_main = None

    def _validate_mcmc_inputs(self, num_samples, warmup, thin_by, initial_state):
        if thin_by < 1:
            raise ValueError("`thin_by` must be 1 or greater.")
        if warmup < 0:
            raise ValueError("`warmup` must be 0 or greater.")
        if num_samples < 0:
            raise ValueError("`num_samples` must be 0 or greater.")
        if tuple(initial_state.shape) != (self.total_chains, self.dim):
            raise ValueError("`inital_state` needs to have shape ")

    def get_chain(self, discard=0, thin=1, flat=False):
        return self.backend.get_chain(discard=discard, thin=thin, flat=flat)

    def get_logprob(self, discard=0, thin=1, flat=False):
        return self.backend.get_log_prob(discard=discard, thin=thin, flat=flat)

    def get_acceptance_prob(self):
        return self.backend.accepted / self.backend.iteration

    def get_autocorr(self, discard, thin):
        return autocorr.integrated_time(self.get_chain(discard=discard, thin=thin, flat=False))


class HamiltonianEnsembleSampler(BaseSampler):
    def __init__(self, total_chains, dim, log_prob, step_size=0.1, L=10, move=hmc_walk_move, backend=None,
                 adapt_inital_step_size=True, adapt_step_size=True, adapt_length=True, *, device=None, group=None,
                 chain_store="full", exchange="peer", store="full", online_stats=None, verbose=True):
        super().__init__(total_chains, dim, log_prob, move, backend)
        if getattr(move, "__name__", None) not in ("hmc_walk_move", "hmc_side_move"):
            raise ValueError("HamiltonianEnsembleSampler on B200 supports move=hmc_walk_move or hmc_side_move")
        if self.total_chains % 2:
            # the walk move needs equal halves (reference hmc_walk.py:36,91)
            raise ValueError("`total_chains` must be even")
        self.step_size = step_size
        self.L = L
        self.device = _rt.device(device)
        self.dtype = _rt.default_dtype()
        self.impl = _rt.config.rng_impl
        self.verbose = verbose
        self.group = group
        self.chain_store = chain_store
        if store not in ("full", "thinned", "none"):
            raise ValueError("store must be 'full' (every iteration, like the reference), 'thinned' (only the iterations "
                             "run_mcmc returns: no warm-up, every thin_by-th main iteration) or 'none' (no chain leaves the "
                             "device: results through online_stats / get_last_sample)")
        self.store = store
        # on-device chain statistics of the MAIN phase (hemcee_b200.stats.ChainStats: moments over all walkers, lagged products
        # of the tracked series -> integrated_time without a stored chain).  True = defaults, dict = ChainStats keyword arguments.
        if online_stats is True:
            online_stats = {}
        self.online_stats = online_stats if isinstance(online_stats, dict) else None
        self.stats = None
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' (NVLink peer-memory stores from the kernels) or 'nccl' (all-gather)")
        self.exchange = exchange
        if group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            if (self.total_chains // 2) % self.world:
                raise ValueError("total_chains/2 must be divisible by the number of ranks")
            if chain_store == "local":      # this rank stores only the chain of its own rows (half 1 rows, then half 2 rows)
                self.backend.reset(2 * ((self.total_chains // 2) // self.world), self.dim)
        else:
            self.world, self.rank = 1, 0
        self.adapt_inital_step_size = adapt_inital_step_size
        self.adapter = adaptation.select_adapter(adapt_step_size, adapt_length, self.move, self.step_size, self.L,
                                                 self.dtype, self._allreduce)
        self.adapter_state = self.adapter.init(self.dim)
        self.kernel_launch_count = 0

    # -- plumbing ---------------------------------------------------------------------------------
    def _say(self, msg):
        if self.verbose and self.rank == 0:
            print(msg)

    def _allreduce(self, t):
        if self.group is not None and self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, group=self.group)
        return t

    def _allgather(self, full, local):
        import torch.distributed as dist
        dist.all_gather_into_tensor(full, local, group=self.group)

    def _upload_sharded(self, x0_host, n):
        """Host ensemble -> every rank's device copy: each rank uploads only ITS rows of both halves over PCIe and the ranks
        all-gather them over NVLink (one collective per run_mcmc call, off the hot path) instead of every rank pulling the whole
        ensemble through the host bridge at once (262 MB per rank at c5: 8 ranks -> 2.1 GB, ~30 ms of a 10-iteration run)."""
        rows = n // self.world
        r0 = self.rank * rows
        d = x0_host.shape[1]
        loc = torch.empty((2, rows, d), dtype=self.dtype, device=self.device)
        loc[0].copy_(x0_host[r0:r0 + rows], non_blocking=True)
        loc[1].copy_(x0_host[n + r0:n + r0 + rows], non_blocking=True)
        full = torch.empty((self.world * 2, rows, d), dtype=self.dtype, device=self.device)      # rank-major concatenation
        self._allgather(full, loc)
        return full.view(self.world, 2, rows, d).permute(1, 0, 2, 3).reshape(2 * n, d).contiguous()

    def _log_prob_sharded(self, Xf, n):
        """log-probs of the full ensemble, each rank evaluating its own rows of both halves (all-gathered)."""
        rows = n // self.world
        r0 = self.rank * rows
        loc = torch.stack([self.log_prob(Xf[r0:r0 + rows]), self.log_prob(Xf[n + r0:n + r0 + rows])])      # [2, rows]
        full = torch.empty((self.world * 2, rows), dtype=loc.dtype, device=loc.device)
        self._allgather(full, loc.contiguous())
        return full.view(self.world, 2, rows).permute(1, 0, 2).reshape(2 * n)

    def _device_adapt_ok(self):
        """Device-resident adaptation (hx_run_warmup): adapter state and update kernels on the device.  The exact SIMT kernels
        read (step, L) from device memory (no host round trip; ChEES up to dim 128); problems the precision policy sends to the
        tensor-core path (fp32 full-covariance Gaussians) read the two words back once per half-move and take ChEES at any dim
        (cov(group2)^-1 by an fp64 blocked Cholesky + one tensor-core GEMM).  Everything else keeps the host loop."""
        if adaptation.device_params(self.adapter) is None or not _rt.config.device_adaptation:
            return False
        from .targets import GaussianFull
        n = self.total_chains // 2
        tc = (self._move_kind == 0 and _rt.config.precision != "exact" and self.dtype == torch.float32
              and isinstance(self.target, GaussianFull)
              and ((self.dim >= 128 and n >= 2048) if _rt.config.precision == "auto" else (self.dim >= 16 and n >= 64)))
        ch_on = isinstance(self.adapter, (adaptation.ChEESAdapter, adaptation.CompositeAdapter))
        if ch_on and self.dim > 128 and not tc:
            return False
        return True

    @property
    def _peer_ok(self):
        """Peer-memory exchange needs a rank's rows and log-probs to be multiples of 16 bytes (vector stores)."""
        rows = (self.total_chains // 2) // max(self.world, 1)
        esz = 8 if self.dtype == torch.float64 else 4
        return self.exchange == "peer" and (rows * self.dim * esz) % 16 == 0 and (rows * esz) % 16 == 0

    @property
    def _thinned(self):
        return self.store in ("thinned", "none") and self.world == 1

    @property
    def _move_kind(self):
        return 0 if self.move.__name__ == "hmc_walk_move" else 1

    # -- reference hamiltonian_ensemble.py:64-139 ----------------------------------------------------
    def run_mcmc(self, key, initial_state, num_samples, warmup=1000, thin_by=1, batch_size=None, show_progress=False):
        keys = hrandom.split(key, 3, self.device, self.impl)                        # :90
        self._validate_mcmc_inputs(num_samples, warmup, thin_by, initial_state)     # :95
        x0 = torch.as_tensor(np.asarray(initial_state) if not isinstance(initial_state, torch.Tensor) else initial_state)
        n = self.total_chains // 2
        if self.world > 1 and x0.device.type == "cpu" and n % self.world == 0:
            x0 = self._upload_sharded(x0, n)
        else:
            x0 = x0.to(device=self.device, dtype=self.dtype, non_blocking=True).contiguous()
        self._say(f"Using {self.total_chains} total chains: Group 1 ({n}), Group 2 ({self.total_chains - n})")
        group1, group2 = x0[:n].clone(), x0[n:].clone()                             # :107-108
        warmup_bs = calculate_batch_size(self.total_chains, self.dim, warmup, batch_size)
        main_bs = calculate_batch_size(self.total_chains, self.dim, num_samples * thin_by, batch_size)

        import time as _time
        def _now():
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            return _time.perf_counter()
        t_start = _now()
        self.timings = {}                                                           # wall time of the phases [s] (diagnostics)
        if self.adapt_inital_step_size:                                             # :115-118
            self.step_size = adaptation.find_reasonable_step_size(
                keys[0], group1, group2, self.target, self.target, self.step_size, self.move, self.impl)
            self._say(f"Using inital step size of {float(self.step_size)!r}")

        self.timings["initial_step_size_s"] = _now() - t_start
        t_w = _now()
        if warmup > 0:                                                              # :121-129
            self._say("Starting warmup...")
            group1, group2 = self._mcmc_warmup(keys[1], group1, group2, warmup, self.adapter, warmup_bs, show_progress)
            self._say("Warmup complete.")
            step_size, L = self.adapter.finalize(self.adapter_state)
            self._say(f"Found step size {float(step_size)!r} and integration length: {int(L)}")
        else:
            step_size, L = self.step_size, self.L                                   # :131-132
        self.found = (step_size, L)
        self.timings["warmup_s"] = _now() - t_w

        self._say("Starting main sampling...")
        t_m = _now()
        self._mcmc_main(keys[2], group1, group2, num_samples, thin_by, step_size, L, main_bs, show_progress,
                        warmup_offset=warmup)
        self.timings["main_s"] = _now() - t_m
        self._say("Main sampling complete.")
        if self._thinned:
            # only the returned iterations were stored: the backend keeps counting the iterations that were RUN and maps
            # discard / thin requests onto the stored rows (Backend.set_stored_layout)
            self.backend.iteration = warmup + num_samples * thin_by
            self.backend.iteration_main = num_samples * thin_by
            last = None
            if getattr(self, "last_state", None) is not None:
                last = (np.asarray(self.last_state.detach().cpu()), np.asarray(self.last_log_prob.detach().cpu())
                        if getattr(self, "last_log_prob", None) is not None else None)
            self.backend.set_stored_layout(warmup, thin_by, last)
        return self.get_chain(discard=warmup, thin=thin_by)                         # :139

    def _save_prefix(self, host_chain, host_lp, host_flag, starts, offset, copy_stream):
        """A stored log-prob became non-finite in some batch: like the reference (sampler_utils.py:74-82 checks each batch
        BEFORE saving it, so every earlier batch is already in the backend), keep the batches that completed before the
        offending one.  Their accept counters are not separable from the device total and are left out."""
        copy_stream.synchronize()
        bad = next((b for b in range(len(starts)) if int(host_flag[b]) != 0), len(starts))
        rows = starts[bad] if bad < len(starts) else 0
        if rows > 0:
            self.backend.save_slice(host_chain[:rows].numpy().copy(), host_lp[:rows].numpy().copy(),
                                    np.zeros(self.backend.nwalkers), offset)

    # -- batch-level engine (single GPU): T iterations per libhx call ----------------------------------
    def _run_batches(self, key, X, lp, T, step_size, L, batch_size, offset, key_order, show_progress, adapt=None,
                     thin_store=1):
        """Runs T iterations in batches; streams (chain, log-prob) batches to pinned host memory on a
        side stream while the next batch computes.  Returns per-walker accept counts (numpy).
        `adapt` = (hx_adapt_params, hx_adapt_state): warm-up with the adapter resident on the device (hx_run_warmup).
        `thin_store` = k > 1: only iterations 0, k, 2k, ... of this phase are emitted and stored (hx_ctx_set_chain_thinning);
        0: nothing is stored."""
        dev, W, d = self.device, self.total_chains, self.dim
        n = W // 2
        lib = _lib.load()
        accepts = torch.zeros(W, dtype=torch.int32, device=dev)
        if T == 0:
            return accepts.cpu().numpy()
        keys_dev = hrandom.split_device(key, 4 * T, dev, self.impl)                 # :227-228 / :328-330
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        nb = (T + batch_size - 1) // batch_size
        k = int(thin_store)
        nstore = lambda a, b: (b - a) if k == 1 else (0 if k == 0 else (-(-b // k)) - (-(-a // k)))   # stored iterations in [a, b)
        Ts = nstore(0, T)
        # the stored chain of the phase.  Page-locked (asynchronous D2H at full PCIe rate) up to _PIN_LIMIT bytes; beyond that
        # pageable memory, like the reference's host arrays (cudaHostAlloc of tens of GB is slow, counts against memlock /
        # cgroup limits and cannot be swapped)
        pin = Ts > 0 and Ts * W * d * torch.empty((), dtype=self.dtype).element_size() <= _PIN_LIMIT
        host_chain = torch.empty((Ts, W, d), dtype=self.dtype, pin_memory=pin)
        host_lp = torch.empty((Ts, W), dtype=self.dtype, pin_memory=pin)
        host_flag = torch.zeros(nb, dtype=torch.int32, pin_memory=True)
        bmax = max(nstore(b * batch_size, min((b + 1) * batch_size, T)) for b in range(nb))
        bufs = [(torch.empty((max(bmax, 1), W, d), dtype=self.dtype, device=dev),
                 torch.empty((max(bmax, 1), W), dtype=self.dtype, device=dev)) for _ in range(min(2, nb))]
        copy_stream = torch.cuda.Stream(dev)
        done = [None, None]
        starts = []
        main = torch.cuda.current_stream(dev)
        desc = self.target.descriptor(self.dtype, dev)
        it = range(nb)
        if show_progress:
            from tqdm import tqdm
            it = tqdm(it)
        pos = 0
        try:
            for b in it:
                start, stop = b * batch_size, min((b + 1) * batch_size, T)
                cbuf, lbuf = bufs[b % 2]
                ns = nstore(start, stop)
                if done[b % 2] is not None:
                    main.wait_event(done[b % 2])          # the copy out of this buffer must be finished
                if k > 1:
                    _lib.check(lib.hx_ctx_set_chain_thinning(_rt.ctx(dev), k, start % k), "hx_ctx_set_chain_thinning")
                cptr, lptr = (_rt.ptr(cbuf), _rt.ptr(lbuf)) if k != 0 else (None, None)
                if adapt is None:
                    st = lib.hx_run_iterations(_rt.ctx(dev), _rt.dtype_code(self.dtype), _rt.impl_code(self.impl),
                                               self._move_kind, key_order, desc, _rt.ptr(X), _rt.ptr(lp), n, float(step_size),
                                               int(L), _rt.ptr(keys_dev[4 * start:]), stop - start, cptr, lptr,
                                               _rt.ptr(accepts), _rt.ptr(flag), _rt.stream(dev))
                    _lib.check(st, "hx_run_iterations")
                else:
                    import ctypes as C
                    st = lib.hx_run_warmup(_rt.ctx(dev), _rt.dtype_code(self.dtype), _rt.impl_code(self.impl), self._move_kind, desc,
                                           _rt.ptr(X), _rt.ptr(lp), n, C.byref(adapt[0]), C.byref(adapt[1]),
                                           _rt.ptr(keys_dev[4 * start:]), stop - start, cptr, lptr,
                                           _rt.ptr(accepts), _rt.ptr(flag), _rt.stream(dev))
                    _lib.check(st, "hx_run_warmup")
                ev = torch.cuda.Event()
                ev.record(main)
                copy_stream.wait_event(ev)
                with torch.cuda.stream(copy_stream):
                    if ns > 0:
                        host_chain[pos:pos + ns].copy_(cbuf[:ns], non_blocking=True)
                        host_lp[pos:pos + ns].copy_(lbuf[:ns], non_blocking=True)
                    host_flag[b:b + 1].copy_(flag, non_blocking=True)
                    done[b % 2] = torch.cuda.Event()
                    done[b % 2].record(copy_stream)
                starts.append(pos)
                pos += ns
                if b >= 1 and int(host_flag[b - 1]) != 0 and done[(b - 1) % 2].query():
                    self._save_prefix(host_chain, host_lp, host_flag, starts, offset, copy_stream)
                    raise ValueError(_NONFINITE_MSG)                                     # sampler_utils.py:74-75
        finally:
            if k > 1:
                lib.hx_ctx_set_chain_thinning(_rt.ctx(dev), 1, 0)
        copy_stream.synchronize()
        main.synchronize()
        if int(host_flag.max()) != 0:
            self._save_prefix(host_chain, host_lp, host_flag, starts, offset, copy_stream)
            raise ValueError(_NONFINITE_MSG)
        acc = accepts.cpu().numpy()
        # one slice for the whole phase: the pinned host buffer becomes the stored chain without another copy
        if Ts > 0:
            self.backend.save_slice(host_chain.numpy(), host_lp.numpy(), acc, offset)
        else:
            self.backend.accepted = self.backend.accepted + acc
        self.kernel_launch_count += T * 2 * (6 if self._move_kind == 0 else 2)
        return acc

    # -- batch-level engine, row-sharded over the ranks of `group`: exchange over NVLink peer memory -------------
    def _run_batches_sharded(self, key, g1, g2, T, step_size, L, batch_size, offset, key_order, show_progress):
        """hx_run_iterations_sharded (include/hx.h section 5): every rank updates its rows of both halves; the updated rows
        are stored into the peers' copies by the kernels themselves.  Chain slices of the LOCAL rows stream to pinned host
        memory; with chain_store='full' they are all-gathered per batch (NCCL, off the hot path) before the host copy."""
        import torch.distributed as dist
        dev, W, d = self.device, self.total_chains, self.dim
        n = W // 2
        rows = n // self.world
        lib = _lib.load()
        keys_dev = hrandom.split_device(key, 4 * T, dev, self.impl)
        Xf, lpf = _rt.dist_setup(dev, self.group, self.dtype, n, d)
        Xf[:n].copy_(g1); Xf[n:].copy_(g2)
        lpf.copy_(self._log_prob_sharded(Xf, n))
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)            # nobody stores rows into a peer before that peer loaded its initial state
        accepts = torch.zeros(2 * rows, dtype=torch.int32, device=dev)
        full = self.chain_store == "full"
        Wst = W if full else 2 * rows
        if T == 0:
            return np.zeros(Wst)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        nb = (T + batch_size - 1) // batch_size
        B = min(batch_size, T)
        host_chain = torch.empty((T, Wst, d), dtype=self.dtype, pin_memory=True)
        host_lp = torch.empty((T, Wst), dtype=self.dtype, pin_memory=True)
        bufs = [(torch.empty((B, 2, rows, d), dtype=self.dtype, device=dev),
                 torch.empty((B, 2, rows), dtype=self.dtype, device=dev)) for _ in range(min(2, nb))]
        copy_stream = torch.cuda.Stream(dev)
        done = [None, None]
        main = torch.cuda.current_stream(dev)
        desc = self.target.descriptor(self.dtype, dev)
        it = range(nb)
        if show_progress and self.rank == 0:
            from tqdm import tqdm
            it = tqdm(it)
        for b in it:
            start, stop = b * batch_size, min((b + 1) * batch_size, T)
            nbt = stop - start
            cbuf, lbuf = bufs[b % 2]
            if done[b % 2] is not None:
                main.wait_event(done[b % 2])
            st = lib.hx_run_iterations_sharded(_rt.ctx(dev), _rt.dtype_code(self.dtype), _rt.impl_code(self.impl), self._move_kind,
                                               key_order, desc, float(step_size), int(L), _rt.ptr(keys_dev[4 * start:]), nbt,
                                               _rt.ptr(cbuf), _rt.ptr(lbuf), _rt.ptr(accepts), _rt.ptr(flag), _rt.stream(dev))
            _lib.check(st, "hx_run_iterations_sharded")
            if full:
                gc = torch.empty((self.world, nbt, 2, rows, d), dtype=self.dtype, device=dev)
                gl = torch.empty((self.world, nbt, 2, rows), dtype=self.dtype, device=dev)
                dist.all_gather_into_tensor(gc, cbuf[:nbt].contiguous(), group=self.group)
                dist.all_gather_into_tensor(gl, lbuf[:nbt].contiguous(), group=self.group)
                src_c = gc.permute(1, 2, 0, 3, 4).reshape(nbt, W, d)
                src_l = gl.permute(1, 2, 0, 3).reshape(nbt, W)
            else:
                src_c, src_l = cbuf[:nbt].reshape(nbt, 2 * rows, d), lbuf[:nbt].reshape(nbt, 2 * rows)
            ev = torch.cuda.Event()
            ev.record(main)
            copy_stream.wait_event(ev)
            with torch.cuda.stream(copy_stream):
                host_chain[start:stop].copy_(src_c, non_blocking=True)
                host_lp[start:stop].copy_(src_l, non_blocking=True)
                src_c.record_stream(copy_stream); src_l.record_stream(copy_stream)
                done[b % 2] = torch.cuda.Event()
                done[b % 2].record(copy_stream)
        copy_stream.synchronize()
        main.synchronize()
        self._allreduce_max(flag)
        code = int(flag.item())
        if code & 2:
            raise _lib.HxError("row exchange timed out: a peer rank stopped")
        if code & 1:
            raise ValueError(_NONFINITE_MSG)                                         # sampler_utils.py:74-75
        if full:
            ga = torch.empty((self.world, 2, rows), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(ga, accepts.reshape(2, rows).contiguous(), group=self.group)
            acc = ga.permute(1, 0, 2).reshape(W).cpu().numpy()
        else:
            acc = accepts.cpu().numpy()
        self.backend.save_slice(host_chain.numpy(), host_lp.numpy(), acc, offset)
        self.kernel_launch_count += T * 2 * (8 if self._move_kind == 0 else 4)
        g1.copy_(Xf[:n]); g2.copy_(Xf[n:])
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)
        return acc

    def _allreduce_max(self, t):
        if self.group is not None and self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    # -- per-half-move engine (warm-up with adapters; row-sharded multi-GPU) ------------------------------
    def _half(self, Xa_loc, Xb_full, lpa_loc, step, L, k_move, k_acc, acc_loc, n, row0, adapter_state):
        prop, la, pproj, lpp = self.move(Xa_loc, Xb_full, step, k_move, self.target, self.target, int(L), lpa_loc,
                                         row0=row0, n=n, impl=self.impl)
        if adapter_state is not None:                                               # :176-182 (before the accept)
            kw = dict(state=adapter_state, log_accept_rate=la, position_current=Xa_loc, position_proposed=prop,
                      momentum_proposed=pproj, group2=Xb_full)
            if isinstance(self.adapter, (adaptation.ChEESAdapter, adaptation.CompositeAdapter)):
                kw["n_total"] = n
            adapter_state = self.adapter.update(**kw)
        accept_proposal(Xa_loc, prop, lpa_loc, lpp, la, k_acc, acc_loc, row0=row0, n=n, impl=self.impl)
        return adapter_state

    def _run_steps(self, key, g1, g2, T, batch_size, offset, key_order, fixed, adapter_state, show_progress):
        """Python-driven loop: one libhx move + accept per half-move.  Used for warm-up with a live
        adapter and for the row-sharded multi-GPU path."""
        dev, W, d = self.device, self.total_chains, self.dim
        n = W // 2
        rows = n // self.world
        row0 = self.rank * rows
        keys = hrandom.split(key, 4 * T, dev, self.impl).reshape(T, 4, 2) if T > 0 else np.zeros((0, 4, 2), np.uint32)
        sharded = self.world > 1
        if sharded:
            X1f, X2f = g1.contiguous(), g2.contiguous()
            X1, X2 = X1f[row0:row0 + rows].clone(), X2f[row0:row0 + rows].clone()
        else:
            X1, X2 = g1, g2
            X1f, X2f = X1, X2
        lp1, lp2 = self.log_prob(X1), self.log_prob(X2)                              # :224-225 / :325-326
        if sharded:
            lp1f = torch.empty(n, dtype=self.dtype, device=dev)
            lp2f = torch.empty(n, dtype=self.dtype, device=dev)
        acc1 = torch.zeros(rows, dtype=torch.int32, device=dev)
        acc2 = torch.zeros(rows, dtype=torch.int32, device=dev)
        order = (0, 2, 1, 3) if key_order == 0 else (0, 1, 2, 3)     # move1, accept1, move2, accept2
        full_store = (not sharded) or self.chain_store == "full"
        Wst = W if full_store else 2 * rows
        total_acc = np.zeros(Wst)
        it = range(0, T, max(batch_size, 1))
        if show_progress:
            from tqdm import tqdm
            it = tqdm(it)
        for start in it:
            stop = min(start + batch_size, T)
            cbuf = torch.empty((stop - start, Wst, d), dtype=self.dtype, device=dev)
            lbuf = torch.empty((stop - start, Wst), dtype=self.dtype, device=dev)
            acc1.zero_(); acc2.zero_()
            for t in range(start, stop):
                k = keys[t]
                if fixed is None:
                    step, L = self.adapter.value(adapter_state)                      # :169
                else:
                    step, L = fixed
                if self._move_kind == 0:          # the draw of the following half-move only needs its key
                    walk_prefetch(k[order[2]], n, row0, rows, dev, self.impl)
                adapter_state = self._half(X1, X2f, lp1, step, L, k[order[0]], k[order[1]], acc1, n, row0, adapter_state)
                if sharded:
                    self._allgather(X1f, X1)
                if fixed is None:
                    step, L = self.adapter.value(adapter_state)                      # :189
                if self._move_kind == 0 and t + 1 < T:
                    walk_prefetch(keys[t + 1][order[0]], n, row0, rows, dev, self.impl)
                adapter_state = self._half(X2, X1f, lp2, step, L, k[order[2]], k[order[3]], acc2, n, row0, adapter_state)
                if sharded:
                    self._allgather(X2f, X2)
                i = t - start
                if full_store:
                    cbuf[i, :n].copy_(X1f); cbuf[i, n:].copy_(X2f)
                    if sharded:
                        self._allgather(lp1f, lp1); self._allgather(lp2f, lp2)
                        lbuf[i, :n].copy_(lp1f); lbuf[i, n:].copy_(lp2f)
                    else:
                        lbuf[i, :n].copy_(lp1); lbuf[i, n:].copy_(lp2)
                else:
                    cbuf[i, :rows].copy_(X1); cbuf[i, rows:].copy_(X2)
                    lbuf[i, :rows].copy_(lp1); lbuf[i, rows:].copy_(lp2)
            bad = ~torch.isfinite(lbuf).all()
            if sharded:
                badt = bad.to(torch.int32).reshape(1)
                self._allreduce(badt)
                bad = badt[0] > 0
            if bool(bad):                                                            # sampler_utils.py:74-75
                raise ValueError(_NONFINITE_MSG)
            if full_store and sharded:
                a1f = torch.empty(n, dtype=torch.int32, device=dev); a2f = torch.empty(n, dtype=torch.int32, device=dev)
                self._allgather(a1f, acc1); self._allgather(a2f, acc2)
                acc = torch.cat([a1f, a2f]).cpu().numpy()
            else:
                acc = torch.cat([acc1, acc2]).cpu().numpy()
            total_acc += acc
            self.backend.save_slice(cbuf.cpu().numpy(), lbuf.cpu().numpy(), acc, offset + start)
        self.kernel_launch_count += T * 2 * (8 if self._move_kind == 0 else 4)
        if sharded:
            g1.copy_(X1f); g2.copy_(X2f)
        return total_acc, adapter_state

    # -- reference hamiltonian_ensemble.py:141-248 --------------------------------------------------------
    def _mcmc_warmup(self, key, group1, group2, warmup, adapter, batch_size, show_progress):
        noop = isinstance(adapter, adaptation.NoOpAdapter)
        if noop and self.world == 1:
            step, L = adapter.value(self.adapter_state)
            X = torch.cat([group1, group2]).contiguous()
            lp = self.log_prob(X)
            acc = self._run_batches(key, X, lp, warmup, step, L, batch_size, 0, 1, show_progress,
                                    thin_store=0 if self._thinned else 1)
            group1, group2 = X[: self.total_chains // 2].clone(), X[self.total_chains // 2:].clone()
        elif not noop and self.world == 1 and self._device_adapt_ok():
            # adapter resident on the device: (step, L) read by the move kernels, update kernels between move and accept
            params = adaptation.device_params(adapter)
            dstate = adaptation.state_to_device(adapter, self.adapter_state)
            X = torch.cat([group1, group2]).contiguous()
            lp = self.log_prob(X)
            acc = self._run_batches(key, X, lp, warmup, None, None, batch_size, 0, 1, show_progress, adapt=(params, dstate),
                                    thin_store=0 if self._thinned else 1)
            self.adapter_state = adaptation.state_from_device(adapter, self.adapter_state, dstate)
            group1, group2 = X[: self.total_chains // 2].clone(), X[self.total_chains // 2:].clone()
        elif noop and self._peer_ok:
            step, L = adapter.value(self.adapter_state)
            acc = self._run_batches_sharded(key, group1, group2, warmup, step, L, batch_size, 0, 1, show_progress)
        else:
            fixed = adapter.value(self.adapter_state) if noop else None
            acc, st = self._run_steps(key, group1, group2, warmup, batch_size, 0, 1, fixed,
                                      None if noop else self.adapter_state, show_progress)
            if not noop:
                self.adapter_state = st
        self.diagnostics_warmup = {"accepts": acc, "acceptance_rate": acc / warmup}
        self.backend.mark_warmup_end()
        return group1, group2

    # -- reference hamiltonian_ensemble.py:250-343 --------------------------------------------------------
    def _mcmc_main(self, key, group1, group2, num_samples, thin_by, step_size, L, batch_size, show_progress,
                   warmup_offset=0):
        total = num_samples * thin_by
        if self.world == 1:
            X = torch.cat([group1, group2]).contiguous()
            lp = self.log_prob(X)                                                    # :325-326
            if self.online_stats is not None:
                from .stats import ChainStats
                cfg = dict(max_lag=128, walkers=min(self.total_chains, 1024), dims=None if self.dim <= 64 else list(range(16)))
                cfg.update(self.online_stats)
                self.stats = ChainStats(self.total_chains, self.dim, self.dtype, self.device, **cfg)
                self.backend.stats = self.stats
                self.stats.attach()
            try:
                thin_store = 0 if self.store == "none" else (thin_by if self._thinned else 1)
                acc = self._run_batches(key, X, lp, total, step_size, L, batch_size, 0 if self._thinned else warmup_offset, 0,
                                        show_progress, thin_store=thin_store)
            finally:
                if self.stats is not None:
                    self.stats.detach()
            self.last_state, self.last_log_prob = X, lp
        elif self._peer_ok:
            acc = self._run_batches_sharded(key, group1, group2, total, step_size, L, batch_size, warmup_offset, 0, show_progress)
            self.last_state = torch.cat([group1, group2])
        else:
            acc, _ = self._run_steps(key, group1, group2, total, batch_size, warmup_offset, 0, (step_size, L), None,
                                     show_progress)
            self.last_state = torch.cat([group1, group2])
        self.diagnostics_main = {"accepts": acc, "acceptance_rate": acc / total if total else acc}
