"""Derivative-free affine-invariant ensemble sampler and its moves, executed by libhx.

  stretch_move, side_move, walk_move   reference src/hemcee/moves/vanilla/{stretch,side,walk}.py
  EnsembleSampler                      reference src/hemcee/samplers/ensemble.py:9-140

Move protocol of the reference:
    move(group1, group2, key, log_prob, log_prob_group1, stretch=2.0) -> (proposed, log_accept_prob, proposed_log_prob)
`log_prob` is a built-in target (`hemcee_b200.targets`).  The sampler keeps the reference's key schedule
(`split(key, 2 * total)`, the move's key also drives its accept, ensemble.py:79-109), stores every iteration (warm-up
included) and returns `get_chain(discard=warmup, thin=thin_by)`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, _rt
from . import random as hrandom
from .sampler import BaseSampler, calculate_batch_size, _NONFINITE_MSG
from .targets import as_target

_KINDS = {"stretch_move": 0, "side_move": 1, "walk_move": 2}


def _vanilla(kind, group1, group2, key, log_prob, log_prob_group1, stretch, row0, n, impl):
    tgt = as_target(log_prob)
    dev = _rt.device(group1.device)
    g1, g2, lp1 = group1.contiguous(), group2.contiguous(), log_prob_group1.contiguous()
    rows, d = g1.shape
    n = g2.shape[0] if n is None else int(n)
    if g2.shape != (n, d):
        raise ValueError(f"complement must have shape ({n}, {d}), got {tuple(g2.shape)}")
    q = torch.empty_like(g1)
    la = torch.empty(rows, dtype=g1.dtype, device=dev)
    lp = torch.empty(rows, dtype=g1.dtype, device=dev)
    st = _lib.load().hx_vanilla_move(_rt.ctx(dev), _rt.dtype_code(g1.dtype), _rt.impl_code(impl), kind, tgt.descriptor(g1.dtype, dev),
                                     _rt.ptr(g1), _rt.ptr(g2), n, int(row0), rows, float(stretch), _lib.key_arg(_rt.as_key(key)),
                                     _rt.ptr(lp1), _rt.ptr(q), _rt.ptr(la), _rt.ptr(lp), _rt.stream(dev))
    _lib.check(st, "hx_vanilla_move")
    return q, la, lp


def stretch_move(group1, group2, key, log_prob, log_prob_group1, stretch=2.0, *, row0=0, n=None, impl=None):
    """Goodman-Weare stretch move (Alg. 1 of arXiv:2505.02987); reference moves/vanilla/stretch.py:5-56."""
    return _vanilla(0, group1, group2, key, log_prob, log_prob_group1, stretch, row0, n, impl)


def side_move(group1, group2, key, log_prob, log_prob_group1, stretch=2.0, *, row0=0, n=None, impl=None):
    """Derivative-free side move (Alg. 2); reference moves/vanilla/side.py:5-56."""
    return _vanilla(1, group1, group2, key, log_prob, log_prob_group1, stretch, row0, n, impl)


def walk_move(group1, group2, key, log_prob, log_prob_group1, *, row0=0, n=None, impl=None, **_):
    """Derivative-free walk move (Alg. 4); reference moves/vanilla/walk.py:5-42."""
    return _vanilla(2, group1, group2, key, log_prob, log_prob_group1, 2.0, row0, n, impl)


class EnsembleSampler(BaseSampler):
    """reference src/hemcee/samplers/ensemble.py:9-140."""

    def __init__(self, total_chains, dim, log_prob, move=stretch_move, backend=None, *, device=None, verbose=True):
        super().__init__(total_chains, dim, log_prob, move, backend)
        if getattr(move, "__name__", None) not in _KINDS:
            raise ValueError("EnsembleSampler on B200 supports move=stretch_move, side_move or walk_move")
        if self.total_chains % 2:
            raise ValueError("`total_chains` must be even")
        self.device = _rt.device(device)
        self.dtype = _rt.default_dtype()
        self.impl = _rt.config.rng_impl
        self.verbose = verbose

    def run_mcmc(self, key, initial_state, num_samples, warmup=1000, thin_by=1, batch_size=None, show_progress=False,
                 **kwargs):
        self._validate_mcmc_inputs(num_samples, warmup, thin_by, initial_state)     # ensemble.py:58
        stretch = float(kwargs.pop("stretch", 2.0))
        if kwargs:
            raise TypeError(f"unexpected move arguments: {sorted(kwargs)}")
        total = warmup + num_samples * thin_by                                      # :66
        dev, W, d = self.device, self.total_chains, self.dim
        n = W // 2
        if self.verbose:
            print(f"Using {W} total chains: Group 1 ({n}), Group 2 ({W - n})")
        bs = calculate_batch_size(W, d, total, batch_size) if total > 0 else 1
        x0 = torch.as_tensor(np.asarray(initial_state) if not isinstance(initial_state, torch.Tensor) else initial_state)
        X = x0.to(device=dev, dtype=self.dtype).contiguous().clone()
        lp = self.log_prob(X)                                                       # :93-94
        lib = _lib.load()
        keys_dev = hrandom.split_device(key, 2 * total, dev, self.impl)            # :79-81
        accepts = torch.zeros(W, dtype=torch.int32, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        desc = self.target.descriptor(self.dtype, dev)
        kind = _KINDS[self.move.__name__]
        host_chain = torch.empty((total, W, d), dtype=self.dtype, pin_memory=total > 0)
        host_lp = torch.empty((total, W), dtype=self.dtype, pin_memory=total > 0)
        nb = (total + bs - 1) // bs if total else 0
        bufs = [(torch.empty((min(bs, total), W, d), dtype=self.dtype, device=dev),
                 torch.empty((min(bs, total), W), dtype=self.dtype, device=dev)) for _ in range(min(2, nb))]
        copy_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        done = [None, None]
        it = range(nb)
        if show_progress:
            from tqdm import tqdm
            it = tqdm(it)
        for b in it:
            start, stop = b * bs, min((b + 1) * bs, total)
            cbuf, lbuf = bufs[b % 2]
            if done[b % 2] is not None:
                main.wait_event(done[b % 2])
            st = lib.hx_run_vanilla_iterations(_rt.ctx(dev), _rt.dtype_code(self.dtype), _rt.impl_code(self.impl), kind, desc,
                                               _rt.ptr(X), _rt.ptr(lp), n, stretch, _rt.ptr(keys_dev[2 * start:]), stop - start,
                                               _rt.ptr(cbuf), _rt.ptr(lbuf), _rt.ptr(accepts), _rt.ptr(flag), _rt.stream(dev))
            _lib.check(st, "hx_run_vanilla_iterations")
            ev = torch.cuda.Event()
            ev.record(main)
            copy_stream.wait_event(ev)
            with torch.cuda.stream(copy_stream):
                host_chain[start:stop].copy_(cbuf[: stop - start], non_blocking=True)
                host_lp[start:stop].copy_(lbuf[: stop - start], non_blocking=True)
                done[b % 2] = torch.cuda.Event()
                done[b % 2].record(copy_stream)
        copy_stream.synchronize()
        main.synchronize()
        if int(flag.item()) != 0:
            raise ValueError(_NONFINITE_MSG)                                         # sampler_utils.py:74-75
        acc = accepts.cpu().numpy()
        if total:
            self.backend.save_slice(host_chain.numpy(), host_lp.numpy(), acc, 0)
        self.last_state = X
        self.diagnostics_main = {"accepts": acc, "acceptance_rate": acc / total if total else acc}     # :137-138
        return self.get_chain(discard=warmup, thin=thin_by)                          # :140
