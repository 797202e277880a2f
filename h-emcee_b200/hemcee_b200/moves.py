"""Hamiltonian walk / side moves with the reference's move protocol, executed by libhx.

    move(group1, group2, step_size, key, log_prob, grad_log_prob, L, log_prob_group1)
        -> (proposed, log_accept_prob, momentum_projected, proposed_log_prob)

(reference src/hemcee/moves/hamiltonian/hmc_walk.py:6-59, hmc_side.py:6-77).  `log_prob` is a
`targets.Target`; `grad_log_prob` is either the same target (analytic gradient in-kernel) or a
`targets.FrozenGrad` marker.  The functions keep the reference's `__name__`s because
`select_adapter` dispatches on them (reference src/hemcee/adaptation/adapter.py:101-103).

Multi-GPU: pass `row0=` / `n=` to update only a row shard of group 1 against the full complement;
the RNG counters use global row indices, so the result is independent of the sharding.
"""
from __future__ import annotations

import torch

from . import _lib, _rt
from .targets import FrozenGrad, as_target


def _move(fn_name, group1, group2, step_size, key, log_prob, grad_log_prob, L, log_prob_group1, row0, n, impl):
    tgt = as_target(log_prob)
    flags = _lib.HX_MOVE_FROZEN_GRAD if isinstance(grad_log_prob, FrozenGrad) else 0
    dev = _rt.device(group1.device)
    g1 = group1.contiguous()
    g2 = group2.contiguous()
    lp1 = log_prob_group1.contiguous()
    rows, d = g1.shape
    n = g2.shape[0] if n is None else int(n)
    if g2.shape != (n, d):
        raise ValueError(f"complement must have shape ({n}, {d}), got {tuple(g2.shape)}")
    if g1.dtype != g2.dtype or lp1.dtype != g1.dtype:
        raise TypeError("group1, group2 and log_prob_group1 must share a dtype")
    q = torch.empty_like(g1)
    pproj = torch.empty_like(g1)
    la = torch.empty(rows, dtype=g1.dtype, device=dev)
    lp = torch.empty(rows, dtype=g1.dtype, device=dev)
    desc = tgt.descriptor(g1.dtype, dev)
    fn = getattr(_lib.load(), fn_name)
    st = fn(_rt.ctx(dev), _rt.dtype_code(g1.dtype), _rt.impl_code(impl), desc, _rt.ptr(g1), _rt.ptr(g2), n, int(row0),
            rows, float(step_size), _lib.key_arg(_rt.as_key(key)), int(L), _rt.ptr(lp1), _rt.ptr(q), _rt.ptr(la),
            _rt.ptr(pproj), _rt.ptr(lp), flags, _rt.stream(dev))
    _lib.check(st, fn_name)
    return q, la, pproj, lp


def hmc_walk_move(group1, group2, step_size, key, log_prob, grad_log_prob, L, log_prob_group1, *, row0=0, n=None,
                  impl=None):
    """Hamiltonian walk move (Alg. 3 of arXiv:2505.02987); reference hmc_walk.py:6-59."""
    return _move("hx_walk_move", group1, group2, step_size, key, log_prob, grad_log_prob, L, log_prob_group1, row0, n, impl)


def walk_prefetch(key, n, row0, rows, device=None, impl=None):
    """Hint for the tensor-core walk path (include/hx.h hx_walk_prefetch): the walk half-move after the next
    `hmc_walk_move` call will use `key` on the row shard [row0, row0+rows) of a group of n."""
    dev = _rt.device(device)
    _lib.check(_lib.load().hx_walk_prefetch(_rt.ctx(dev), _rt.impl_code(impl), int(n), int(row0), int(rows),
                                            _lib.key_arg(_rt.as_key(key))), "hx_walk_prefetch")


def hmc_side_move(group1, group2, step_size, key, log_prob, grad_log_prob, L, log_prob_group1, *, row0=0, n=None,
                  impl=None):
    """Hamiltonian side move (Alg. 4 of arXiv:2505.02987); reference hmc_side.py:6-77."""
    return _move("hx_side_move", group1, group2, step_size, key, log_prob, grad_log_prob, L, log_prob_group1, row0, n, impl)


def leapfrog_walk_move(q, p_projected, log_prob, beta_eps, L, centered):
    """Walk leapfrog in projected form (reference hmc_walk.py:61-107).

    The reference carries the n x n momentum `p`; every use of it is through `S = p @ centered`,
    which is what this entry takes and returns: (q_L, S_L).  S_L equals the reference's
    `p_projected` (hmc_walk.py:105)."""
    tgt = as_target(log_prob)
    dev = _rt.device(q.device)
    q = q.clone().contiguous()
    S = p_projected.clone().contiguous()
    C = centered.contiguous()
    _lib.check(_lib.load().hx_leapfrog_walk(_rt.ctx(dev), _rt.dtype_code(q.dtype), tgt.descriptor(q.dtype, dev),
                                            _rt.ptr(q), _rt.ptr(S), _rt.ptr(C), C.shape[0], q.shape[0],
                                            float(beta_eps), int(L), None, _rt.stream(dev)), "hx_leapfrog_walk")
    return q, S


def leapfrog_side_move(q, p, log_prob, beta_eps, L, diff_particles_group2):
    """Side leapfrog (reference hmc_side.py:80-135): scalar momentum per walker along `diff` rows."""
    tgt = as_target(log_prob)
    dev = _rt.device(q.device)
    q = q.clone().contiguous()
    p = p.clone().contiguous()
    D = diff_particles_group2
    if D.ndim == 1:
        D = D[None, :].expand(q.shape[0], -1)
    D = D.contiguous()
    _lib.check(_lib.load().hx_leapfrog_side(_rt.ctx(dev), _rt.dtype_code(q.dtype), tgt.descriptor(q.dtype, dev),
                                            _rt.ptr(q), _rt.ptr(p), _rt.ptr(D), q.shape[0], float(beta_eps), int(L),
                                            _rt.stream(dev)), "hx_leapfrog_side")
    return q, p, D * p[:, None]


def accept_proposal(current, proposed, lp_current, lp_proposed, log_accept_prob, key, accepts=None, *, row0=0,
                    n=None, impl=None):
    """Metropolis accept (reference src/hemcee/samplers/sampler_utils.py:91-124) applied to positions and
    log-probs with one key, like the two calls at hamiltonian_ensemble.py:295-296.  Updates `current`
    and `lp_current` IN PLACE and returns (current, lp_current, accepts_int32)."""
    dev = _rt.device(current.device)
    rows, d = current.shape
    n = rows if n is None else int(n)
    if accepts is None:
        accepts = torch.zeros(rows, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().hx_accept(_rt.ctx(dev), _rt.dtype_code(current.dtype), _rt.impl_code(impl), _rt.ptr(current),
                                     _rt.ptr(proposed.contiguous()), _rt.ptr(lp_current), _rt.ptr(lp_proposed.contiguous()),
                                     _rt.ptr(log_accept_prob.contiguous()), _lib.key_arg(_rt.as_key(key)), n, int(row0),
                                     rows, d, _rt.ptr(accepts), _rt.stream(dev)), "hx_accept")
    return current, lp_current, accepts
