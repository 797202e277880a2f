"""Warm-up adapters with the reference's Adapter protocol (init / update / value / finalize).

  DAParameters, DualAveragingAdapter   reference src/hemcee/adaptation/dual_averaging.py:16-96
  ChEESParameters, ChEESAdapter        reference src/hemcee/adaptation/chees.py:18-219
  NoOpAdapter, CompositeAdapter        reference src/hemcee/adaptation/adapter.py:14-78
  select_adapter                       reference src/hemcee/adaptation/adapter.py:81-125
  find_reasonable_step_size            reference src/hemcee/adaptation/reasonable_stepsize.py:5-83

These are warm-up hooks, not the hot path: the per-walker reductions run on the
GPU (torch ops over the tensors the move kernels produced) and the handful of scalars of each update
is finished on the host in the run's float type.  With a row-sharded ensemble the partial sums are
combined through `allreduce` (identity on one GPU).  Reference quirks are kept.
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np
import torch

from . import random as hrandom
from .targets import FrozenGrad


class DAParameters(NamedTuple):
    target_accept: float = 0.651
    stepsize_inter: float = 0.9
    t0: float = 10.0
    gamma: float = 0.05
    kappa: float = 0.75
    agg: str = "harmonic"   # 'mean' or 'harmonic'


class DAState(NamedTuple):
    iteration: int
    log_stepsize: float
    log_stepsize_bar: float
    H_bar: float


class ChEESParameters(NamedTuple):
    T_min: float = 0.25
    T_max: float = 10.0
    T_interpolation: float = 0.9
    jitter: float = 0.6
    lr_T: float = 0.025
    beta1: float = 0.0
    beta2: float = 0.95
    regularization: float = 1e-7


class ChEESState(NamedTuple):
    log_T: float
    log_T_bar: float
    first_moment: float
    second_moment: float
    iteration: float
    halton: float


class NoOpState(NamedTuple):
    constant_step_size: float
    constant_integration_time: float


class CompositeState(NamedTuple):
    da_state: DAState
    chees_state: ChEESState


def _np_dt(dtype):
    return np.float64 if dtype in (torch.float64, np.float64) else np.float32


def _identity(t):
    return t


class Adapter:
    """Protocol of reference src/hemcee/adaptation/base.py:11-36."""

    def init(self, dim):
        raise NotImplementedError

    def update(self, state, log_accept_rate, **kwargs):
        raise NotImplementedError

    def value(self, state):
        raise NotImplementedError

    def finalize(self, state):
        raise NotImplementedError


def get_halton(index, base: int = 2) -> np.float32:
    """Radical inverse of int32(index) accumulated in float32 (chees.py:193-219)."""
    i = int(np.int32(index))
    f = np.float32(1.0)
    r = np.float32(0.0)
    while i > 0:
        f = np.float32(f / np.float32(base))
        r = np.float32(r + f * np.float32(i % base))
        i //= base
    return r


class DualAveragingAdapter(Adapter):
    def __init__(self, parameters: DAParameters, initial_step_size, initial_L, dtype=torch.float32,
                 allreduce=_identity):
        if parameters.agg not in ("mean", "harmonic"):
            raise ValueError('DAParmeter.agg must be either "mean" or "harmonic"')
        self.parameters = parameters
        self.initial_step_size = initial_step_size
        self.passthrough_L = initial_L
        self.dt = _np_dt(dtype)
        self.allreduce = allreduce

    def average_over_rate(self, log_accept):                      # dual_averaging.py:43-50
        a = torch.clamp(torch.exp(log_accept), 0.0, 1.0)
        if self.parameters.agg == "mean":
            s = self.allreduce(torch.stack([a.sum(), torch.tensor(float(a.numel()), dtype=a.dtype, device=a.device)]))
            s = s.cpu().numpy().astype(self.dt)
            return self.dt(s[0] / s[1])
        s = self.allreduce(torch.stack([(1.0 / a).sum(), torch.tensor(float(a.numel()), dtype=a.dtype, device=a.device)]))
        s = s.cpu().numpy().astype(self.dt)
        with np.errstate(divide="ignore"):
            return self.dt(s[1] / s[0])

    def init(self, dim):
        l0 = self.dt(np.log(self.dt(self.initial_step_size)))
        return DAState(0, l0, l0, self.dt(0.0))

    def update(self, state, log_accept_rate, **_):                # dual_averaging.py:64-89
        dt, P = self.dt, self.parameters
        it = state.iteration + 1
        rate = self.average_over_rate(log_accept_rate)
        eta = dt(1.0) / dt(it + P.t0)
        H_bar = (dt(1.0) - eta) * state.H_bar + eta * (dt(P.target_accept) - rate)
        soft_t = dt(it + P.t0)
        log_eps = dt(np.log(dt(self.initial_step_size))) - (np.sqrt(dt(it)) / (soft_t * dt(P.gamma))) * H_bar
        eta = dt(np.power(dt(it), dt(-P.kappa)))
        log_eps_bar = eta * log_eps + (dt(1.0) - eta) * state.log_stepsize_bar
        return DAState(it, dt(log_eps), dt(log_eps_bar), dt(H_bar))

    def value(self, state):
        return self.dt(np.exp(state.log_stepsize)), self.passthrough_L

    def finalize(self, state):
        return self.dt(np.exp(state.log_stepsize_bar)), self.passthrough_L


class ChEESAdapter(Adapter):
    def __init__(self, parameters: ChEESParameters, move_type, initial_step_size, initial_L, dtype=torch.float32,
                 allreduce=_identity):
        if move_type not in ("hmc_move", "hmc_walk_move", "hmc_side_move"):
            raise ValueError('Move type for ChEES tuner is not suppported. Choose between ["hmc", "hmc_walk", "hmc_side"]')
        self.parameters = parameters
        self.passthrough_step_size = initial_step_size
        self.initial_L = initial_L
        self.move_type = move_type
        self.dt = _np_dt(dtype)
        self.allreduce = allreduce

    def init(self, dim):
        l0 = self.dt(np.log(self.dt(self.initial_L * self.passthrough_step_size)))
        return ChEESState(l0, l0, self.dt(0.0), self.dt(0.0), self.dt(1.0), get_halton(1.0))

    def innerproduct(self, group1, group2, projected_momentum1):  # chees.py:54-73
        if self.move_type == "hmc_move":
            return (group1 * projected_momentum1).sum(dim=1)
        cov = torch.atleast_2d(torch.cov(group2.T))
        sol = torch.linalg.solve(cov, projected_momentum1.T).T
        return (group1 * sol).sum(dim=1)

    def update(self, state, log_accept_rate, position_current, position_proposed, momentum_proposed, group2,
               n_total=None):                                      # chees.py:89-166
        dt, P = self.dt, self.parameters
        acc = torch.clamp(torch.exp(log_accept_rate), 0.0, 1.0)
        rows = position_current.shape[0]
        n_total = rows if n_total is None else n_total
        sums = self.allreduce(torch.stack([position_current.sum(dim=0), position_proposed.sum(dim=0)]))
        mean_prev, mean_prop = sums[0] / n_total, sums[1] / n_total
        cen_prev = position_current - mean_prev
        cen_prop = position_proposed - mean_prop
        diff_sqnorm = (cen_prop ** 2).sum(dim=1) - (cen_prev ** 2).sum(dim=1)
        T = dt(np.exp(state.log_T))
        t_n = dt(state.halton) * T
        g_m = float(t_n) * diff_sqnorm * self.innerproduct(cen_prop, group2, momentum_proposed)
        g_m = torch.where(acc > 1e-4, g_m, torch.zeros_like(g_m))
        g_m = torch.where(torch.isfinite(g_m), g_m, torch.zeros_like(g_m))
        s = self.allreduce(torch.stack([(acc * g_m).sum(), acc.sum()])).cpu().numpy().astype(dt)
        g_hat = dt(s[0] / (s[1] + dt(P.regularization)))
        iteration = dt(state.iteration + dt(1.0))
        m1 = dt(P.beta1) * state.first_moment + dt(1 - P.beta1) * g_hat
        m2 = dt(P.beta2) * state.second_moment + dt(1 - P.beta2) * g_hat ** 2
        m1_hat = m1 / (dt(1.0) - dt(np.power(dt(P.beta1), iteration)))
        m2_hat = m2 / (dt(1.0) - dt(np.power(dt(P.beta2), iteration)))
        upd = dt(P.lr_T) * m1_hat / np.sqrt(m2_hat + dt(P.regularization))
        lo, hi = dt(np.log(dt(P.T_min))), dt(np.log(dt(P.T_max)))
        log_T = dt(np.clip(state.log_T + upd, lo, hi))
        a = P.T_interpolation
        log_T_bar = dt(np.logaddexp(dt(np.log(dt(a))) + state.log_T_bar, dt(np.log(dt(1.0 - a))) + log_T))
        log_T_bar = dt(np.clip(log_T_bar, lo, hi))
        return ChEESState(log_T, log_T_bar, dt(m1), dt(m2), iteration, get_halton(iteration))

    def _get_T(self, state, bar=False):                           # chees.py:182-190
        dt, P = self.dt, self.parameters
        T = dt(np.exp(state.log_T_bar if bar else state.log_T))
        return dt(dt(1.0 - P.jitter) * T + dt(P.jitter) * dt(state.halton) * T)

    def _length(self, T, step):
        return int(max(np.ceil(self.dt(T) / self.dt(step)), 1))

    def value(self, state):
        return self.passthrough_step_size, self._length(self._get_T(state), self.passthrough_step_size)

    def finalize(self, state):
        return self.passthrough_step_size, self._length(self._get_T(state, bar=True), self.passthrough_step_size)


class NoOpAdapter(Adapter):
    def __init__(self, initial_step_size, initial_L, **_):
        self.constant_step_size = initial_step_size
        self.constant_integration_time = initial_L

    def init(self, dim):
        return NoOpState(self.constant_step_size, self.constant_integration_time)

    def update(self, state, **kwargs):
        return state

    def value(self, state):
        return state.constant_step_size, state.constant_integration_time

    def finalize(self, state):
        return state.constant_step_size, state.constant_integration_time


class CompositeAdapter(Adapter):
    def __init__(self, da_parameters, chees_parameters, move_type, initial_step_size, initial_L, dtype=torch.float32,
                 allreduce=_identity):
        self.da_adapter = DualAveragingAdapter(da_parameters, initial_step_size, initial_L, dtype, allreduce)
        self.chees_adapter = ChEESAdapter(chees_parameters, move_type, initial_step_size, initial_L, dtype, allreduce)

    def init(self, dim):
        return CompositeState(self.da_adapter.init(dim), self.chees_adapter.init(dim))

    def update(self, state, **kwargs):
        da = self.da_adapter.update(state.da_state, log_accept_rate=kwargs["log_accept_rate"])
        ch = self.chees_adapter.update(state.chees_state, **kwargs)
        return CompositeState(da, ch)

    def value(self, state):
        step, _ = self.da_adapter.value(state.da_state)
        T = self.chees_adapter._get_T(state.chees_state)
        return step, self.chees_adapter._length(T, step)

    def finalize(self, state):
        step, _ = self.da_adapter.finalize(state.da_state)
        T = self.chees_adapter._get_T(state.chees_state, bar=True)
        return step, self.chees_adapter._length(T, step)


def select_adapter(da_parameters, chees_parameters, move, initial_step_size, initial_L, dtype=torch.float32,
                   allreduce=_identity) -> Adapter:
    """reference adapter.py:81-125 — dispatch on `move.__name__`."""
    move_type = getattr(move, "__name__", None)
    if move_type not in ["hmc_move", "hmc_walk_move", "hmc_side_move"]:
        raise ValueError(f'Move type for ChEES tuner is not suppported. Choose between ["hmc_move", "hmc_walk_move", '
                         f'"hmc_side_move"]. Got {move_type}')
    if da_parameters is True:
        da_parameters = DAParameters()
    if chees_parameters is True:
        chees_parameters = ChEESParameters()
    da_on = da_parameters is not False
    ch_on = chees_parameters is not False
    if not da_on and not ch_on:
        return NoOpAdapter(initial_step_size, initial_L)
    if da_on and not ch_on:
        return DualAveragingAdapter(da_parameters, initial_step_size, initial_L, dtype, allreduce)
    if ch_on and not da_on:
        return ChEESAdapter(chees_parameters, move_type, initial_step_size, initial_L, dtype, allreduce)
    return CompositeAdapter(da_parameters, chees_parameters, move_type, initial_step_size, initial_L, dtype, allreduce)


def find_reasonable_step_size(key, group1, group2, log_prob, grad_log_prob, init_step_size, proposal,
                              impl=None, max_iter=200):
    """reference reasonable_stepsize.py:5-83: double/halve the step until the direction flips, with
    L=1, the gradient frozen at the initial positions and the `N - logsumexp(-log_acc)` statistic."""
    dt = _np_dt(group1.dtype)
    target = dt(np.log(dt(0.8)))
    lp1 = log_prob.log_prob(group1)
    lp2 = log_prob.log_prob(group2)
    frozen = FrozenGrad(log_prob)
    finfo = np.finfo(dt)
    step, last_dir, direction = dt(init_step_size), 0, 0
    key_state = np.asarray(key, dtype=np.uint32)

    def cond(step, last_dir, direction):
        not_small = (step > finfo.tiny) or (direction >= 0)
        not_large = (step < finfo.max) or (direction <= 0)
        return not_small and not_large and (last_dir == 0 or direction == last_dir)

    it = 0
    while cond(step, last_dir, direction) and it < max_iter:
        ks = hrandom.split(key_state, 3, group1.device, impl)
        key_state, k1, k2 = ks[0], ks[1], ks[2]
        with np.errstate(over="ignore"):
            step = dt(dt(2.0) ** direction * step)
        _, la1, _, _ = proposal(group1, group2, step, k1, log_prob, frozen, L=1, log_prob_group1=lp1)
        _, la2, _, _ = proposal(group2, group1, step, k2, log_prob, frozen, L=1, log_prob_group1=lp2)
        la = torch.cat([la1, la2])
        avg = dt(la.shape[0]) - dt(torch.logsumexp(-la, dim=0).item())       # :63 (N, not log N)
        new_dir = 1 if target < avg else -1
        last_dir, direction = direction, new_dir
        it += 1
    return step


# ---- device-resident warm-up (include/hx.h section 4b): adapter objects <-> the C structs ------------------------------
def device_params(adapter):
    """hx_adapt_params for a DualAveraging / ChEES / Composite adapter, or None if the adapter has no device form."""
    from . import _lib
    da = ch = None
    if isinstance(adapter, CompositeAdapter):
        da, ch = adapter.da_adapter, adapter.chees_adapter
    elif isinstance(adapter, DualAveragingAdapter):
        da = adapter
    elif isinstance(adapter, ChEESAdapter):
        ch = adapter
    else:
        return None
    if ch is not None and ch.move_type == "hmc_move":
        return None
    p = _lib.hx_adapt_params()
    p.use_da, p.use_chees = int(da is not None), int(ch is not None)
    if da is not None:
        P = da.parameters
        p.agg_harmonic = int(P.agg == "harmonic")
        p.da_target_accept, p.da_t0, p.da_gamma, p.da_kappa = P.target_accept, P.t0, P.gamma, P.kappa
        p.da_init_step = float(da.dt(da.initial_step_size))
        p.pass_L = int(da.passthrough_L)
    if ch is not None:
        P = ch.parameters
        p.ch_T_min, p.ch_T_max, p.ch_T_interp, p.ch_jitter = P.T_min, P.T_max, P.T_interpolation, P.jitter
        p.ch_lr, p.ch_beta1, p.ch_beta2, p.ch_reg = P.lr_T, P.beta1, P.beta2, P.regularization
        p.pass_step = float(ch.dt(ch.passthrough_step_size))
    return p


def state_to_device(adapter, state):
    from . import _lib
    s = _lib.hx_adapt_state()
    da = state.da_state if isinstance(state, CompositeState) else (state if isinstance(state, DAState) else None)
    ch = state.chees_state if isinstance(state, CompositeState) else (state if isinstance(state, ChEESState) else None)
    if da is not None:
        s.da_iteration, s.da_log_step, s.da_log_step_bar, s.da_H_bar = int(da.iteration), float(da.log_stepsize), \
            float(da.log_stepsize_bar), float(da.H_bar)
    if ch is not None:
        s.ch_log_T, s.ch_log_T_bar, s.ch_m1, s.ch_m2 = float(ch.log_T), float(ch.log_T_bar), float(ch.first_moment), \
            float(ch.second_moment)
        s.ch_iteration, s.ch_halton = float(ch.iteration), float(ch.halton)
    return s


def state_from_device(adapter, state, s):
    """Rebuild the adapter's NamedTuple state from the C struct (values already are in the run's float type)."""
    dt = adapter.da_adapter.dt if isinstance(adapter, CompositeAdapter) else adapter.dt
    da = DAState(int(s.da_iteration), dt(s.da_log_step), dt(s.da_log_step_bar), dt(s.da_H_bar))
    ch = ChEESState(dt(s.ch_log_T), dt(s.ch_log_T_bar), dt(s.ch_m1), dt(s.ch_m2), dt(s.ch_iteration), np.float32(s.ch_halton))
    if isinstance(state, CompositeState):
        return CompositeState(da, ch)
    return da if isinstance(state, DAState) else ch
