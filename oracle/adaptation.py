"""NumPy restatement of the reference's warm-up adapters and step-size heuristic.

TEST INFRASTRUCTURE (see oracle/__init__.py).

  DAParameters / DualAveragingAdapter   reference src/hemcee/adaptation/dual_averaging.py:16-96
  ChEESParameters / ChEESAdapter        reference src/hemcee/adaptation/chees.py:18-190
  get_halton                            reference src/hemcee/adaptation/chees.py:193-219
  NoOpAdapter / CompositeAdapter        reference src/hemcee/adaptation/adapter.py:14-78
  select_adapter                        reference src/hemcee/adaptation/adapter.py:81-125
  find_reasonable_step_size             reference src/hemcee/adaptation/reasonable_stepsize.py:5-83

Quirks of the reference are preserved on purpose: DA divides by
(t + t0) twice, the harmonic-mean acceptance statistic, ChEES `iteration`
starting at 1.0 with an f32 Halton sequence, `finalize` applying the current
jitter, and the heuristic's `N - logsumexp` "average" with a frozen gradient.

Every adapter is constructed with the float type `dt` the run computes in
(np.float32 = JAX default, np.float64 = x64) so scalar state rounds like the
reference's traced scalars do.
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np

from . import jaxrng


class DAParameters(NamedTuple):
    target_accept: float = 0.651
    stepsize_inter: float = 0.9
    t0: float = 10.0
    gamma: float = 0.05
    kappa: float = 0.75
    agg: str = "harmonic"


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


def get_halton(index, base: int = 2) -> np.float32:
    """chees.py:193-219: radical inverse of int32(index), accumulated in float32."""
    i = int(np.int32(index))
    f = np.float32(1.0)
    r = np.float32(0.0)
    while i > 0:
        f = np.float32(f / np.float32(base))
        r = np.float32(r + f * np.float32(i % base))
        i //= base
    return r


def _accept_prob(log_accept, dt):
    with np.errstate(over="ignore"):
        return np.clip(np.exp(np.asarray(log_accept, dtype=dt)), dt(0.0), dt(1.0))


class DualAveragingAdapter:
    def __init__(self, parameters: DAParameters, initial_step_size, initial_L, dt=np.float64):
        if parameters.agg not in ("mean", "harmonic"):
            raise ValueError('DAParmeter.agg must be either "mean" or "harmonic"')
        self.parameters = parameters
        self.initial_step_size = initial_step_size
        self.passthrough_L = initial_L
        self.dt = np.dtype(dt).type

    def average_over_rate(self, log_accept):          # dual_averaging.py:43-50
        dt = self.dt
        a = _accept_prob(log_accept, dt)
        if self.parameters.agg == "mean":
            return dt(np.mean(a, dtype=dt))
        with np.errstate(divide="ignore"):
            return dt(dt(a.size) / np.sum(dt(1.0) / a, dtype=dt))

    def init(self, dim):                              # :55-62
        l0 = self.dt(np.log(self.dt(self.initial_step_size)))
        return DAState(0, l0, l0, self.dt(0.0))

    def update(self, state, log_accept_rate, **_):    # :64-89
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


class ChEESAdapter:
    def __init__(self, parameters: ChEESParameters, move_type, initial_step_size, initial_L, dt=np.float64):
        if move_type not in ("hmc_move", "hmc_walk_move", "hmc_side_move"):
            raise ValueError("Move type for ChEES tuner is not suppported.")
        self.parameters = parameters
        self.passthrough_step_size = initial_step_size
        self.initial_L = initial_L
        self.move_type = move_type
        self.dt = np.dtype(dt).type

    def innerproduct(self, group1, group2, projected_momentum1):     # chees.py:54-73
        dt = self.dt
        if self.move_type == "hmc_move":
            return np.einsum("bd,bd->b", group1, projected_momentum1)
        cov = np.atleast_2d(np.cov(np.asarray(group2, dtype=dt), rowvar=False)).astype(dt)
        sol = np.linalg.solve(cov, np.asarray(projected_momentum1, dtype=dt).T).T.astype(dt)
        return np.sum(group1 * sol, axis=1, dtype=dt)

    def init(self, dim):                                             # :78-87
        l0 = self.dt(np.log(self.dt(self.initial_L * self.passthrough_step_size)))
        return ChEESState(l0, l0, self.dt(0.0), self.dt(0.0), self.dt(1.0), get_halton(1.0))

    def update(self, state, log_accept_rate, position_current, position_proposed,
               momentum_proposed, group2):                           # :89-166
        dt, P = self.dt, self.parameters
        acc = _accept_prob(log_accept_rate, dt)
        xc = np.asarray(position_current, dtype=dt)
        xp = np.asarray(position_proposed, dtype=dt)
        cen_prev = xc - np.mean(xc, axis=0, dtype=dt)
        cen_prop = xp - np.mean(xp, axis=0, dtype=dt)
        with np.errstate(invalid="ignore", over="ignore"):
            diff_sqnorm = np.sum(cen_prop ** 2, axis=1, dtype=dt) - np.sum(cen_prev ** 2, axis=1, dtype=dt)
            T = dt(np.exp(state.log_T))
            t_n = dt(state.halton) * T
            g_m = t_n * diff_sqnorm * self.innerproduct(cen_prop, group2, momentum_proposed)
            g_m = np.where(acc > dt(1e-4), g_m, dt(0.0))
            g_m = np.where(np.isfinite(g_m), g_m, dt(0.0)).astype(dt)
            g_hat = dt(np.sum(acc * g_m, dtype=dt) / (np.sum(acc, dtype=dt) + dt(P.regularization)))
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

    def _get_T(self, state, bar=False):                              # :182-190
        dt, P = self.dt, self.parameters
        T = dt(np.exp(state.log_T_bar if bar else state.log_T))
        return dt(dt(1.0 - P.jitter) * T + dt(P.jitter) * dt(state.halton) * T)

    def _length(self, T, step):
        return int(max(np.ceil(self.dt(T) / self.dt(step)), 1))

    def value(self, state):
        return self.passthrough_step_size, self._length(self._get_T(state), self.passthrough_step_size)

    def finalize(self, state):
        return self.passthrough_step_size, self._length(self._get_T(state, bar=True), self.passthrough_step_size)


class NoOpAdapter:
    def __init__(self, initial_step_size, initial_L, dt=np.float64):
        self.constant_step_size = initial_step_size
        self.constant_integration_time = initial_L

    def init(self, dim):
        return NoOpState(self.constant_step_size, self.constant_integration_time)

    def update(self, state, **_):
        return state

    def value(self, state):
        return state.constant_step_size, state.constant_integration_time

    finalize = value


class CompositeAdapter:
    def __init__(self, da_parameters, chees_parameters, move_type, initial_step_size, initial_L, dt=np.float64):
        self.da_adapter = DualAveragingAdapter(da_parameters, initial_step_size, initial_L, dt)
        self.chees_adapter = ChEESAdapter(chees_parameters, move_type, initial_step_size, initial_L, dt)

    def init(self, dim):
        return CompositeState(self.da_adapter.init(dim), self.chees_adapter.init(dim))

    def update(self, state, **kw):                                   # adapter.py:58-62
        da = self.da_adapter.update(state.da_state, log_accept_rate=kw["log_accept_rate"])
        ch = self.chees_adapter.update(state.chees_state, **kw)
        return CompositeState(da, ch)

    def value(self, state):                                          # adapter.py:64-70
        step, _ = self.da_adapter.value(state.da_state)
        T = self.chees_adapter._get_T(state.chees_state)
        return step, self.chees_adapter._length(T, step)

    def finalize(self, state):                                       # adapter.py:72-78
        step, _ = self.da_adapter.finalize(state.da_state)
        T = self.chees_adapter._get_T(state.chees_state, bar=True)
        return step, self.chees_adapter._length(T, step)


def select_adapter(da_parameters, chees_parameters, move, initial_step_size, initial_L, dt=np.float64):
    """adapter.py:81-125 (dispatch on `move.__name__`)."""
    move_type = getattr(move, "__name__", None)
    if move_type not in ("hmc_move", "hmc_walk_move", "hmc_side_move"):
        raise ValueError(f"Move type for ChEES tuner is not suppported. Got {move_type}")
    if da_parameters is True:
        da_parameters = DAParameters()
    if chees_parameters is True:
        chees_parameters = ChEESParameters()
    da_on = da_parameters is not False
    ch_on = chees_parameters is not False
    if not da_on and not ch_on:
        return NoOpAdapter(initial_step_size, initial_L, dt)
    if da_on and not ch_on:
        return DualAveragingAdapter(da_parameters, initial_step_size, initial_L, dt)
    if ch_on and not da_on:
        return ChEESAdapter(chees_parameters, move_type, initial_step_size, initial_L, dt)
    return CompositeAdapter(da_parameters, chees_parameters, move_type, initial_step_size, initial_L, dt)


def _logsumexp(a, dt):
    a = np.asarray(a, dtype=dt)
    m = np.max(a)
    if not np.isfinite(m):
        return dt(m)
    return dt(m + np.log(np.sum(np.exp(a - m), dtype=dt)))


def find_reasonable_step_size(key, group1, group2, log_prob, grad_log_prob, init_step_size, proposal,
                              impl=jaxrng.DEFAULT_IMPL, trace=None, max_iter=200):
    """reasonable_stepsize.py:5-83.  `proposal` is an oracle move (oracle.moves)."""
    dt = np.asarray(group1).dtype.type
    target = dt(np.log(dt(0.8)))
    lp1, lp2 = log_prob(group1), log_prob(group2)
    g1, g2 = grad_log_prob(group1), grad_log_prob(group2)
    frozen1 = lambda x: g1            # :37-38: the gradient is frozen at the initial positions
    frozen2 = lambda x: g2
    finfo = np.finfo(dt)
    step, last_dir, direction, key_state = dt(init_step_size), 0, 0, np.asarray(key, dtype=np.uint32)

    def cond(step, last_dir, direction):
        not_small = (step > finfo.tiny) or (direction >= 0)
        not_large = (step < finfo.max) or (direction <= 0)
        return not_small and not_large and (last_dir == 0 or direction == last_dir)

    it = 0
    while cond(step, last_dir, direction) and it < max_iter:
        ks = jaxrng.split(key_state, 3, impl)                       # :44
        key_state, k1, k2 = ks[0], ks[1], ks[2]
        with np.errstate(over="ignore"):
            step = dt(dt(2.0) ** direction * step)                   # :47
        _, la1, _, _ = proposal(group1, group2, step, k1, log_prob, frozen1, 1, lp1, impl=impl)
        _, la2, _, _ = proposal(group2, group1, step, k2, log_prob, frozen2, 1, lp2, impl=impl)
        la = np.concatenate([la1, la2])
        avg = dt(la.shape[0]) - _logsumexp(-la, dt)                  # :63 (N, not log N)
        new_dir = 1 if target < avg else -1                          # :66
        if trace is not None:
            trace.append((float(step), float(avg), new_dir))
        last_dir, direction = direction, new_dir
        it += 1
    return step
