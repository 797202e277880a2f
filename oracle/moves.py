"""NumPy restatement of the reference's Hamiltonian walk / side moves (dense form).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows, statement by statement:
  walk move        reference src/hemcee/moves/hamiltonian/hmc_walk.py:6-59
  walk leapfrog    reference src/hemcee/moves/hamiltonian/hmc_walk.py:61-107
  side move        reference src/hemcee/moves/hamiltonian/hmc_side.py:6-77
  side leapfrog    reference src/hemcee/moves/hamiltonian/hmc_side.py:80-135
  accept_proposal  reference src/hemcee/samplers/sampler_utils.py:91-124

The walk move here keeps the reference's *dense* n x n momentum on purpose: the
CUDA path uses the algebraically equivalent projected form, so the
oracle is an independent check of that algebra as well as of the arithmetic.
All arrays carry the dtype of `group1` (float32 = JAX default mode,
float64 = `jax_enable_x64`).
"""
from __future__ import annotations

import numpy as np

from . import jaxrng


def _dt(a):
    return np.asarray(a).dtype.type


def leapfrog_walk_move(q, p, grad_log_prob, beta_eps, L, centered):
    """hmc_walk.py:61-107.  q[n,d], p[n,n], centered[n,d] -> (q, p, p @ centered)."""
    dt = _dt(q)
    eps = dt(beta_eps)
    q = np.array(q, dtype=dt, copy=True)
    p = np.array(p, dtype=dt, copy=True)
    C = np.asarray(centered, dtype=dt)
    g = -grad_log_prob(q)                       # :85
    p = p - dt(0.5) * eps * (g @ C.T)           # :86
    for _ in range(int(L) - 1):                 # :89-96
        q = q + eps * (p @ C)
        g = -grad_log_prob(q)
        p = p - eps * (g @ C.T)
    q = q + eps * (p @ C)                       # :99
    g = -grad_log_prob(q)                       # :101
    p = p - dt(0.5) * eps * (g @ C.T)           # :102
    return q, p, p @ C                          # :105


def hmc_walk_move(group1, group2, step_size, key, log_prob, grad_log_prob, L, log_prob_group1,
                  impl=jaxrng.DEFAULT_IMPL, rng_dtype=None):
    """hmc_walk.py:6-59 -> (proposed, log_accept_prob, momentum_projected, proposed_log_prob).

    `rng_dtype` (test aid): draw the momentum in that dtype and upcast, so an f32 random stream can
    be integrated in f64 arithmetic as a high-precision yardstick for the f32 kernels."""
    dt = _dt(group1)
    rdt = dt if rng_dtype is None else np.dtype(rng_dtype).type
    g1 = np.asarray(group1, dtype=dt)
    g2 = np.asarray(group2, dtype=dt)
    n = g1.shape[0]
    centered2 = (g2 - np.mean(g2, axis=0, dtype=dt)[None, :]) / np.sqrt(dt(n))      # :35
    momentum = jaxrng.normal(key, (n, n), rdt, impl).astype(dt)                      # :36
    q, p, p_proj = leapfrog_walk_move(g1, momentum, grad_log_prob, step_size, L, centered2)
    current_U = -np.asarray(log_prob_group1, dtype=dt)                               # :47
    current_K = dt(0.5) * np.sum(momentum ** 2, axis=1, dtype=dt)                    # :48
    proposed_lp = np.asarray(log_prob(q), dtype=dt)                                  # :50
    proposed_U = -proposed_lp
    proposed_K = dt(0.5) * np.sum(p ** 2, axis=1, dtype=dt)                          # :52
    with np.errstate(invalid="ignore"):
        dH = (proposed_U + proposed_K) - (current_U + current_K)                     # :54
    dH = np.where(np.isnan(dH), dt(np.inf), dH)                                      # :55
    log_acc = np.minimum(dt(0.0), -dH)                                               # :56
    return q, log_acc.astype(dt), p_proj, proposed_lp


def leapfrog_side_move(q, p, grad_log_prob, beta_eps, L, diff):
    """hmc_side.py:80-135.  q[n,d], p[n], diff[n,d] (or [d]) -> (q, p, diff * p[:,None])."""
    dt = _dt(q)
    eps = dt(beta_eps)
    q = np.array(q, dtype=dt, copy=True)
    p = np.array(p, dtype=dt, copy=True)
    D = np.asarray(diff, dtype=dt)
    g = -grad_log_prob(q)
    p = p - dt(0.5) * eps * np.sum(g * D, axis=1, dtype=dt)       # :107-110
    for _ in range(int(L) - 1):                                   # :113-121
        q = q + eps * (p[:, None] * D)
        g = -grad_log_prob(q)
        p = p - eps * np.sum(g * D, axis=1, dtype=dt)
    q = q + eps * (p[:, None] * D)                                # :124
    g = -grad_log_prob(q)
    p = p - dt(0.5) * eps * np.sum(g * D, axis=1, dtype=dt)       # :127-130
    return q, p, D * p[:, None]                                   # :133


def side_choices(key_choices, n, impl=jaxrng.DEFAULT_IMPL):
    """hmc_side.py:43-47: per-walker `choice(k_i, arange(n), (2,), replace=False)` -> int[n,2]."""
    out = np.empty((len(key_choices), 2), dtype=np.int64)
    for i, k in enumerate(key_choices):
        out[i] = jaxrng.choice_no_replace(k, n, 2, impl)
    return out


def hmc_side_move(group1, group2, step_size, key, log_prob, grad_log_prob, L, log_prob_group1,
                  impl=jaxrng.DEFAULT_IMPL, rng_dtype=None):
    """hmc_side.py:6-77.  `rng_dtype`: see hmc_walk_move."""
    dt = _dt(group1)
    rdt = dt if rng_dtype is None else np.dtype(rng_dtype).type
    g1 = np.asarray(group1, dtype=dt)
    g2 = np.asarray(group2, dtype=dt)
    n, dim = g1.shape
    keys = jaxrng.split(key, n + 2, impl)                          # :37
    key_choices = keys[0:n]
    key_momentum = keys[-2]                                        # :39 (keys[-1] is never used)
    ch = side_choices(key_choices, n, impl)                        # :43-47
    diff = (g2[ch[:, 0]] - g2[ch[:, 1]]) / np.sqrt(dt(2 * dim))    # :52
    momentum = jaxrng.normal(key_momentum, (n,), rdt, impl).astype(dt)   # :54
    q, p, p_proj = leapfrog_side_move(g1, momentum, grad_log_prob, step_size, L, diff)
    current_U = -np.asarray(log_prob_group1, dtype=dt)
    current_K = dt(0.5) * momentum ** 2
    proposed_lp = np.asarray(log_prob(q), dtype=dt)
    proposed_U = -proposed_lp
    proposed_K = dt(0.5) * p ** 2
    with np.errstate(invalid="ignore"):
        dH = (proposed_U + proposed_K) - (current_U + current_K)   # :72
    dH = np.where(np.isnan(dH), dt(np.inf), dH)                    # :73
    log_acc = np.minimum(dt(0.0), -dH)                             # :74
    return q, log_acc.astype(dt), p_proj, proposed_lp


def accept_uniform_log(key, n, dt, impl=jaxrng.DEFAULT_IMPL):
    """sampler_utils.py:114: log(uniform(key, (n,), minval=1e-10, maxval=1.0))."""
    u = jaxrng.uniform(key, (n,), dt, 1e-10, 1.0, impl)
    return np.log(u).astype(dt)


def accept_proposal(current, proposed, log_accept_prob, key, impl=jaxrng.DEFAULT_IMPL):
    """sampler_utils.py:91-124 -> (updated, accepts int[n]).  Strict `<` compare (:115)."""
    dt = _dt(log_accept_prob)
    log_u = accept_uniform_log(key, np.shape(log_accept_prob)[0], dt, impl)
    mask = log_u < np.asarray(log_accept_prob)
    cur = np.asarray(current)
    m = mask.reshape((-1,) + (1,) * (cur.ndim - 1))
    return np.where(m, proposed, cur), mask.astype(np.int64)


hmc_walk_move.__name__ = "hmc_walk_move"
hmc_side_move.__name__ = "hmc_side_move"


# ---- derivative-free ensemble moves -----------------------------------------------------------------
def sample_inv_sqrt_density(key, a, shape, dt, impl=jaxrng.DEFAULT_IMPL):
    """moves/vanilla/stretch.py:58-76: z ~ z^{-1/2} on [1/a, a] by inversion, in the default float type."""
    a = dt(a)
    u = jaxrng.uniform(key, shape, dt, 0.0, 1.0, impl)
    sqrt_a = np.sqrt(a)
    inv_sqrt_a = dt(1.0) / sqrt_a
    t = inv_sqrt_a + u * (sqrt_a - inv_sqrt_a)
    return (t * t).astype(dt)


def stretch_move(group1, group2, key, log_prob, log_prob_group1, stretch=2.0, impl=jaxrng.DEFAULT_IMPL):
    """moves/vanilla/stretch.py:5-56 (Goodman-Weare stretch move, Alg. 1 of arXiv:2505.02987)."""
    dt = _dt(group1)
    n, dim = group1.shape
    keys = jaxrng.split(key, n + 2, impl)                                       # :35-38
    idx = np.array([jaxrng.choice_no_replace(k, n, 1, impl)[0] for k in keys[:n]])   # :40-46
    z = sample_inv_sqrt_density(keys[-1], stretch, (n,), dt, impl)              # :48
    xj = group2[idx]
    prop = (xj + z[:, None] * (group1 - xj)).astype(dt)                         # :49
    lpp = log_prob(prop)
    la = (dt(dim - 1) * np.log(z) + lpp - log_prob_group1).astype(dt)           # :53
    return prop, la, lpp


def side_move(group1, group2, key, log_prob, log_prob_group1, stretch=2.0, impl=jaxrng.DEFAULT_IMPL):
    """moves/vanilla/side.py:5-56 (Alg. 2)."""
    dt = _dt(group1)
    n = group1.shape[0]
    keys = jaxrng.split(key, n + 2, impl)                                       # :30-33
    ch = side_choices(keys[:n], n, impl)                                        # :35-39
    z = jaxrng.normal(keys[-2], (n, 1), dt, impl)                               # :44
    prop = (group1 + (dt(stretch) * z) * (group2[ch[:, 0]] - group2[ch[:, 1]])).astype(dt)   # :49
    lpp = log_prob(prop)
    return prop, (lpp - log_prob_group1).astype(dt), lpp                        # :52


def walk_move(group1, group2, key, log_prob, log_prob_group1, impl=jaxrng.DEFAULT_IMPL, **_):
    """moves/vanilla/walk.py:5-42 (Alg. 4): ONE shared random combination of the centred complement shifts every walker."""
    dt = _dt(group1)
    n = group1.shape[0]
    key_noise = jaxrng.split(key, 2, impl)[0]                                   # :29
    m = np.mean(group2, axis=0, dtype=dt)                                       # :32
    xi = jaxrng.normal(key_noise, (n,), dt, impl)                               # :33
    shift = (np.einsum("bi,b->i", group2 - m, xi) / np.sqrt(dt(n))).astype(dt)  # :35
    prop = (group1 + shift[None, :]).astype(dt)                                 # :36
    lpp = log_prob(prop)
    return prop, (lpp - log_prob_group1).astype(dt), lpp                        # :39


stretch_move.__name__ = "stretch_move"
side_move.__name__ = "side_move"
walk_move.__name__ = "walk_move"
