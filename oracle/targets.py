"""Log-densities (vectorised over walkers) + analytic gradients for the oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference differentiates an arbitrary per-sample callable with
`jax.vmap(jax.grad(log_prob))` (reference src/hemcee/samplers/sampler.py:52,55).
The oracle restates the *built-in* targets of the reference's test-suite with
hand-derived gradients (finite-difference checked in tests/test_oracle_moves.py):

  gaussian        reference src/hemcee/tests/distribution.py:30-48  (-1/2 (x-mu)^T Sigma^{-1} (x-mu);
                  the argument the reference calls `precision` is used as the covariance, :44-47)
  skewed cov      reference src/hemcee/tests/distribution.py:76-97  (Q diag(0.1*linspace(1,kappa,d)) Q^T)
  rosenbrock      reference src/hemcee/tests/distribution.py:143-156
  funnel          standard Neal funnel (nu ~ N(0,3^2), x_i|nu ~ N(0,e^nu)); the reference's own
                  `make_neal_funnel` (:158-175) has the wrong sign and is unused

Each factory returns `(log_prob, grad_log_prob)`, both mapping x[n,d] -> [n] / [n,d]
in the dtype of x.
"""
from __future__ import annotations

import numpy as np

from . import jaxrng


def make_gaussian(cov, mean):
    """distribution.py:30-48 with `precision`(sic)=cov.  Uses a solve like the reference."""
    cov64 = np.asarray(cov, dtype=np.float64)
    mean64 = np.asarray(mean, dtype=np.float64)
    prec64 = np.linalg.inv(cov64)
    prec64 = 0.5 * (prec64 + prec64.T)

    def log_prob(x):
        dt = x.dtype.type
        c = x - mean64.astype(dt)
        return (dt(-0.5) * np.einsum("nj,nj->n", c, c @ prec64.astype(dt))).astype(dt)

    def grad_log_prob(x):
        dt = x.dtype.type
        c = x - mean64.astype(dt)
        return (-(c @ prec64.astype(dt))).astype(dt)

    log_prob.precision = prec64
    log_prob.mean = mean64
    return log_prob, grad_log_prob


def make_covariance_skewed(key, dim, cond_number=1000, dtype=np.float64, impl=jaxrng.DEFAULT_IMPL):
    """distribution.py:76-97: Q diag(0.1*linspace(1, cond, dim)) Q^T, symmetrised."""
    dt = np.dtype(dtype).type
    eig = dt(0.1) * np.linspace(1, cond_number, dim, dtype=dt)
    H = jaxrng.normal(key, (dim, dim), dt, impl)
    Q, _ = np.linalg.qr(H)
    S = Q @ np.diag(eig) @ Q.T
    return (dt(0.5) * (S + S.T)).astype(dt)


def make_gaussian_skewed(key, dim, cond_number=1000, dtype=np.float64, impl=jaxrng.DEFAULT_IMPL):
    """distribution.py:50-74."""
    keys = jaxrng.split(key, 2, impl)
    cov = make_covariance_skewed(keys[0], dim, cond_number, dtype, impl)
    mean = jaxrng.normal(keys[1], (dim,), dtype, impl)
    return make_gaussian(cov, mean)


def make_iso_gaussian():
    """-1/2 |x|^2 (quickstart.ipynb cell 4; tests/test_proposal.py:47)."""
    def log_prob(x):
        dt = x.dtype.type
        return (dt(-0.5) * np.sum(x * x, axis=1, dtype=dt)).astype(dt)

    def grad_log_prob(x):
        return -x

    return log_prob, grad_log_prob


def make_rosenbrock(a=1.0, b=100.0):
    """distribution.py:155-156: -sum_i [ b (x_{i+1} - x_i^2)^2 + (a - x_i)^2 ]."""
    def log_prob(x):
        dt = x.dtype.type
        x0, x1 = x[:, :-1], x[:, 1:]
        return (-np.sum(dt(b) * (x1 - x0 ** 2) ** 2 + (dt(a) - x0) ** 2, axis=1, dtype=dt)).astype(dt)

    def grad_log_prob(x):
        dt = x.dtype.type
        x0, x1 = x[:, :-1], x[:, 1:]
        r = x1 - x0 ** 2
        g = np.zeros_like(x)
        g[:, :-1] += dt(4.0 * b) * r * x0 + dt(2.0) * (dt(a) - x0)
        g[:, 1:] += dt(-2.0 * b) * r
        return g

    return log_prob, grad_log_prob


def make_funnel(sigma=3.0):
    """Neal's funnel, x[:,0]=nu: -nu^2/(2 sigma^2) - 1/2 e^{-nu} sum x_i^2 - (d-1) nu / 2."""
    def log_prob(x):
        dt = x.dtype.type
        nu = x[:, 0]
        s = np.sum(x[:, 1:] ** 2, axis=1, dtype=dt)
        m = dt(x.shape[1] - 1)
        with np.errstate(over="ignore", invalid="ignore"):
            return (-nu * nu / dt(2.0 * sigma * sigma) - dt(0.5) * np.exp(-nu) * s - dt(0.5) * m * nu).astype(dt)

    def grad_log_prob(x):
        dt = x.dtype.type
        nu = x[:, 0]
        s = np.sum(x[:, 1:] ** 2, axis=1, dtype=dt)
        m = dt(x.shape[1] - 1)
        g = np.empty_like(x)
        with np.errstate(over="ignore", invalid="ignore"):
            e = np.exp(-nu)
            g[:, 0] = -nu / dt(sigma * sigma) + dt(0.5) * e * s - dt(0.5) * m
            g[:, 1:] = -(e[:, None] * x[:, 1:])
        return g

    return log_prob, grad_log_prob
