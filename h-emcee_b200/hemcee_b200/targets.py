"""Built-in targets (log-density + analytic gradient evaluated by libhx kernels).

The reference takes an arbitrary per-sample JAX callable and differentiates it with
`jax.vmap(jax.grad(...))` (reference src/hemcee/samplers/sampler.py:52,55).  This package has no
autodiff and no CPU fallback; instead the reference's own test targets
(reference src/hemcee/tests/distribution.py:30-48,76-97,143-156, funnel per SURVEY §8a row T) are
descriptors handed to the CUDA kernels.  Each target is also a plain NumPy callable
`target(x[d]) -> float`, so the same object can be given to an independent implementation.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, _rt


class Target:
    kind = -1

    def __init__(self, dim: int, lower0=None):
        self.dim = int(dim)
        self.lower0 = lower0
        self._cache = {}

    # -- host-side (NumPy, float64) definition ------------------------------------------------
    def _logp_np(self, x):
        raise NotImplementedError

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        if self.lower0 is not None and x[0] < self.lower0:
            return -np.inf
        return float(self._logp_np(x))

    # -- device descriptor -----------------------------------------------------------------------
    def _fill(self, desc, dtype, dev, keep):
        pass

    def descriptor(self, dtype: torch.dtype, dev: torch.device) -> _lib.hx_target:
        k = (dtype, dev.index)
        if k not in self._cache:
            d = _lib.hx_target()
            d.kind, d.dim, d.ld = self.kind, self.dim, 0
            d.has_lower0 = 0 if self.lower0 is None else 1
            d.lower0 = 0.0 if self.lower0 is None else float(self.lower0)
            d.mean = None
            d.precision = None
            d.a = d.b = 0.0
            keep = []
            self._fill(d, dtype, dev, keep)
            self._cache[k] = (d, keep)
        return self._cache[k][0]

    # -- vectorised device evaluation (what BaseSampler builds with jit(vmap(...))) ----------------
    def log_prob(self, X: torch.Tensor) -> torch.Tensor:
        return self._eval(X, False)[0]

    def grad_log_prob(self, X: torch.Tensor) -> torch.Tensor:
        return self._eval(X, True)[1]

    def _eval(self, X, want_grad):
        dev = _rt.device(X.device)
        X = X.contiguous()
        if X.ndim != 2 or X.shape[1] != self.dim:
            raise ValueError(f"expected positions of shape (n, {self.dim}), got {tuple(X.shape)}")
        lp = torch.empty(X.shape[0], dtype=X.dtype, device=dev)
        g = torch.empty_like(X) if want_grad else None
        d = self.descriptor(X.dtype, dev)
        _lib.check(_lib.load().hx_log_prob(_rt.ctx(dev), _rt.dtype_code(X.dtype), d, _rt.ptr(X), X.shape[0],
                                           _rt.ptr(lp), _rt.ptr(g), _rt.stream(dev)), "hx_log_prob")
        return lp, g


class GaussianIso(Target):
    """-1/2 |x - mean|^2 (docs/tutorials/quickstart.ipynb cell 4)."""
    kind = _lib.HX_TARGET_GAUSS_ISO

    def __init__(self, dim, mean=None, lower0=None):
        super().__init__(dim, lower0)
        self.mean = None if mean is None else np.asarray(mean, dtype=np.float64).reshape(self.dim)

    def _logp_np(self, x):
        c = x if self.mean is None else x - self.mean
        return -0.5 * float(c @ c)

    def _fill(self, desc, dtype, dev, keep):
        if self.mean is not None:
            m = torch.as_tensor(self.mean, dtype=dtype, device=dev).contiguous()
            keep.append(m)
            desc.mean = m.data_ptr()


class GaussianFull(Target):
    """-1/2 (x - mean)^T P (x - mean).  Give `precision=P` or `cov=Sigma` (P = Sigma^-1 computed in
    float64 on the host), like reference src/hemcee/tests/distribution.py:30-48 which solves against
    the covariance."""
    kind = _lib.HX_TARGET_GAUSS_FULL

    def __init__(self, precision=None, cov=None, mean=None, lower0=None):
        if (precision is None) == (cov is None):
            raise ValueError("give exactly one of precision= or cov=")
        if precision is None:
            cov = np.asarray(cov, dtype=np.float64)
            precision = np.linalg.inv(cov)
        P = np.asarray(precision, dtype=np.float64)
        if P.ndim != 2 or P.shape[0] != P.shape[1]:
            raise ValueError("precision must be square")
        self.precision = 0.5 * (P + P.T)
        super().__init__(P.shape[0], lower0)
        self.mean = None if mean is None else np.asarray(mean, dtype=np.float64).reshape(self.dim)

    def _logp_np(self, x):
        c = x if self.mean is None else x - self.mean
        return -0.5 * float(c @ self.precision @ c)

    def _fill(self, desc, dtype, dev, keep):
        ld = (self.dim + 3) // 4 * 4
        P = torch.zeros((self.dim, ld), dtype=dtype, device=dev)
        P[:, : self.dim] = torch.as_tensor(self.precision, dtype=dtype, device=dev)
        keep.append(P)
        desc.precision = P.data_ptr()
        desc.ld = ld
        if self.mean is not None:
            m = torch.as_tensor(self.mean, dtype=dtype, device=dev).contiguous()
            keep.append(m)
            desc.mean = m.data_ptr()


class Rosenbrock(Target):
    """-sum_i [ b (x_{i+1} - x_i^2)^2 + (a - x_i)^2 ] (reference src/hemcee/tests/distribution.py:155-156)."""
    kind = _lib.HX_TARGET_ROSENBROCK

    def __init__(self, dim, a=1.0, b=100.0, lower0=None):
        super().__init__(dim, lower0)
        self.a, self.b = float(a), float(b)

    def _logp_np(self, x):
        return -float(np.sum(self.b * (x[1:] - x[:-1] ** 2) ** 2 + (self.a - x[:-1]) ** 2))

    def _fill(self, desc, dtype, dev, keep):
        desc.a, desc.b = self.a, self.b


class Funnel(Target):
    """Neal's funnel with x[0] = nu ~ N(0, sigma^2), x[i>0] | nu ~ N(0, e^nu)."""
    kind = _lib.HX_TARGET_FUNNEL

    def __init__(self, dim, sigma=3.0, lower0=None):
        super().__init__(dim, lower0)
        self.sigma = float(sigma)

    def _logp_np(self, x):
        nu = x[0]
        return float(-nu * nu / (2 * self.sigma ** 2) - 0.5 * np.exp(-nu) * np.sum(x[1:] ** 2) - 0.5 * (self.dim - 1) * nu)

    def _fill(self, desc, dtype, dev, keep):
        desc.a = self.sigma


class FrozenGrad:
    """Marker handed as `grad_log_prob` to freeze the gradient at the initial positions
    (reference src/hemcee/adaptation/reasonable_stepsize.py:34-38)."""

    def __init__(self, target: Target):
        self.target = target


def as_target(log_prob) -> Target:
    t = getattr(log_prob, "target", log_prob)
    if not isinstance(t, Target):
        raise TypeError(
            "hemcee_b200 evaluates built-in targets on the GPU (GaussianIso, GaussianFull, Rosenbrock, Funnel); "
            "arbitrary Python callables are not supported (no autodiff, no CPU fallback).")
    return t
