"""Integrated autocorrelation time (emcee's estimator), as in reference src/hemcee/autocorr.py:7-128.

Post-processing for the ESS metric, not a kernel target: batched torch.fft over
walkers, on whatever device the chain lives on.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["function_1d", "integrated_time", "AutocorrError"]


class AutocorrError(Exception):
    """Raised if the chain is too short to estimate an autocorrelation time (`.tau` has the estimate)."""

    def __init__(self, tau, *args, **kwargs):
        self.tau = tau
        super().__init__(*args, **kwargs)


def next_pow_two(n: int) -> int:
    i = 1
    while i < n:
        i <<= 1
    return i


def _acf(z: torch.Tensor) -> torch.Tensor:
    """Normalised ACF along axis 0 of z[n_t, ...] via zero-padded FFT (autocorr.py:15-34)."""
    n_t = z.shape[0]
    n = next_pow_two(n_t)
    f = torch.fft.fft(z - z.mean(dim=0, keepdim=True), n=2 * n, dim=0)
    acf = torch.fft.ifft(f * torch.conj(f), dim=0)[:n_t].real
    return acf / acf[0:1]


def function_1d(x):
    x = torch.atleast_1d(torch.as_tensor(x))
    if x.ndim != 1:
        raise ValueError("invalid dimensions for 1D autocorrelation function")
    return _acf(x.to(torch.float64))


def auto_window(taus: torch.Tensor, c) -> int:
    m = torch.arange(len(taus), device=taus.device) < c * taus
    if bool(torch.any(m)):
        return int(torch.argmin(m.to(torch.int8)))
    return len(taus) - 1


def integrated_time(x, c=5, tol=50, quiet=False, has_walkers=True):
    """x[n_step, n_walker, n_dim] -> tau[n_dim] (autocorr.py:44-118)."""
    x = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x)
    x = torch.atleast_1d(x).to(torch.float64)
    if x.ndim == 1:
        x = x[:, None, None]
    elif x.ndim == 2:
        x = x[:, None, :] if not has_walkers else x[:, :, None]
    if x.ndim != 3:
        raise ValueError("invalid dimensions")
    n_t, n_w, n_d = x.shape
    tau = np.empty(n_d)
    for d in range(n_d):
        f = _acf(x[:, :, d]).mean(dim=1)
        taus = 2.0 * torch.cumsum(f, 0) - 1.0
        tau[d] = float(taus[auto_window(taus, c)])
    flag = tol * tau > n_t
    if np.any(flag):
        msg = (f"The chain is shorter than {tol} times the integrated autocorrelation time for "
               f"{int(np.sum(flag))} parameter(s). Use this estimate with caution and run a longer chain!\n"
               f"N/{tol} = {n_t / tol:.0f};\ntau: {tau}")
        if not quiet:
            raise AutocorrError(tau, msg)
        print(msg)
    return tau
