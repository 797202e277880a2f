"""Integrated autocorrelation time (emcee's estimator), as in reference src/hemcee/autocorr.py:7-128.

Batched over walkers AND parameters: one real FFT of the zero-padded (n_step, n_walker x n_param) block per
chunk of columns (cuFFT when the chain is a CUDA tensor — the ESS leg of bench.py keeps its chain subset on the
device — pocketfft for host chains), the window search vectorised over parameters.  The reference loops over
parameters and vmaps `function_1d` over walkers (autocorr.py:95-104); the arithmetic per series is the same:
subtract the series mean, FFT with n = 2 * next_pow_two(n_step), |f|^2, inverse FFT, normalise by lag 0, average
over walkers, tau(M) = 2 cumsum - 1, first M with M >= c tau(M).
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["function_1d", "integrated_time", "AutocorrError"]

_CHUNK_ELEMS = 1 << 26        # padded (time x column) elements per FFT batch: 512 MB of float64 per buffer


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
    f = torch.fft.rfft(z - z.mean(dim=0, keepdim=True), n=2 * n, dim=0)
    acf = torch.fft.irfft(f.real * f.real + f.imag * f.imag, n=2 * n, dim=0)[:n_t]
    return acf / acf[0:1]


def function_1d(x):
    x = torch.atleast_1d(torch.as_tensor(x))
    if x.ndim != 1:
        raise ValueError("invalid dimensions for 1D autocorrelation function")
    return _acf(x.to(torch.float64))


def auto_window(taus: torch.Tensor, c) -> torch.Tensor:
    """taus[n_t, n_d] -> window[n_d]: first M with M >= c taus[M] (autocorr.py:37-41, including its all-true case)."""
    n_t = taus.shape[0]
    m = torch.arange(n_t, device=taus.device)[:, None] < c * taus
    first_false = torch.argmin(m.to(torch.int8), dim=0)          # first minimal value: the first False, 0 if all True
    return torch.where(m.any(dim=0), first_false, torch.full_like(first_false, n_t - 1))


def mean_acf(x: torch.Tensor) -> torch.Tensor:
    """x[n_t, n_w, n_d] (float64) -> walker-averaged normalised ACF [n_t, n_d]."""
    n_t, n_w, n_d = x.shape
    out = torch.empty((n_t, n_d), dtype=torch.float64, device=x.device)
    per = max(1, _CHUNK_ELEMS // (2 * next_pow_two(n_t) * max(n_w, 1)))
    for d0 in range(0, n_d, per):
        out[:, d0:d0 + per] = _acf(x[:, :, d0:d0 + per]).mean(dim=1)
    return out


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
    n_t = x.shape[0]
    taus = 2.0 * torch.cumsum(mean_acf(x), dim=0) - 1.0
    win = auto_window(taus, c)
    tau = taus.gather(0, win[None, :])[0].cpu().numpy()
    flag = tol * tau > n_t
    if np.any(flag):
        msg = (f"The chain is shorter than {tol} times the integrated autocorrelation time for "
               f"{int(np.sum(flag))} parameter(s). Use this estimate with caution and run a longer chain!\n"
               f"N/{tol} = {n_t / tol:.0f};\ntau: {tau}")
        if not quiet:
            raise AutocorrError(tau, msg)
        print(msg)
    return tau
