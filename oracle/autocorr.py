"""NumPy restatement of the reference's integrated autocorrelation time.

TEST INFRASTRUCTURE (see oracle/__init__.py).
Follows reference src/hemcee/autocorr.py:7-128 (itself derived from emcee).
"""
from __future__ import annotations

import numpy as np


class AutocorrError(Exception):
    def __init__(self, tau, *args):
        self.tau = tau
        super().__init__(*args)


def next_pow_two(n: int) -> int:
    i = 1
    while i < n:
        i <<= 1
    return i


def function_1d(x):
    """autocorr.py:15-34: normalised ACF of a 1-D series via zero-padded FFT."""
    x = np.atleast_1d(x)
    if x.ndim != 1:
        raise ValueError("invalid dimensions for 1D autocorrelation function")
    n = next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[: len(x)].real
    return acf / acf[0]


def auto_window(taus, c):
    """autocorr.py:37-41."""
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return int(np.argmin(m))
    return len(taus) - 1


def integrated_time(x, c=5, tol=50, quiet=False, has_walkers=True):
    """autocorr.py:44-118: x[n_step, n_walker, n_dim] -> tau[n_dim]."""
    x = np.atleast_1d(np.asarray(x))
    if x.ndim == 1:
        x = x[:, None, None]
    elif x.ndim == 2:
        x = x[:, None, :] if not has_walkers else x[:, :, None]
    if x.ndim != 3:
        raise ValueError("invalid dimensions")
    n_t, n_w, n_d = x.shape
    tau = np.empty(n_d)
    for d in range(n_d):
        z = x[:, :, d]
        n = next_pow_two(n_t)
        f = np.fft.fft(z - z.mean(axis=0), n=2 * n, axis=0)
        acf = np.fft.ifft(f * np.conjugate(f), axis=0)[:n_t].real
        acf = acf / acf[0]
        fm = acf.mean(axis=1)
        taus = 2.0 * np.cumsum(fm) - 1.0
        tau[d] = taus[auto_window(taus, c)]
    flag = tol * tau > n_t
    if np.any(flag):
        msg = (f"The chain is shorter than {tol} times the integrated autocorrelation time for "
               f"{int(np.sum(flag))} parameter(s). N/{tol} = {n_t / tol:.0f}; tau: {tau}")
        if not quiet:
            raise AutocorrError(tau, msg)
    return tau
