"""On-device chain statistics: moments and the integrated autocorrelation time WITHOUT a stored chain.

The reference keeps every iteration (`backend/backend.py:174-211`; 262 MB per iteration at c5) and estimates tau from the
stored chain (`autocorr.py:44-118`).  `ChainStats` owns the device accumulators that `hx_stats_feed` (csrc/hx_stats.cu)
updates once per iteration -- from inside the batch loops when attached to a sampler (`online_stats=`), or explicitly
through `feed(X)` -- and finishes the reference's estimator from them on the device:

    sum_t (x_t - m)(x_{t+k} - m) = A_k - m (head_k + tail_k) + (T - k) m^2,     k < max_lag
    A_k = sum_{t>=k} x_t x_{t-k},  head_k = total - (last k values),  tail_k = total - (first k values)

which is the quantity `function_1d` gets from its FFT (autocorr.py:30-31), so tau(M) for every window M < max_lag equals the
reference's number up to f64 rounding.  If the automatic window (first M >= c tau(M)) is not reached within max_lag lags the
estimate is refused (`AutocorrError`): raise `max_lag`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, _rt
from .autocorr import AutocorrError, auto_window

__all__ = ["ChainStats", "acf_from_accumulators"]


def acf_from_accumulators(count: int, lagsum, total, first, ring):
    """Normalised autocorrelation per tracked series from the accumulators (any device, f64).
    lagsum[K, S], total[S], first[K, S] = x_0..x_{K-1}, ring[K, S] = ring buffer of the last K values (slot t mod K).
    Returns acf[Kv, S] with Kv = min(K, count) -- lag k of autocorr.py:15-34 `function_1d` for each series."""
    K, S = lagsum.shape
    T = int(count)
    Kv = min(K, T)
    lagsum, total = lagsum.to(torch.float64), total.to(torch.float64)
    first, ring = first.to(torch.float64), ring.to(torch.float64)
    m = total / T
    # last values in reverse time order: x_{T-1}, x_{T-2}, ... live in ring slots (T-1) % K, (T-2) % K, ...
    slots = (T - 1 - torch.arange(Kv, device=lagsum.device)) % K
    last_rev = ring[slots]                                         # [Kv, S]: x_{T-1-i}
    zero = torch.zeros((1, S), dtype=torch.float64, device=lagsum.device)
    sum_last = torch.cat([zero, torch.cumsum(last_rev, 0)[:-1]])   # [Kv, S]: sum of the last k values
    sum_first = torch.cat([zero, torch.cumsum(first[:Kv], 0)[:-1]])
    head = total[None, :] - sum_last
    tail = total[None, :] - sum_first
    k = torch.arange(Kv, device=lagsum.device, dtype=torch.float64)[:, None]
    acov = lagsum[:Kv] - m[None, :] * (head + tail) + (T - k) * (m * m)[None, :]
    return acov / acov[0:1]


class ChainStats:
    """Device accumulators for one ensemble [nwalkers, ndim].

    walkers : how many walkers are tracked for the autocorrelation estimate (evenly strided rows; the reference averages
              the normalised ACF over ALL walkers -- a subset is a statistical, not an arithmetic, approximation; None = all)
    dims    : tracked parameters (sequence of indices, or None = all)
    max_lag : lags 0 .. max_lag-1 are accumulated"""

    def __init__(self, nwalkers, ndim, dtype=None, device=None, max_lag=128, walkers=None, dims=None):
        self.device = _rt.device(device)
        self.dtype = _rt.default_dtype() if dtype is None else dtype
        self.nwalkers, self.ndim, self.max_lag = int(nwalkers), int(ndim), int(max_lag)
        nt = self.nwalkers if walkers is None else max(0, min(int(walkers), self.nwalkers))
        self.stride = max(1, self.nwalkers // nt) if nt else 1
        self.n_track = nt
        self.dims = None if dims is None else torch.as_tensor(np.asarray(dims, dtype=np.int32), device=self.device)
        self.n_dims = self.ndim if dims is None else int(self.dims.numel())
        if self.dims is not None and (int(self.dims.min()) < 0 or int(self.dims.max()) >= self.ndim):
            raise ValueError("tracked dims out of range")
        self.reset()

    def reset(self):
        dev, S, K = self.device, self.n_track * self.n_dims, self.max_lag
        rb = int(_lib.load().hx_stats_row_blocks())
        f64 = dict(dtype=torch.float64, device=dev)
        self.count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sum, self.sumsq = torch.zeros(self.ndim, **f64), torch.zeros(self.ndim, **f64)
        self._partial = torch.zeros(2 * rb * self.ndim, **f64)
        self.lagsum = torch.zeros((K, S), **f64)
        self.ring = torch.zeros((K, S), dtype=self.dtype, device=dev)
        self.first = torch.zeros((K, S), dtype=self.dtype, device=dev)
        self.total = torch.zeros(S, **f64)
        st = _lib.hx_chain_stats()
        st.nwalkers, st.dim = self.nwalkers, self.ndim
        st.n_track, st.walker_stride, st.n_dims, st.max_lag = self.n_track, self.stride, self.n_dims, K
        st.dims_dev = None if self.dims is None else self.dims.data_ptr()
        st.count_dev, st.sum_dev, st.sumsq_dev = self.count.data_ptr(), self.sum.data_ptr(), self.sumsq.data_ptr()
        st.partial_dev, st.lagsum_dev = self._partial.data_ptr(), self.lagsum.data_ptr()
        st.ring_dev, st.first_dev, st.total_dev = self.ring.data_ptr(), self.first.data_ptr(), self.total.data_ptr()
        self.struct = st

    # -- feeding ------------------------------------------------------------------------------------------
    def feed(self, X: torch.Tensor):
        """Add one iteration: X[nwalkers, ndim] on this device, run dtype, contiguous."""
        if tuple(X.shape) != (self.nwalkers, self.ndim) or X.dtype != self.dtype or not X.is_contiguous():
            raise ValueError(f"feed expects a contiguous {self.dtype} tensor of shape {(self.nwalkers, self.ndim)}")
        _lib.check(_lib.load().hx_stats_feed(_rt.ctx(self.device), _rt.dtype_code(self.dtype), C.byref(self.struct), _rt.ptr(X),
                                             _rt.stream(self.device)), "hx_stats_feed")

    def attach(self):
        """The single-GPU batch loops of this device feed these accumulators after every iteration until detach()."""
        _lib.check(_lib.load().hx_ctx_set_chain_stats(_rt.ctx(self.device), C.byref(self.struct)), "hx_ctx_set_chain_stats")

    def detach(self):
        _lib.check(_lib.load().hx_ctx_set_chain_stats(_rt.ctx(self.device), None), "hx_ctx_set_chain_stats")

    # -- results ------------------------------------------------------------------------------------------
    @property
    def iterations(self) -> int:
        return int(self.count.item())

    def mean(self) -> np.ndarray:
        """Mean of every parameter over all walkers and iterations fed."""
        return (self.sum / (self.iterations * self.nwalkers)).cpu().numpy()

    def var(self) -> np.ndarray:
        """Variance (population form) of every parameter over all walkers and iterations fed."""
        N = self.iterations * self.nwalkers
        m = self.sum / N
        return (self.sumsq / N - m * m).cpu().numpy()

    def acf(self) -> torch.Tensor:
        """Walker-averaged normalised autocorrelation [lags, n_dims] (autocorr.py:98-100), on the device."""
        a = acf_from_accumulators(self.iterations, self.lagsum, self.total, self.first, self.ring)
        return a.reshape(a.shape[0], self.n_track, self.n_dims).mean(dim=1)

    def integrated_time(self, c=5, tol=50, quiet=False) -> np.ndarray:
        """tau per tracked parameter: `integrated_time` of autocorr.py:44-118 evaluated from the accumulators."""
        n_t = self.iterations
        if n_t < 1 or self.n_track * self.n_dims == 0:
            raise AttributeError("no iterations were fed / no series are tracked")
        taus = 2.0 * torch.cumsum(self.acf(), dim=0) - 1.0
        m = torch.arange(taus.shape[0], device=taus.device)[:, None] < c * taus
        reached = (~m).any(dim=0)
        if n_t > taus.shape[0] and not bool(reached.all()):
            t_last = taus[-1].cpu().numpy()
            raise AutocorrError(t_last, f"the automatic window (M >= {c} tau) was not reached within max_lag = {self.max_lag} lags for "
                                        f"{int((~reached).sum())} parameter(s): tau({self.max_lag - 1}) = {t_last}; raise max_lag")
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
