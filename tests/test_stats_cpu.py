"""The finalisation of the on-device chain statistics (hemcee_b200/stats.py acf_from_accumulators) against the oracle's FFT
estimator (oracle/autocorr.py, a restatement of reference src/hemcee/autocorr.py:15-118).  The accumulators are emulated in
NumPy exactly as csrc/hx_stats.cu builds them (lagged products, total, first K values, ring of the last K values)."""
import numpy as np
import pytest
import torch

from oracle import autocorr as OA


def _accumulate(x, K):
    """x[T, S] -> (lagsum[K, S], total[S], first[K, S], ring[K, S]) the way k_stats_lags does, one iteration at a time."""
    T, S = x.shape
    lagsum, total = np.zeros((K, S)), np.zeros(S)
    first, ring = np.zeros((K, S), x.dtype), np.zeros((K, S), x.dtype)
    for t in range(T):
        v = x[t]
        total += v
        if t < K:
            first[t] = v
        lagsum[0] += v.astype(np.float64) ** 2
        for k in range(1, min(t, K - 1) + 1):
            lagsum[k] += v.astype(np.float64) * ring[(t - k) % K]
        ring[t % K] = v
    return lagsum, total, first, ring


def _ar1(T, S, rho, seed, mean=0.0):
    rng = np.random.default_rng(seed)
    x = np.zeros((T, S))
    x[0] = rng.standard_normal(S)
    for t in range(1, T):
        x[t] = rho * x[t - 1] + np.sqrt(1 - rho * rho) * rng.standard_normal(S)
    return x + mean


@pytest.mark.parametrize("T,K,dtype", [(400, 64, np.float64), (50, 64, np.float64), (64, 64, np.float32), (1000, 17, np.float32)])
def test_acf_from_accumulators_equals_the_fft_estimator(T, K, dtype):
    from hemcee_b200.stats import acf_from_accumulators
    x = _ar1(T, 6, 0.7, 1, mean=3.0).astype(dtype)                     # a non-zero mean exercises the head / tail corrections
    acc = _accumulate(x, K)
    got = acf_from_accumulators(T, *[torch.from_numpy(a) for a in acc]).numpy()
    Kv = min(K, T)
    assert got.shape == (Kv, 6)
    for s in range(6):
        ref = OA.function_1d(x[:, s].astype(np.float64))[:Kv]
        np.testing.assert_allclose(got[:, s], ref, rtol=1e-9, atol=1e-10)


def test_windowed_tau_equals_integrated_time_when_the_window_fits():
    """tau from max_lag lags equals the reference's integrated_time whenever its automatic window M < max_lag."""
    from hemcee_b200.stats import acf_from_accumulators
    from hemcee_b200.autocorr import auto_window
    T, W, D, K = 3000, 12, 2, 96
    x = _ar1(T, W * D, 0.8, 7).reshape(T, W, D)                        # tau = (1 + rho) / (1 - rho) = 9, window ~ 45 < 96
    ref = OA.integrated_time(x, quiet=True)
    acc = _accumulate(x.reshape(T, W * D), K)
    a = acf_from_accumulators(T, *[torch.from_numpy(v) for v in acc]).reshape(K, W, D).mean(1)
    taus = 2.0 * torch.cumsum(a, 0) - 1.0
    win = auto_window(taus, 5)
    assert int(win.max()) < K - 1
    np.testing.assert_allclose(taus.gather(0, win[None])[0].numpy(), ref, rtol=1e-9)
