"""On-device chain statistics (csrc/hx_stats.cu, hemcee_b200/stats.py) through the C-ABI: the accumulators
fed per iteration reproduce the reference's estimators of the STORED chain -- moments, function_1d's normalised ACF and
integrated_time (reference src/hemcee/autocorr.py:15-118, oracle/autocorr.py) -- without the chain leaving the device."""
import numpy as np
import pytest
import torch

from oracle import autocorr as OA
from oracle import jaxrng as R

pytestmark = pytest.mark.gpu


def _ar1(T, W, D, rho, seed, mean):
    rng = np.random.default_rng(seed)
    x = np.zeros((T, W, D))
    x[0] = rng.standard_normal((W, D))
    for t in range(1, T):
        x[t] = rho * x[t - 1] + np.sqrt(1 - rho * rho) * rng.standard_normal((W, D))
    return x + mean


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("T,K", [(300, 40), (25, 40), (40, 40)])
def test_feed_matches_the_stored_chain_estimators(hx, dtype, T, K):
    from hemcee_b200.stats import ChainStats
    W, D = 48, 7
    npdt = np.float32 if dtype == torch.float32 else np.float64
    x = _ar1(T, W, D, 0.6, 3, mean=np.arange(D) - 2.0).astype(npdt)
    dims = [5, 0, 3]
    st = ChainStats(W, D, dtype, max_lag=K, walkers=16, dims=dims)         # walkers 0, 3, 6, ..., 45
    assert st.stride == 3 and st.n_track == 16
    X = torch.from_numpy(x).cuda()
    for t in range(T):
        st.feed(X[t])
    assert st.iterations == T
    x64 = x.astype(np.float64)
    np.testing.assert_allclose(st.mean(), x64.mean((0, 1)), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(st.var(), x64.var((0, 1)), rtol=1e-10)
    Kv = min(K, T)
    sub = x64[:, 0:48:3][:, :, dims]                                        # [T, 16, 3]
    ref = np.stack([np.mean([OA.function_1d(sub[:, w, i])[:Kv] for w in range(16)], axis=0) for i in range(3)], axis=1)
    np.testing.assert_allclose(st.acf().cpu().numpy(), ref, rtol=1e-9, atol=1e-10)
    if T == 300:
        np.testing.assert_allclose(st.integrated_time(quiet=True), OA.integrated_time(sub, quiet=True), rtol=1e-9)


def test_window_beyond_max_lag_is_refused(hx):
    from hemcee_b200.stats import ChainStats
    x = _ar1(600, 8, 1, 0.97, 5, 0.0).astype(np.float32)                    # tau ~ 65: the window needs ~300 lags
    st = ChainStats(8, 1, torch.float32, max_lag=16)
    X = torch.from_numpy(x).cuda()
    for t in range(600):
        st.feed(X[t])
    with pytest.raises(hx.autocorr.AutocorrError, match="max_lag"):
        st.integrated_time(quiet=True)


@pytest.mark.parametrize("x64", [False, True])
def test_sampler_online_stats_equal_the_stored_chain(hx, x64):
    """run_mcmc with online_stats: tau, mean and variance from the accumulators equal those of the chain the same run stored."""
    hx.config.update("enable_x64", x64)
    try:
        dt = np.float64 if x64 else np.float32
        W, d = 64, 4
        keys = R.split(R.PRNGKey(4), 2)
        x0 = R.normal(keys[0], (W, d), dt)
        s = hx.HamiltonianEnsembleSampler(W, d, hx.targets.GaussianIso(d), step_size=0.4, L=3, adapt_inital_step_size=False,
                                          adapt_step_size=False, adapt_length=False, verbose=False,
                                          online_stats=dict(max_lag=64, walkers=None, dims=None))
        chain = s.run_mcmc(keys[1], x0, 1500, warmup=50, thin_by=1, batch_size=256)
        assert s.stats.iterations == 1500
        c64 = np.asarray(chain, dtype=np.float64)
        np.testing.assert_allclose(s.stats.mean(), c64.mean((0, 1)), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(s.stats.var(), c64.var((0, 1)), rtol=1e-9)
        np.testing.assert_allclose(s.stats.integrated_time(), hx.autocorr.integrated_time(chain), rtol=1e-8)
        np.testing.assert_allclose(s.stats.integrated_time(), OA.integrated_time(c64), rtol=1e-8)
    finally:
        hx.config.update("enable_x64", False)


def test_store_none_keeps_no_chain_and_the_same_final_state(hx):
    W, d = 128, 6
    keys = R.split(R.PRNGKey(8), 2)
    x0 = R.normal(keys[0], (W, d), np.float32)
    kw = dict(step_size=0.3, L=4, adapt_inital_step_size=False, adapt_step_size=False, adapt_length=False, verbose=False)
    full = hx.HamiltonianEnsembleSampler(W, d, hx.targets.GaussianIso(d), online_stats=True, **kw)
    chain = full.run_mcmc(keys[1], x0, 800, warmup=100)
    none = hx.HamiltonianEnsembleSampler(W, d, hx.targets.GaussianIso(d), store="none", online_stats=True, **kw)
    out = none.run_mcmc(keys[1], x0, 800, warmup=100)
    assert out.shape == (0, W, d) and sum(len(c) for c in none.backend.chain) == 0
    xl, lpl = none.backend.get_last_sample()
    np.testing.assert_array_equal(xl, chain[-1])                            # same draws, same chain: only the D2H is gone
    np.testing.assert_array_equal(none.stats.integrated_time(), full.stats.integrated_time())
    np.testing.assert_allclose(none.backend.get_autocorr_time(), hx.autocorr.integrated_time(chain), rtol=1e-8)
    np.testing.assert_array_equal(none.backend.accepted, full.backend.accepted)


def test_online_stats_on_the_tensor_core_path(hx):
    """The tcgen05 batch loop feeds the accumulators too (c5-shaped problem at test size)."""
    from oracle import targets as OT
    n, d = 256, 128
    ks = R.split(R.PRNGKey(1), 4)
    cov = OT.make_covariance_skewed(ks[0], d, 100.0, np.float64)
    x0 = (R.normal(ks[1], (2 * n, d), np.float32) * 0.5).astype(np.float32)
    hx.config.update("precision", "bf16x3")
    try:
        s = hx.HamiltonianEnsembleSampler(2 * n, d, hx.targets.GaussianFull(cov=cov), step_size=0.03, L=4, verbose=False,
                                          adapt_inital_step_size=False, adapt_step_size=False, adapt_length=False,
                                          online_stats=dict(max_lag=32, walkers=64, dims=[0, 17, 127]))
        chain = s.run_mcmc(ks[2], x0, 60, warmup=0)
    finally:
        hx.config.update("precision", "auto")
    sub = np.asarray(chain, np.float64)[:, 0:512:8][:, :, [0, 17, 127]]
    ref = np.stack([np.mean([OA.function_1d(sub[:, w, i])[:32] for w in range(64)], axis=0) for i in range(3)], axis=1)
    np.testing.assert_allclose(s.stats.acf().cpu().numpy(), ref, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(s.stats.mean(), np.asarray(chain, np.float64).mean((0, 1)), rtol=1e-9, atol=1e-12)
