"""hemcee_b200.autocorr.integrated_time (batched rFFT over walkers x parameters) against the reference's own unit
tests (src/hemcee/tests/test_autocorr.py: AR(1) chains with tau ~ 19 +- 20 %, walkers / no-walkers equivalence,
AutocorrError on a short chain, multi-parameter = per-parameter) and against the NumPy oracle restatement of
src/hemcee/autocorr.py:15-118 on seeded ensembles (rtol 1e-9: same estimator, rfft instead of fft)."""
import numpy as np
import pytest
import torch

from hemcee_b200.autocorr import AutocorrError, function_1d, integrated_time
from oracle import autocorr as OA


def get_chain(seed=1234, ndim=3, N=100000):
    """The reference test's AR(1) series, vectorised (x_i = 0.9 x_{i-1} + U(0,1))."""
    rng = np.random.RandomState(seed)
    e = rng.rand(N, ndim)
    x = np.empty((N, ndim))
    x[0] = 0.0
    for i in range(1, N):
        x[i] = x[i - 1] * 0.9 + e[i]
    return x


def test_1d():
    tau = integrated_time(get_chain(ndim=1, N=250000))
    assert np.all(np.abs(tau - 19.0) / 19.0 < 0.2)


def test_nd():
    tau = integrated_time(get_chain(ndim=3, N=150000))
    assert np.all(np.abs(tau - 19.0) / 19.0 < 0.2)


def test_nd_without_walkers():
    x = get_chain(ndim=3, N=10000)
    assert np.allclose(integrated_time(x[:, np.newaxis]), integrated_time(x, has_walkers=False))


def test_too_short():
    x = get_chain(ndim=3, N=100)
    with pytest.raises(AutocorrError) as ei:
        integrated_time(x)
    assert ei.value.tau.shape == (1,)      # (n_step, n_walker) by default: one parameter
    tau = integrated_time(x, quiet=True)
    np.testing.assert_array_equal(tau, ei.value.tau)


def test_autocorr_multi_works():
    xs = np.random.RandomState(42).randn(16384, 2)
    multi = integrated_time(xs[:, np.newaxis])
    single = np.array([integrated_time(xs[:, i]) for i in range(2)]).squeeze()
    assert np.allclose(multi, single)


def test_invalid_dimensions():
    with pytest.raises(ValueError):
        integrated_time(np.zeros((4, 3, 2, 2)))
    with pytest.raises(ValueError):
        function_1d(np.zeros((4, 3)))


@pytest.mark.parametrize("T,W,D,seed", [(2000, 8, 3, 0), (777, 5, 4, 1), (4096, 16, 2, 2), (130, 3, 1, 3)])
def test_matches_oracle(T, W, D, seed):
    rng = np.random.RandomState(seed)
    rho = rng.uniform(0.3, 0.95, size=(1, W, D))
    x = np.empty((T, W, D))
    x[0] = rng.randn(W, D)
    for t in range(1, T):
        x[t] = rho[0] * x[t - 1] + rng.randn(W, D)
    got = integrated_time(x, quiet=True)
    ref = OA.integrated_time(x, quiet=True)
    np.testing.assert_allclose(got, ref, rtol=1e-9)
    np.testing.assert_allclose(function_1d(x[:, 0, 0]).numpy(), OA.function_1d(x[:, 0, 0]), rtol=1e-9, atol=1e-12)
    # window factor and thinning conventions
    np.testing.assert_allclose(integrated_time(x, c=7, quiet=True), OA.integrated_time(x, c=7, quiet=True), rtol=1e-9)
    np.testing.assert_allclose(integrated_time(x[::3], quiet=True), OA.integrated_time(x[::3], quiet=True), rtol=1e-9)


def test_column_chunking_is_transparent(monkeypatch):
    """The FFT batch is cut into column chunks to bound memory; any chunk size gives the same tau."""
    import hemcee_b200.autocorr as A
    x = np.random.RandomState(5).randn(512, 4, 9).cumsum(axis=0) * 0.05 + np.random.RandomState(6).randn(512, 4, 9)
    full = A.integrated_time(x, quiet=True)
    monkeypatch.setattr(A, "_CHUNK_ELEMS", 2 * 1024 * 4 * 2)      # two parameters per chunk
    np.testing.assert_allclose(A.integrated_time(x, quiet=True), full, rtol=1e-12)


@pytest.mark.gpu
def test_matches_oracle_on_the_device():
    """The ESS leg of bench.py keeps its chain subset on the GPU: cuFFT path against the oracle."""
    assert torch.cuda.is_available()
    rng = np.random.RandomState(11)
    T, W, D = 1500, 64, 16
    x = np.empty((T, W, D), dtype=np.float32)
    x[0] = rng.randn(W, D)
    for t in range(1, T):
        x[t] = 0.8 * x[t - 1] + rng.randn(W, D)
    got = integrated_time(torch.from_numpy(x).cuda(), quiet=True)
    ref = OA.integrated_time(x.astype(np.float64), quiet=True)
    np.testing.assert_allclose(got, ref, rtol=1e-8)
