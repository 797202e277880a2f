"""Device-resident adaptation on the tensor-core path (hx_run_warmup): dual averaging and ChEES with the
adapter state and its update kernels on the device for fp32 full-covariance Gaussians -- the c5 family -- including ChEES
beyond dim 128, where cov(group2)^-1 comes from an fp64 blocked Cholesky (csrc/hx_spd.cuh) and the inner product of all
walkers from one tensor-core GEMM (reference adaptation/chees.py:57-73,89-166, dual_averaging.py:64-89, adapter.py:58-78)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import jaxrng as R, targets as OT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d", [1, 5, 63, 64, 65, 130, 500, 1000])
def test_spd_inverse_matches_numpy(hx, d):
    from hemcee_b200 import _lib, _rt
    rng = np.random.default_rng(d)
    n = 4 * d + 8
    c = rng.standard_normal((n, d)) * np.sqrt(np.geomspace(1.0, 1e4, d))[None, :]          # kappa ~ 1e4 like the c5 complement
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    m32 = ((c @ q).T @ (c @ q) / n).astype(np.float32)
    m32 = 0.5 * (m32 + m32.T)
    dev = torch.device("cuda", 0)
    M = torch.from_numpy(m32).to(dev)
    out = torch.empty((d, d), dtype=torch.float32, device=dev)
    scale = n / (n - 1.0)
    _lib.check(_lib.load().hx_spd_inverse(_rt.ctx(dev), _rt.ptr(M), d, d, C.c_double(scale), _rt.ptr(out), d, _rt.stream(dev)),
               "hx_spd_inverse")
    ref = np.linalg.inv(scale * m32.astype(np.float64))
    got = out.cpu().numpy().astype(np.float64)
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()         # fp64 factorisation, result rounded to fp32
    np.testing.assert_allclose(got, got.T, rtol=0, atol=1e-7 * np.abs(ref).max())


def _setup(hx, W, d, kappa=100.0, seed=3):
    ks = R.split(R.PRNGKey(seed), 3)
    cov = OT.make_covariance_skewed(ks[0], d, kappa, np.float64)
    x0 = (R.normal(ks[1], (W, d), np.float32).astype(np.float64) @ np.linalg.cholesky(cov).T).astype(np.float32)
    return hx.targets.GaussianFull(cov=cov), x0, ks[2]


@pytest.mark.parametrize("W,d,adapt", [(512, 32, (True, False)), (512, 32, (True, True)), (512, 160, (True, True)),
                                       (512, 160, (False, True)), (1024, 300, (True, True))])
def test_tensor_core_warmup_device_adapter_tracks_the_host_loop(hx, W, d, adapt):
    """The same warm-up through hx_run_warmup (adapter on the device, (step, L) read back per half-move) and through the
    Python host loop (torch adapters: cov + LU solve like the reference).  Both run the same tensor-core moves; the adapter
    arithmetic differs in rounding (fp64 Cholesky + bf16 hi/lo GEMM against torch's fp32 solve), so the trajectories agree
    until an integer L flips: after 4 iterations the dual-averaging state agrees to 2e-3 and ChEES' log T to 2e-2."""
    tgt, x0, key = _setup(hx, W, d)
    hx.config.update("precision", "bf16x3")
    st = {}
    try:
        for dev_adapt in (True, False):
            hx.config.update("device_adaptation", dev_adapt)
            s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=0.1, L=4, adapt_inital_step_size=False,
                                              adapt_step_size=adapt[0], adapt_length=adapt[1], verbose=False)
            assert s._device_adapt_ok() == dev_adapt
            chain = s.run_mcmc(key, x0, 2, warmup=4)
            assert np.all(np.isfinite(chain))
            st[dev_adapt] = (s.adapter_state, s.found, s.diagnostics_warmup["accepts"])
    finally:
        hx.config.update("device_adaptation", True)
        hx.config.update("precision", "auto")
    a, b = st[True][0], st[False][0]
    da_a = a if adapt == (True, False) else getattr(a, "da_state", None)
    da_b = b if adapt == (True, False) else getattr(b, "da_state", None)
    ch_a = a if adapt == (False, True) else getattr(a, "chees_state", None)
    ch_b = b if adapt == (False, True) else getattr(b, "chees_state", None)
    if adapt[0]:
        assert abs(float(da_a.log_stepsize) - float(da_b.log_stepsize)) < 2e-3, (da_a, da_b)
        assert abs(float(da_a.log_stepsize_bar) - float(da_b.log_stepsize_bar)) < 2e-3
    if adapt[1]:
        assert abs(float(ch_a.log_T) - float(ch_b.log_T)) < 2e-2, (ch_a, ch_b)
        assert float(ch_a.iteration) == float(ch_b.iteration) == 9.0
    assert abs(int(st[True][2].sum()) - int(st[False][2].sum())) <= 0.02 * W * 4


def test_tensor_core_warmup_at_scale_runs_on_the_device(hx):
    """c5-shaped (auto precision: d >= 128, >= 2048 walkers per group): DA + ChEES warm-up through hx_run_warmup on the
    tcgen05 path, no chain stored; the tuned values are sane and the main phase continues from it."""
    W, d = 4096, 256
    tgt, x0, key = _setup(hx, W, d, kappa=1e3)
    s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=0.1, L=5, adapt_inital_step_size=False, verbose=False, store="none",
                                      online_stats=True)
    assert s._device_adapt_ok()
    s.run_mcmc(key, x0, 20, warmup=40)
    from hemcee_b200 import _lib, _rt
    assert _lib.load().hx_ctx_last_path(_rt.ctx(torch.device("cuda", torch.cuda.current_device()))) == 1      # tcgen05 path ran
    step, L = float(s.found[0]), int(s.found[1])
    assert 1e-3 < step < 2.0 and 1 <= L <= 5000, (step, L)
    acc = s.diagnostics_warmup["acceptance_rate"].mean()
    assert 0.2 < acc <= 1.0, acc
    assert s.stats.iterations == 20
