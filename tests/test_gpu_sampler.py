"""End-to-end parity of the drop-in sampler (run_mcmc through the C-ABI) against the oracle sampler,
plus the reference's own behavioural tests re-targeted at the drop-in API:
  tests/test_proposal.py:27-88 (N(0,1) statistics), tests/test_backend.py:29-37 (shapes, counters),
  tests/test_adapter.py:43-121 (adapter dispatch), tests/test_log_prob_validation.py (NaN / -inf).
"""
import numpy as np
import pytest
import torch
from scipy.stats import kstest

from oracle import jaxrng as R, moves as OM, targets as OT, sampler as OS

pytestmark = pytest.mark.gpu


@pytest.fixture()
def x64(hx):
    hx.config.update("enable_x64", True)
    yield hx
    hx.config.update("enable_x64", False)


def _pair(hx, kind, d):
    if kind == "iso":
        return hx.targets.GaussianIso(d), OT.make_iso_gaussian()
    if kind == "rosen":
        return hx.targets.Rosenbrock(d), OT.make_rosenbrock()
    if kind == "funnel":
        return hx.targets.Funnel(d), OT.make_funnel()
    ks = R.split(R.PRNGKey(1), 2)
    cov = OT.make_covariance_skewed(ks[0], d, 1000.0, np.float64)
    return hx.targets.GaussianFull(cov=cov), OT.make_gaussian(cov, np.zeros(d))


@pytest.mark.parametrize("move", ["walk", "side"])
@pytest.mark.parametrize("kind,W,d", [("gauss", 32, 10), ("iso", 20, 3), ("rosen", 24, 2), ("funnel", 16, 5)])
def test_main_phase_parity_x64(x64, move, kind, W, d):
    """c1-sized runs in x64: accept decisions identical, chain within 1e-9 over T iterations."""
    hx = x64
    T = 25
    tgt, otgt = _pair(hx, kind, d)
    key = R.PRNGKey(0)
    k_init, k_run = R.split(key, 2)
    x0 = R.normal(k_init, (W, d), np.float64) * 0.5
    eps = 0.05 if kind != "rosen" else 0.003
    pm = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    om = OM.hmc_walk_move if move == "walk" else OM.hmc_side_move
    s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=eps, L=6, move=pm, adapt_inital_step_size=False,
                                      adapt_step_size=False, adapt_length=False, verbose=False)
    chain = s.run_mcmc(k_run, x0, T, warmup=0, batch_size=7)
    o = OS.HamiltonianEnsembleSampler(W, d, otgt, step_size=eps, L=6, move=om, adapt_inital_step_size=False,
                                      adapt_step_size=False, adapt_length=False, dtype=np.float64)
    ref = o.run_mcmc(k_run, x0, T, warmup=0, batch_size=7)
    assert chain.shape == (T, W, d)
    np.testing.assert_array_equal(s.diagnostics_main["accepts"], o.diagnostics_main["accepts"])
    np.testing.assert_allclose(chain, ref, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(s.get_logprob(), o.get_logprob(), rtol=1e-9, atol=1e-9)
    assert s.backend.iteration == T


@pytest.mark.parametrize("move", ["walk", "side"])
def test_main_phase_parity_f32(hx, move):
    """f32: first iterations agree with the f32 oracle within 1e-3 of the ensemble scale; accept masks
    may differ only where |log u - log_acc| is inside the f32 dH noise (checked via the accept counts)."""
    W, d, T = 64, 10, 6
    tgt, otgt = _pair(hx, "gauss", d)
    k_init, k_run = R.split(R.PRNGKey(3), 2)
    x0 = R.normal(k_init, (W, d), np.float32)
    pm = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    om = OM.hmc_walk_move if move == "walk" else OM.hmc_side_move
    s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=0.03, L=8, move=pm, adapt_inital_step_size=False,
                                      adapt_step_size=False, adapt_length=False, verbose=False)
    chain = s.run_mcmc(k_run, x0, T, warmup=0)
    o = OS.HamiltonianEnsembleSampler(W, d, otgt, step_size=0.03, L=8, move=om, adapt_inital_step_size=False,
                                      adapt_step_size=False, adapt_length=False, dtype=np.float32)
    ref = o.run_mcmc(k_run, x0, T, warmup=0)
    same = np.all(np.abs(chain - ref) <= 1e-3 * np.abs(ref).max(), axis=2)      # per (t, walker)
    assert same.mean() > 0.97, same.mean()
    assert abs(int(s.diagnostics_main["accepts"].sum()) - int(o.diagnostics_main["accepts"].sum())) <= 0.03 * W * T


@pytest.mark.parametrize("move", ["walk", "side"])
def test_warmup_adaptation_parity_x64(x64, move):
    """Warm-up with DA + ChEES: same key order, adapter trajectory and final (step, L) as the oracle."""
    hx = x64
    W, d = 20, 3
    tgt, otgt = _pair(hx, "iso", d)
    keys = R.split(R.PRNGKey(2), 2)
    x0 = R.normal(keys[0], (W, d), np.float64)
    pm = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    om = OM.hmc_walk_move if move == "walk" else OM.hmc_side_move
    s = hx.HamiltonianEnsembleSampler(W, d, tgt, move=pm, verbose=False)
    chain = s.run_mcmc(keys[1], x0, 10, warmup=40)
    o = OS.HamiltonianEnsembleSampler(W, d, otgt, move=om, dtype=np.float64)
    ref = o.run_mcmc(keys[1], x0, 10, warmup=40)
    assert float(s.step_size) == float(o.step_size)                   # initial step-size heuristic
    assert int(s.found[1]) == int(o.found[1])
    np.testing.assert_allclose(float(s.found[0]), float(o.found[0]), rtol=1e-10)
    np.testing.assert_allclose(chain, ref, rtol=1e-8, atol=1e-9)
    assert s.backend.iteration == 50 and chain.shape == (10, W, d)


@pytest.mark.parametrize("move", ["walk", "side"])
@pytest.mark.parametrize("adapt", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("target", ["iso", "rosenbrock"])
def test_device_adaptation_equals_host_loop_x64(x64, move, adapt, target):
    """hx_run_warmup (adapter state, value() and update() on the device) against the host-driven warm-up loop: same adapter
    trajectory (step to 1e-10, same L) and chain, for DA, ChEES and both, walk and side moves."""
    hx = x64
    W, d = 24, 4
    tgt = hx.targets.GaussianIso(d) if target == "iso" else hx.targets.Rosenbrock(d)
    keys = R.split(R.PRNGKey(4), 2)
    x0 = R.normal(keys[0], (W, d), np.float64) * (1.0 if target == "iso" else 0.3) + (0.0 if target == "iso" else 1.0)
    pm = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    out = {}
    try:
        for dev_adapt in (True, False):
            hx.config.update("device_adaptation", dev_adapt)
            s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=0.05, L=5, move=pm, adapt_inital_step_size=False,
                                              adapt_step_size=adapt[0], adapt_length=adapt[1], verbose=False)
            chain = s.run_mcmc(keys[1], x0, 6, warmup=30, batch_size=7)
            out[dev_adapt] = (chain, s.found, s.get_chain(), s.diagnostics_warmup["accepts"])
    finally:
        hx.config.update("device_adaptation", True)
    assert int(out[True][1][1]) == int(out[False][1][1])
    np.testing.assert_allclose(float(out[True][1][0]), float(out[False][1][0]), rtol=1e-10)
    np.testing.assert_allclose(out[True][2], out[False][2], rtol=1e-8, atol=1e-9)        # warm-up + main chain
    np.testing.assert_array_equal(out[True][3], out[False][3])


def test_device_adaptation_f32_runs(hx):
    """f32 warm-up on the device: finite chain, sane tuned values (draw-level agreement with the host loop is not expected in
    f32: exp/log/pow differ in the last bit between NumPy and CUDA and one flipped accept decorrelates the trajectories)."""
    W, d = 64, 8
    tgt = hx.targets.Rosenbrock(d)
    keys = R.split(R.PRNGKey(6), 2)
    x0 = (R.normal(keys[0], (W, d), np.float32) * 0.3 + 1.0).astype(np.float32)
    s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=0.01, L=5, adapt_inital_step_size=False, verbose=False)
    chain = s.run_mcmc(keys[1], x0, 20, warmup=100)
    assert np.all(np.isfinite(chain)) and chain.shape == (20, W, d)
    assert 1e-4 < float(s.found[0]) < 1.0 and 1 <= int(s.found[1]) <= 10000
    assert s.backend.iteration == 120


@pytest.mark.parametrize("move", ["walk", "side"])
@pytest.mark.parametrize("warm,adapt", [(0, False), (5, False), (6, True)])
def test_thinned_store_equals_full_store(hx, move, warm, adapt):
    """store='thinned' (hx_ctx_set_chain_thinning: only the iterations run_mcmc returns leave the device) returns bit for bit
    what the reference-style full store + get_chain(discard=warmup, thin=thin_by) returns."""
    W, d = 32, 6
    tgt = hx.targets.Rosenbrock(d)
    keys = R.split(R.PRNGKey(9), 2)
    x0 = (R.normal(keys[0], (W, d), np.float32) * 0.3 + 1.0).astype(np.float32)
    pm = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    out = {}
    for store in ("full", "thinned"):
        s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=0.01, L=4, move=pm, adapt_inital_step_size=False,
                                          adapt_step_size=adapt, adapt_length=adapt, store=store, verbose=False)
        out[store] = s.run_mcmc(keys[1], x0, 7, warmup=warm, thin_by=3, batch_size=4)
        assert s.backend.iteration == warm + 21
    assert out["thinned"].shape == (7, W, d)
    np.testing.assert_array_equal(out["thinned"], out["full"])


def test_quickstart_initial_step(x64):
    """docs/tutorials/quickstart.ipynb cell 8 prints 'Using inital step size of 1.6'."""
    hx = x64
    keys = R.split(R.PRNGKey(2), 2)
    x0 = R.normal(keys[0], (20, 3), np.float64)
    s = hx.HamiltonianEnsembleSampler(20, 3, hx.targets.GaussianIso(3), verbose=False)
    s.run_mcmc(keys[1], x0, 2, warmup=0)
    assert float(s.step_size) == 1.6


def test_adaptation_notebook_initial_step(hx):
    """docs/tutorials/adaptation.ipynb cell 5 prints 'Using inital step size of 0.004999999888241291' (f32)."""
    keys = R.split(R.PRNGKey(0), 2)
    x0 = R.normal(keys[0], (100, 2), np.float32)
    s = hx.HamiltonianEnsembleSampler(100, 2, hx.targets.Rosenbrock(2), step_size=0.01, verbose=False)
    s.run_mcmc(keys[1], x0, 2, warmup=0)
    assert float(s.step_size) == 0.004999999888241291


@pytest.mark.parametrize("move", ["walk", "side"])
def test_normal(x64, move, ndim=1, nwalkers=20, nsteps=3000, seed=1):
    """reference tests/test_proposal.py:27-88 (same sizes, seed and thresholds)."""
    hx = x64
    keys = R.split(R.PRNGKey(seed), 2)
    init = R.normal(keys[0], (nwalkers, ndim), np.float64)
    pm = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    sampler = hx.HamiltonianEnsembleSampler(total_chains=nwalkers, dim=ndim, log_prob=hx.targets.GaussianIso(ndim),
                                            move=pm, verbose=False)
    samples = sampler.run_mcmc(keys[1], init, nsteps, warmup=0, thin_by=1)
    assert samples.shape == (nsteps, nwalkers, ndim)
    samples = samples.reshape(-1, ndim)
    assert np.all(np.abs(np.mean(samples, axis=0)) < 0.08), "Incorrect mean"
    assert np.all(np.abs(np.std(samples, axis=0) - 1) < 0.05), "Incorrect standard deviation"
    ks_stat, _ = kstest(samples.reshape(-1), "norm")
    assert ks_stat < 0.05, "The K-S test failed"


def test_iterations_backend(hx):
    """reference tests/test_backend.py:29-37."""
    keys = R.split(R.PRNGKey(0), 2)
    x0 = R.normal(keys[0], (4, 1), np.float32)
    sampler = hx.HamiltonianEnsembleSampler(4, 1, hx.targets.GaussianIso(1), backend=hx.Backend(), verbose=False)
    samples = sampler.run_mcmc(keys[0], x0, num_samples=10, warmup=35, thin_by=2)
    assert samples.shape == (10, 4, 1)
    assert sampler.backend.iteration == 35 + 10 * 2
    assert sampler.get_acceptance_prob().shape == (4,)


@pytest.mark.parametrize("move", ["walk", "side"])
@pytest.mark.parametrize("da,ch,cls", [(True, True, "CompositeAdapter"), (True, False, "DualAveragingAdapter"),
                                       (False, True, "ChEESAdapter"), (False, False, "NoOpAdapter")])
def test_adapter_with_move(hx, move, da, ch, cls):
    """reference tests/test_adapter.py:79-121 (W=4, d=1, warm-up 5)."""
    from hemcee_b200 import adaptation
    keys = R.split(R.PRNGKey(0), 30)
    x0 = R.normal(keys[0], (4, 1), np.float32)
    pm = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    sampler = hx.HamiltonianEnsembleSampler(4, 1, hx.targets.GaussianIso(1), move=pm, adapt_step_size=da,
                                            adapt_length=ch, verbose=False)
    sampler.run_mcmc(keys[3], x0, num_samples=1, warmup=5)
    assert isinstance(sampler.adapter, getattr(adaptation, cls))


def test_nan_log_prob_raises(hx):
    """reference tests/test_log_prob_validation.py:28-89: a NaN stored log-prob raises ValueError."""
    x0 = np.array([[np.nan, 0.5], [0.5, 0.5], [-0.5, 0.5], [0.5, -0.5]], dtype=np.float32)
    sampler = hx.HamiltonianEnsembleSampler(4, 2, hx.targets.GaussianIso(2), adapt_inital_step_size=False, verbose=False)
    with pytest.raises(ValueError, match="Log probability returned NaN"):
        sampler.run_mcmc(R.PRNGKey(1), x0, num_samples=5, warmup=0)


def test_neginf_is_rejected(hx):
    """reference tests/test_log_prob_validation.py:95-155: -inf regions are rejected silently."""
    key = R.PRNGKey(4)
    x0 = np.abs(R.normal(key, (4, 2), np.float32))
    tgt = hx.targets.GaussianIso(2, lower0=0.0)
    sampler = hx.HamiltonianEnsembleSampler(4, 2, tgt, verbose=False)
    samples = sampler.run_mcmc(key, x0, num_samples=5, warmup=10)
    assert np.all(samples[:, :, 0] >= 0)


def test_bad_arguments(hx):
    tgt = hx.targets.GaussianIso(2)
    with pytest.raises(ValueError):
        hx.HamiltonianEnsembleSampler(2, 2, tgt)
    s = hx.HamiltonianEnsembleSampler(4, 2, tgt, verbose=False)
    with pytest.raises(ValueError):
        s.run_mcmc(R.PRNGKey(0), np.zeros((4, 3), np.float32), 2, warmup=0)
    with pytest.raises(ValueError):
        s.run_mcmc(R.PRNGKey(0), np.zeros((4, 2), np.float32), 2, warmup=0, thin_by=0)
    with pytest.raises(hx.HxError):
        hx.hmc_walk_move(torch.zeros(4, 2).cuda(), torch.zeros(4, 2).cuda(), 0.1, R.PRNGKey(0), tgt, tgt, 0,
                         torch.zeros(4).cuda())
