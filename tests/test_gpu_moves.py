"""GPU parity of the fused half-move kernels (through the C-ABI) against the dense-form oracle.

Tolerances (stated, per dtype):
  f64: positions / log-probs / log-accept rtol 1e-9 (the oracle integrates the dense n x n momentum,
       the kernel the projected form; they differ only by summation order)
  f32: rtol 2e-3 on positions relative to the trajectory scale, atol 5e-3 on log-accept
"""
import numpy as np
import pytest
import torch

from oracle import jaxrng as R, moves as OM, targets as OT

pytestmark = pytest.mark.gpu


def make_target(hx, kind, d, dt, seed=3):
    """-> (product target, oracle (log_prob, grad))"""
    if kind == "iso":
        return hx.targets.GaussianIso(d), OT.make_iso_gaussian()
    if kind == "gauss":
        key = R.PRNGKey(seed)
        ks = R.split(key, 2)
        cov = OT.make_covariance_skewed(ks[0], d, 100.0, np.float64)
        mean = R.normal(ks[1], (d,), np.float64)
        return hx.targets.GaussianFull(cov=cov, mean=mean), OT.make_gaussian(cov, mean)
    if kind == "rosen":
        return hx.targets.Rosenbrock(d), OT.make_rosenbrock()
    if kind == "funnel":
        return hx.targets.Funnel(d), OT.make_funnel()
    raise ValueError(kind)


def setup(n, d, dt, seed=0):
    ks = R.split(R.PRNGKey(seed), 3)
    g1 = R.normal(ks[0], (n, d), dt) * dt(0.7)
    g2 = R.normal(ks[1], (n, d), dt) * dt(0.9) + dt(0.1)
    return g1.astype(dt), g2.astype(dt), ks[2]


CASES = [("iso", 16, 3), ("gauss", 16, 10), ("gauss", 40, 5), ("rosen", 24, 2), ("rosen", 20, 7),
         ("funnel", 32, 6), ("gauss", 64, 33), ("gauss", 24, 130), ("iso", 300, 4)]


@pytest.mark.parametrize("move", ["walk", "side"])
@pytest.mark.parametrize("kind,n,d", CASES)
@pytest.mark.parametrize("tdt", [torch.float64, torch.float32])
def test_move_matches_oracle(hx, move, kind, n, d, tdt):
    dt = np.float64 if tdt == torch.float64 else np.float32
    tgt, (olp, ograd) = make_target(hx, kind, d, dt)
    g1, g2, key = setup(n, d, dt)
    eps, L = (0.05, 7) if kind != "rosen" else (0.004, 5)
    omove = OM.hmc_walk_move if move == "walk" else OM.hmc_side_move
    pmove = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    lp1 = olp(g1)
    q_ref, la_ref, pp_ref, lp_ref = omove(g1, g2, eps, key, olp, ograd, L, lp1)
    G1, G2 = torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda()
    LP1 = tgt.log_prob(G1)
    q, la, pp, lp = pmove(G1, G2, eps, key, tgt, tgt, L, LP1)
    q, la, pp, lp, LP1 = (t.cpu().numpy() for t in (q, la, pp, lp, LP1))
    if dt == np.float64:
        np.testing.assert_allclose(LP1, lp1, rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(q, q_ref, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(pp, pp_ref, rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(lp, lp_ref, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(la, la_ref, rtol=1e-7, atol=1e-8)
    else:
        scale = np.abs(q_ref).max()
        np.testing.assert_allclose(q, q_ref, rtol=0, atol=2e-3 * scale)
        np.testing.assert_allclose(lp, lp_ref, rtol=2e-3, atol=2e-2)
        # The f32 dense oracle carries K0, K1 ~ n/2 cancellation noise in dH, so log-accept is
        # checked against the SAME f32 random stream integrated in f64 arithmetic.
        _, (olp64, og64) = make_target(hx, kind, d, np.float64)
        a64, b64 = g1.astype(np.float64), g2.astype(np.float64)
        q64, la64, _, lp64 = omove(a64, b64, eps, key, olp64, og64, L, olp64(a64), rng_dtype=np.float32)
        np.testing.assert_allclose(q, q64, rtol=0, atol=5e-4 * scale)
        np.testing.assert_allclose(la, la64, rtol=2e-3, atol=2e-3 * max(1.0, np.abs(lp64).max() * 1e-2))


@pytest.mark.parametrize("move", ["walk", "side"])
def test_frozen_gradient_L1(hx, move):
    """reasonable_stepsize.py:34-59: L=1 with the gradient frozen at the initial positions."""
    dt = np.float64
    tgt, (olp, ograd) = make_target(hx, "rosen", 4, dt)
    g1, g2, key = setup(20, 4, dt, seed=5)
    g0 = ograd(g1)
    omove = OM.hmc_walk_move if move == "walk" else OM.hmc_side_move
    pmove = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    q_ref, la_ref, _, lp_ref = omove(g1, g2, 0.01, key, olp, lambda x: g0, 1, olp(g1))
    G1, G2 = torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda()
    q, la, _, lp = pmove(G1, G2, 0.01, key, tgt, hx.targets.FrozenGrad(tgt), L=1, log_prob_group1=tgt.log_prob(G1))
    np.testing.assert_allclose(q.cpu().numpy(), q_ref, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(lp.cpu().numpy(), lp_ref, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(la.cpu().numpy(), la_ref, rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("move", ["walk", "side"])
@pytest.mark.parametrize("impl", ["partitionable", "legacy"])
def test_shard_invariance(hx, move, impl):
    """Updating row shards separately gives bit-identical results to the single call."""
    dt = np.float32
    tgt, _ = make_target(hx, "gauss", 12, dt)
    g1, g2, key = setup(48, 12, dt, seed=9)
    pmove = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    G1, G2 = torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda()
    LP1 = tgt.log_prob(G1)
    full = pmove(G1, G2, 0.05, key, tgt, tgt, 5, LP1, impl=impl)
    for G in (2, 3, 4):
        rows = 48 // G
        parts = [pmove(G1[r * rows:(r + 1) * rows], G2, 0.05, key, tgt, tgt, 5, LP1[r * rows:(r + 1) * rows],
                       row0=r * rows, n=48, impl=impl) for r in range(G)]
        for i in range(4):
            assert torch.equal(torch.cat([p[i] for p in parts]), full[i])


def test_determinism(hx):
    dt = np.float32
    tgt, _ = make_target(hx, "gauss", 100, dt)
    g1, g2, key = setup(512, 100, dt, seed=11)
    G1, G2 = torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda()
    LP1 = tgt.log_prob(G1)
    a = hx.hmc_walk_move(G1, G2, 0.02, key, tgt, tgt, 10, LP1)
    b = hx.hmc_walk_move(G1, G2, 0.02, key, tgt, tgt, 10, LP1)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_nan_proposal_is_rejected(hx):
    """hmc_walk.py:55: NaN dH -> +inf -> log_accept = -inf (funnel overflow with a huge step)."""
    tgt = hx.targets.Funnel(5)
    g1, g2, key = setup(16, 5, np.float32, seed=2)
    G1, G2 = torch.from_numpy(g1 * 30).cuda(), torch.from_numpy(g2 * 30).cuda()
    q, la, _, lp = hx.hmc_walk_move(G1, G2, 50.0, key, tgt, tgt, 10, tgt.log_prob(G1))
    la = la.cpu().numpy()
    assert not np.any(np.isnan(la))
    assert np.all(la <= 0)
    assert np.any(np.isneginf(la))
