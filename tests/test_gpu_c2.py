"""BASELINE.json's configuration c2 exactly as stated — 100-dim ill-conditioned Gaussian (condition number 1e4,
Sigma = Q diag(0.1 linspace(1, 1e4, 100)) Q^T with Q = qr(normal(PRNGKey(1), (100, 100))), reference
src/hemcee/tests/distribution.py:76-97), 1024 walkers, walk and side moves, L = 10 — through run_mcmc (C-ABI batch loop)
against the oracle sampler, iteration by iteration, in x64 and in f32; and the Hamiltonian
side move against the oracle at n = 1700 walkers per group, where jax.random.choice shuffles in TWO rounds
(n > 1625: reference moves/hamiltonian/hmc_side.py:43-47)."""
import numpy as np
import pytest
import torch

from oracle import jaxrng as R, moves as OM, targets as OT, sampler as OS

pytestmark = pytest.mark.gpu

W, D, KAPPA, L_STEPS, EPS = 1024, 100, 1e4, 10, 0.1


def _c2(hx, dtype, start="prior"):
    cov = OT.make_covariance_skewed(R.PRNGKey(1), D, KAPPA, np.float64)
    k_init, k_run = R.split(R.PRNGKey(0), 2)
    x0 = R.normal(k_init, (W, D), dtype)                      # tutorial idiom (docs/index.rst:29-36)
    if start == "equilibrated":                               # a draw from the target: non-trivial accept decisions
        x0 = (x0.astype(np.float64) @ np.linalg.cholesky(cov).T).astype(dtype)
    return hx.targets.GaussianFull(cov=cov), OT.make_gaussian(cov, np.zeros(D)), x0, k_run


# (start, step size per move): BASELINE's c2 run (N(0, I) start, eps = 0.1: acceptance ~ 1), and an equilibrated ensemble with
# step sizes that reject 15 % (walk) / 7 % (side) of the proposals, so that the accept decisions are a real check
CASES = [("prior", {"walk": EPS, "side": EPS}), ("equilibrated", {"walk": 0.4, "side": 1.0})]


@pytest.mark.parametrize("start,eps_by_move", CASES)
@pytest.mark.parametrize("move", ["walk", "side"])
def test_c2_chain_parity_x64(hx, move, start, eps_by_move):
    """x64: identical accept decisions; positions within 1e-7 of the ensemble scale and log-probs within 1e-6 relative over
    T iterations (kappa = 1e4: one leapfrog kick amplifies a position rounding error by up to eps * lambda_max(P M) ~ 1e2)."""
    hx.config.update("enable_x64", True)
    try:
        T, EPS = 6, eps_by_move[move]
        tgt, otgt, x0, k_run = _c2(hx, np.float64, start)
        pm = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
        om = OM.hmc_walk_move if move == "walk" else OM.hmc_side_move
        s = hx.HamiltonianEnsembleSampler(W, D, tgt, step_size=EPS, L=L_STEPS, move=pm, adapt_inital_step_size=False,
                                          adapt_step_size=False, adapt_length=False, verbose=False)
        chain = s.run_mcmc(k_run, x0, T, warmup=0)
        o = OS.HamiltonianEnsembleSampler(W, D, otgt, step_size=EPS, L=L_STEPS, move=om, adapt_inital_step_size=False,
                                          adapt_step_size=False, adapt_length=False, dtype=np.float64)
        ref = o.run_mcmc(k_run, x0, T, warmup=0)
        np.testing.assert_array_equal(s.diagnostics_main["accepts"], o.diagnostics_main["accepts"])
        np.testing.assert_allclose(chain, ref, rtol=0, atol=1e-7 * np.abs(ref).max())
        np.testing.assert_allclose(s.get_logprob(), o.get_logprob(), rtol=1e-6, atol=1e-6)
        if start == "equilibrated":
            assert 0.5 < np.mean(o.diagnostics_main["acceptance_rate"]) < 0.97
    finally:
        hx.config.update("enable_x64", False)


@pytest.mark.parametrize("start,eps_by_move", CASES)
@pytest.mark.parametrize("move", ["walk", "side"])
def test_c2_chain_parity_f32(hx, move, start, eps_by_move):
    """f32 (the dtype the throughput numbers are quoted in) under precision "exact" (draws bit-equal to the oracle's): over the
    first 4 iterations every half-move of the main-phase schedule (keys[t,0] move1, [t,2] accept1, [t,1] move2, [t,3] accept2,
    hamiltonian_ensemble.py:284-318) is run from the ORACLE's current f32 state by both sides.  Per half-move: proposals within
    2e-3 of the ensemble scale and log-accept within 2e-2 for every walker (f32 energy noise of a kappa = 1e4 target), log-probs
    within 2e-3 relative; accept decisions equal except where log u lies inside that noise (< 1 % of the walkers).  Continuing
    from the oracle's state keeps a single flipped decision from decorrelating the rest of the run (one changed complement row
    moves every walk proposal by ~4 % of the scale), which is what a free-running f32 comparison measures instead."""
    T, EPS = 4, eps_by_move[move]
    tgt, (olp, og), x0, k_run = _c2(hx, np.float32, start)
    pm = hx.hmc_walk_move if move == "walk" else hx.hmc_side_move
    om = OM.hmc_walk_move if move == "walk" else OM.hmc_side_move
    n = W // 2
    keys = R.split(R.split(k_run, 3)[2], 4 * T).reshape(T, 4, 2)            # run_mcmc: keys[2] drives the main phase
    g = [x0[:n].copy(), x0[n:].copy()]
    lp = [olp(g[0]).astype(np.float32), olp(g[1]).astype(np.float32)]
    scale = np.abs(x0).max()
    flips = total = 0
    hx.config.update("precision", "exact")
    try:
        for t in range(T):
            for half, (im, ia) in enumerate(((0, 2), (1, 3))):
                cur, comp, lpc = g[half], g[1 - half], lp[half]
                qo, lao, _, lpo = om(cur, comp, np.float32(EPS), keys[t, im], olp, og, L_STEPS, lpc)
                C, Cc, LPC = torch.from_numpy(cur).cuda(), torch.from_numpy(comp).cuda(), torch.from_numpy(lpc).cuda()
                qg, lag, _, lpg = pm(C, Cc, EPS, keys[t, im], tgt, tgt, L_STEPS, LPC)
                np.testing.assert_allclose(qg.cpu().numpy(), qo, rtol=0, atol=2e-3 * scale)
                np.testing.assert_allclose(lpg.cpu().numpy(), lpo, rtol=2e-3, atol=2e-2)
                np.testing.assert_allclose(lag.cpu().numpy(), lao, rtol=0, atol=2e-2)
                new, acc_o = OM.accept_proposal(cur, qo, lao, keys[t, ia])
                new_lp, _ = OM.accept_proposal(lpc, lpo, lao, keys[t, ia])
                _, _, acc_g = hx.moves.accept_proposal(C.clone(), qg, LPC.clone(), lpg, lag, keys[t, ia])
                flips += int(np.sum(acc_g.cpu().numpy().astype(bool) != acc_o.astype(bool)))
                total += n
                g[half], lp[half] = new.astype(np.float32), new_lp.astype(np.float32)
    finally:
        hx.config.update("precision", "auto")
    assert flips <= 0.01 * total, (flips, total)


def test_side_move_two_shuffle_rounds_matches_oracle(hx):
    """n = 1700 > 1625: `choice(key, n, (2,), replace=False)` runs two sort rounds; the whole side half-move (choices, direction
    rows, scalar-momentum leapfrog, dH) against the oracle in x64."""
    hx.config.update("enable_x64", True)
    try:
        n, d, L, eps = 1700, 8, 5, 0.05
        ks = R.split(R.PRNGKey(7), 4)
        cov = OT.make_covariance_skewed(ks[0], d, 100.0, np.float64)
        olp, og = OT.make_gaussian(cov, np.zeros(d))
        tgt = hx.targets.GaussianFull(cov=cov)
        g1 = R.normal(ks[1], (n, d), np.float64)
        g2 = R.normal(ks[2], (n, d), np.float64) * 1.5
        G1, G2 = torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda()
        q, la, pp, lp = hx.hmc_side_move(G1, G2, eps, ks[3], tgt, tgt, L, tgt.log_prob(G1))
        qr, lar, ppr, lpr = OM.hmc_side_move(g1, g2, eps, ks[3], olp, og, L, olp(g1))
        np.testing.assert_allclose(q.cpu().numpy(), qr, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(pp.cpu().numpy(), ppr, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(la.cpu().numpy(), lar, rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(lp.cpu().numpy(), lpr, rtol=1e-9, atol=1e-9)
    finally:
        hx.config.update("enable_x64", False)
