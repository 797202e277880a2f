"""CPU tests of the oracle itself: the reference's property tests re-stated on the NumPy restatement
(reference tests/test_leapfrog.py:73-116, tests/affine_invariance/test_leapfrog.py:26-100,
tests/affine_invariance/test_proposal_hamiltonian.py:17-77), finite-difference checks of the analytic
gradients, and the projected-form identities the CUDA path relies on."""
import numpy as np
import pytest

from oracle import adaptation as OA, jaxrng as R, moves as OM, sampler as OS, targets as OT

dt = np.float64


def _fd_grad(lp, x, h=1e-6):
    g = np.zeros_like(x)
    for j in range(x.shape[1]):
        e = np.zeros_like(x); e[:, j] = h
        g[:, j] = (lp(x + e) - lp(x - e)) / (2 * h)
    return g


@pytest.mark.parametrize("name", ["iso", "gauss", "rosen", "funnel"])
def test_analytic_gradients(name):
    d = 5
    x = R.normal(R.PRNGKey(1), (7, d), dt) * 0.7
    if name == "iso":
        lp, g = OT.make_iso_gaussian()
    elif name == "gauss":
        lp, g = OT.make_gaussian_skewed(R.PRNGKey(3), d, 100, dt)
    elif name == "rosen":
        lp, g = OT.make_rosenbrock()
    else:
        lp, g = OT.make_funnel()
    np.testing.assert_allclose(g(x), _fd_grad(lp, x), rtol=1e-5, atol=1e-6)


def _walk_setup(n, d, seed=0):
    ks = R.split(R.PRNGKey(seed), 4)
    g1 = R.normal(ks[0], (n, d), dt)
    g2 = R.normal(ks[1], (n, d), dt)
    C = (g2 - g2.mean(0)) / np.sqrt(n)
    p = R.normal(ks[2], (n, n), dt)
    return g1, g2, C, p, ks[3]


@pytest.mark.parametrize("cond", [1, 100, 1000])
def test_time_reversal_walk_and_side(cond):
    """reference tests/test_leapfrog.py:73-116 (L=30, eps=0.05, 5-d Gaussians)."""
    lp, grad = OT.make_gaussian_skewed(R.PRNGKey(42), 5, cond, dt)
    q, _, C, p, _ = _walk_setup(10, 5)
    qL, pL, _ = OM.leapfrog_walk_move(q, p, grad, 0.05, 30, C)
    qR, pR, _ = OM.leapfrog_walk_move(qL, -pL, grad, 0.05, 30, C)
    np.testing.assert_allclose(qR, q, atol=1e-8); np.testing.assert_allclose(-pR, p, atol=1e-8)
    D = (R.normal(R.PRNGKey(5), (5,), dt) - R.normal(R.PRNGKey(6), (5,), dt)) / np.sqrt(10)
    ps = R.normal(R.PRNGKey(7), (10,), dt)
    qL, pL, _ = OM.leapfrog_side_move(q, ps, grad, 0.05, 30, D[None, :].repeat(10, 0))
    qR, pR, _ = OM.leapfrog_side_move(qL, -pL, grad, 0.05, 30, D[None, :].repeat(10, 0))
    np.testing.assert_allclose(qR, q, atol=1e-8); np.testing.assert_allclose(-pR, ps, atol=1e-8)


@pytest.mark.parametrize("move", [OM.hmc_walk_move, OM.hmc_side_move])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_affine_invariance_of_moves(move, seed):
    """reference tests/affine_invariance/test_proposal_hamiltonian.py:17-77 (eps=0.1, L=10, n=20, d=5)."""
    d, n = 5, 20
    ks = R.split(R.PRNGKey(seed), 7)
    A = R.normal(ks[0], (d, d), dt); b = R.normal(ks[1], (d,), dt)
    lp, grad = OT.make_gaussian(np.eye(d), np.zeros(d))
    lpT, gradT = OT.make_gaussian(A @ A.T, b)
    g1 = R.normal(ks[4], (n, d), dt); g2 = R.normal(ks[5], (n, d), dt)
    qa, *_ = move(g1, g2, 0.1, ks[6], lp, grad, 10, lp(g1))
    g1t, g2t = g1 @ A.T + b, g2 @ A.T + b
    qb, *_ = move(g1t, g2t, 0.1, ks[6], lpT, gradT, 10, lpT(g1t))
    np.testing.assert_allclose(qa @ A.T + b, qb, rtol=1e-7, atol=1e-8)


def test_projected_form_equals_dense_form():
    """S = p C, kick S -= a g M, dK = -1/2 sum a <g, S- + S+> reproduce the dense move."""
    n, d, L, eps = 24, 6, 9, 0.07
    lp, grad = OT.make_gaussian_skewed(R.PRNGKey(9), d, 100, dt)
    g1, g2, C, _, key = _walk_setup(n, d, seed=4)
    q_ref, la_ref, pp_ref, lp_ref = OM.hmc_walk_move(g1, g2, eps, key, lp, grad, L, lp(g1))
    P0 = R.normal(key, (n, n), dt)
    M = C.T @ C
    S = P0 @ C
    q = g1.copy(); dK = np.zeros(n)
    for k in range(L + 1):
        a = 0.5 * eps if k in (0, L) else eps
        g = -grad(q)
        S_new = S - a * (g @ M)
        dK += a * np.sum(g * (S + S_new), axis=1)
        S = S_new
        if k < L:
            q = q + eps * S
    dH = (lp(g1) - lp(q)) - 0.5 * dK
    np.testing.assert_allclose(q, q_ref, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(S, pp_ref, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(np.minimum(0, -dH), la_ref, rtol=1e-8, atol=1e-9)
    # B-form used by the tensor-core path: one product per kick, energies from the impulse sum
    mu, Pm = lp.mean, lp.precision
    B = Pm @ M
    S0 = P0 @ C; S = S0.copy(); q = g1.copy(); Qs = np.zeros_like(q)
    for k in range(L + 1):
        a = 0.5 * eps if k in (0, L) else eps
        Qs += a * (q - mu)
        S = S - a * ((q - mu) @ B)
        if k < L:
            q = q + eps * S
    dK2 = -0.5 * np.sum((Qs @ Pm) * (S0 + S), axis=1)
    np.testing.assert_allclose(q, q_ref, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(dK2, -0.5 * dK, rtol=1e-8, atol=1e-9)
    # App. A.4: cov(X2, ddof=1) = n/(n-1) M
    np.testing.assert_allclose(np.cov(g2, rowvar=False), n / (n - 1) * M, rtol=1e-10, atol=1e-12)


def test_accept_is_strict_and_same_key_same_mask():
    """sampler_utils.py:114-115 and hamiltonian_ensemble.py:295-296."""
    key = R.PRNGKey(3)
    la = np.log(R.uniform(R.PRNGKey(4), (50,), dt, 1e-3, 1.0))
    cur, prop = np.zeros((50, 2)), np.ones((50, 2))
    new, acc = OM.accept_proposal(cur, prop, la, key)
    lpn, acc2 = OM.accept_proposal(np.zeros(50), np.ones(50), la, key)
    np.testing.assert_array_equal(acc, acc2)
    np.testing.assert_array_equal(new[:, 0], lpn)
    logu = OM.accept_uniform_log(key, 50, dt)
    _, acc3 = OM.accept_proposal(cur, prop, logu, key)      # log_u < log_u is False everywhere
    assert acc3.sum() == 0


def test_nan_and_neginf_contract():
    """hmc_walk.py:55 (NaN dH -> reject) and tests/test_log_prob_validation.py (-inf rejected, NaN stored raises)."""
    lp, grad = OT.make_funnel()
    g1 = R.normal(R.PRNGKey(1), (8, 4), np.float32) * 40
    g2 = R.normal(R.PRNGKey(2), (8, 4), np.float32) * 40
    with np.errstate(all="ignore"):
        _, la, _, _ = OM.hmc_walk_move(g1, g2, 50.0, R.PRNGKey(3), lp, grad, 10, lp(g1))
    assert not np.any(np.isnan(la)) and np.all(la <= 0)
    s = OS.HamiltonianEnsembleSampler(4, 2, OT.make_iso_gaussian(), adapt_inital_step_size=False, dtype=np.float32)
    x0 = np.array([[np.nan, 0.5], [0.5, 0.5], [-0.5, 0.5], [0.5, -0.5]], np.float32)
    with pytest.raises(ValueError, match="Log probability returned NaN"):
        s.run_mcmc(R.PRNGKey(1), x0, 5, warmup=0)


@pytest.mark.parametrize("da,ch,cls", [(True, True, OA.CompositeAdapter), (True, False, OA.DualAveragingAdapter),
                                       (False, True, OA.ChEESAdapter), (False, False, OA.NoOpAdapter)])
@pytest.mark.parametrize("move", [OM.hmc_walk_move, OM.hmc_side_move])
def test_adapter_dispatch_and_run(da, ch, cls, move):
    """reference tests/test_adapter.py:79-121 + tests/test_backend.py:29-37 on the oracle sampler."""
    s = OS.HamiltonianEnsembleSampler(4, 1, OT.make_iso_gaussian(), move=move, adapt_step_size=da, adapt_length=ch,
                                      dtype=np.float32)
    x0 = R.normal(R.PRNGKey(0), (4, 1), np.float32)
    out = s.run_mcmc(R.PRNGKey(1), x0, 10, warmup=35, thin_by=2)
    assert isinstance(s.adapter, cls)
    assert out.shape == (10, 4, 1) and s.backend.iteration == 35 + 20


def test_halton_and_da_quirks():
    assert [float(OA.get_halton(i)) for i in (1, 2, 3, 4)] == [0.5, 0.25, 0.75, 0.125]
    da = OA.DualAveragingAdapter(OA.DAParameters(), 0.1, 10, np.float64)
    st = da.init(1)
    for _ in range(200000):
        st = da.update(st, np.log(np.full(8, 0.999)))
    # DA divides by (t + t0) twice, so the step drifts back to the initial value (App. C.4): the quickstart
    # notebook ends at 0.1016 after 2e5 updates with 99.9 % acceptance
    assert abs(float(da.finalize(st)[0]) - 0.1016) < 5e-4
    assert da.average_over_rate(np.array([0.0, -np.inf])) == 0.0  # harmonic mean: any zero accept prob -> 0


# ---- derivative-free moves: affine invariance (reference tests/affine_invariance/test_proposal_vanilla.py:17-90) ------------
@pytest.mark.parametrize("name", ["stretch_move", "side_move", "walk_move"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_vanilla_moves_affine_invariant_oracle(name, seed):
    """R(A X + b; phi_# pi) = A R(X; pi) + b for the same key, and identical log-accept ratios."""
    from oracle import jaxrng as R, moves as OM
    d, n = 5, 20
    ks = R.split(R.PRNGKey(seed), 7)
    A = R.normal(ks[0], (d, d), np.float64) + 3.0 * np.eye(d)
    b = R.normal(ks[1], (d,), np.float64)
    Sig = np.cov(R.normal(ks[2], (4 * d, d), np.float64).T) + np.eye(d)
    P = np.linalg.inv(Sig)
    Ainv = np.linalg.inv(A)
    lp = lambda x: -0.5 * np.einsum("bi,ij,bj->b", x, P, x)
    lp_push = lambda y: lp((y - b) @ Ainv.T)          # density of A x + b up to a constant
    g1 = R.normal(ks[4], (n, d), np.float64)
    g2 = R.normal(ks[5], (n, d), np.float64)
    t1, t2 = g1 @ A.T + b, g2 @ A.T + b
    mv = getattr(OM, name)
    q, la, _ = mv(g1, g2, ks[6], lp, lp(g1))
    qt, lat, _ = mv(t1, t2, ks[6], lp_push, lp_push(t1))
    np.testing.assert_allclose(qt, q @ A.T + b, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(lat, la, rtol=1e-8, atol=1e-8)
