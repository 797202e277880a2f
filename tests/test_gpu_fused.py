"""Fused draw + projection kernel (csrc/hx_tc_fused.cuh: P0 = normal(key, (n, n)) generated inside the tcgen05 GEMM,
hmc_walk.py:36) against the staged path (k_rng_planes -> HBM -> k_tc_gemm) and the dense oracle.

The fused kernel accumulates the three products of the bf16 hi/lo split in ONE TMEM accumulator; with
hx_ctx_set_tuning(18, 1) the staged kernel does the same, and the two must then agree bit for bit (same draws, same
operand split, same MMA order per output element).  Against the default staged kernel (two accumulators) and the
f64 oracle the stated tensor-core tolerances apply (test_gpu_tensorcore.py)."""
import numpy as np
import pytest
import torch

from oracle import jaxrng as R, moves as OM, targets as OT

pytestmark = pytest.mark.gpu


@pytest.fixture()
def knobs(hx):
    from hemcee_b200 import _lib, _rt
    lib, ctx = _lib.load(), _rt.ctx(torch.device("cuda", 0))

    def set_(what, value):
        _lib.check(lib.hx_ctx_set_tuning(ctx, what, value), "hx_ctx_set_tuning")
    yield set_
    for what, value in ((1, 0), (15, 1), (16, 0), (17, 0), (18, 0), (19, 0), (21, 0)):
        lib.hx_ctx_set_tuning(ctx, what, value)
    hx.config.update("precision", "auto")
    hx.config.update("leapfrog_form", "auto")


def _problem(hx, n, d, seed=0, kappa=100.0):
    ks = R.split(R.PRNGKey(seed), 4)
    cov = OT.make_covariance_skewed(ks[0], d, kappa, np.float64)
    mean = R.normal(ks[3], (d,), np.float64) * 0.1
    g1 = (R.normal(ks[1], (n, d), np.float32) * 0.5).astype(np.float32)
    g2 = (R.normal(ks[2], (n, d), np.float32) * 0.8).astype(np.float32)
    return hx.targets.GaussianFull(cov=cov, mean=mean), OT.make_gaussian(cov, mean), g1, g2, ks[3]


# (n, d): one unit / two units without sharing (d <= 512), CTA pair sharing the draw (d > 512), ragged rows and columns
SHAPES = [(256, 200), (640, 300), (384, 512), (384, 600), (1000, 1000), (200, 72)]


@pytest.mark.parametrize("n,d", SHAPES)
@pytest.mark.parametrize("nw", [28, 24, 20, 16, 8])
def test_fused_equals_staged_one_accumulator(hx, knobs, n, d, nw):
    hx.config.update("precision", "bf16x3")
    hx.config.update("leapfrog_form", "stepwise")
    tgt, _, g1, g2, key = _problem(hx, n, d)
    G1, G2 = torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda()
    LP1 = tgt.log_prob(G1)
    knobs(15, 0); knobs(18, 1)
    staged = hx.hmc_walk_move(G1, G2, 0.03, key, tgt, tgt, 2, LP1)
    knobs(15, 1); knobs(16, nw)
    fused = hx.hmc_walk_move(G1, G2, 0.03, key, tgt, tgt, 2, LP1)
    for a, b in zip(staged, fused):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n,d,slices", [(2048, 600, 2), (4096, 200, 4), (4160, 1000, 0)])
def test_fused_split_k_equals_staged_split_k(hx, knobs, n, d, slices):
    """Split-K (grid.y slices of K summed in fp32 in a fixed order; default 2 from 4096 walkers per group): the fused kernel
    against the staged kernel with the same slices and one accumulator."""
    hx.config.update("precision", "bf16x3")
    hx.config.update("leapfrog_form", "stepwise")
    tgt, _, g1, g2, key = _problem(hx, n, d)
    rows = 384
    G1, G2 = torch.from_numpy(g1[:rows]).cuda(), torch.from_numpy(g2).cuda()
    LP1 = tgt.log_prob(G1)
    eff = slices if slices else 2
    knobs(15, 0); knobs(18, 1); knobs(1, eff)
    staged = hx.hmc_walk_move(G1, G2, 0.03, key, tgt, tgt, 1, LP1, row0=128, n=n)
    knobs(15, 1); knobs(17, slices); knobs(1, 0)
    fused = hx.hmc_walk_move(G1, G2, 0.03, key, tgt, tgt, 1, LP1, row0=128, n=n)
    for a, b in zip(staged, fused):
        assert torch.equal(a, b)


@pytest.mark.parametrize("fmt", [1, 2])
@pytest.mark.parametrize("n,d", [(384, 600), (640, 300)])
def test_fused_matches_oracle(hx, knobs, n, d, fmt):
    hx.config.update("precision", "bf16x3")
    tgt, (olp, og), g1, g2, key = _problem(hx, n, d)
    G1, G2 = torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda()
    LP1 = tgt.log_prob(G1)
    knobs(15, 1); knobs(21, fmt)
    q, la, pp, lp = hx.hmc_walk_move(G1, G2, 0.03, key, tgt, tgt, 4, LP1)
    a64, b64 = g1.astype(np.float64), g2.astype(np.float64)
    q64, la64, pp64, lp64 = OM.hmc_walk_move(a64, b64, 0.03, key, olp, og, 4, olp(a64), rng_dtype=np.float32)
    np.testing.assert_allclose(q.cpu().numpy(), q64, rtol=0, atol=2e-4 * np.abs(q64).max())
    np.testing.assert_allclose(pp.cpu().numpy(), pp64, rtol=0, atol=2e-4 * np.abs(pp64).max())
    np.testing.assert_allclose(lp.cpu().numpy(), lp64, rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(la.cpu().numpy(), la64, rtol=0, atol=5e-3)


@pytest.mark.parametrize("n,d", SHAPES + [(4160, 1000), (4160, 203), (4100, 64)])      # from 4096 walkers: split-K 2 with red.add (d = 203: scalar path)
def test_fused_projection_error_of_both_operand_forms(hx, knobs, n, d):
    """S0 = P0 C of the fused kernel against an f64 product of the same draws, for the bf16 hi/lo form (FMT 0, three products) and
    the fp16 + fp8 form (FMT 1: fp16 main product, e4m3 x e5m2 correction products at true scale; the default under "auto").
    With step 1e-9 and L = 1 the returned projected momentum IS S0.  Bounds: rms error / rms S0 <= 1e-5 (bf16 hi/lo, measured
    ~3e-6) and <= 3e-5 (fp16 + fp8); both forms agree with each other to 1e-4 of the scale everywhere."""
    hx.config.update("precision", "bf16x3")
    hx.config.update("leapfrog_form", "stepwise")
    tgt, (olp, og), g1, g2, key = _problem(hx, n, d)
    rows = min(n, 384)
    G1, G2 = torch.from_numpy(g1[:rows]).cuda(), torch.from_numpy(g2).cuda()
    LP1 = tgt.log_prob(G1)
    p0 = R.normal(key, (n, n), np.float32)[:rows].astype(np.float64)
    c = g2.astype(np.float64)
    ref = p0 @ ((c - c.mean(0)) / np.sqrt(n))
    out = {}
    for fmt in (1, 2):
        knobs(15, 1); knobs(21, fmt)
        pp = hx.hmc_walk_move(G1, G2, 1e-9, key, tgt, tgt, 1, LP1, row0=0, n=n)[2].cpu().numpy().astype(np.float64)
        out[fmt] = pp
        err = np.sqrt(np.mean((pp - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))
        print(f"n={n} d={d} fmt={fmt}: rms error {err:.2e}, max {np.abs(pp - ref).max() / np.abs(ref).max():.2e}")
        assert err <= (1e-5 if fmt == 1 else 3e-5)
    assert np.abs(out[1] - out[2]).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("fmt", [1, 2])
def test_fused_row_shards_and_reruns_are_bit_identical(hx, knobs, fmt):
    hx.config.update("precision", "bf16x3")
    knobs(15, 1); knobs(21, fmt)
    tgt, _, g1, g2, key = _problem(hx, 512, 600)
    G1, G2 = torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda()
    LP1 = tgt.log_prob(G1)
    full = hx.hmc_walk_move(G1, G2, 0.03, key, tgt, tgt, 3, LP1)
    again = hx.hmc_walk_move(G1, G2, 0.03, key, tgt, tgt, 3, LP1)
    for a, b in zip(full, again):
        assert torch.equal(a, b)
    parts = [hx.hmc_walk_move(G1[r * 128:(r + 1) * 128], G2, 0.03, key, tgt, tgt, 3, LP1[r * 128:(r + 1) * 128],
                              row0=r * 128, n=512) for r in range(4)]
    for i in range(4):
        assert torch.equal(torch.cat([p[i] for p in parts]), full[i])
