"""Pins the oracle's jax.random restatement against the known-answer vectors in golden/rng_kat.json."""
import json
import os

import numpy as np
import pytest
from scipy.special import erfinv

from oracle import jaxrng as R

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rng_kat.json")))


def test_threefry_random123_kat():
    for v in KAT["threefry2x32"]:
        k = [int(x, 16) for x in v["key"]]
        c = [int(x, 16) for x in v["ctr"]]
        y0, y1 = R.threefry2x32(np.uint32(k[0]), np.uint32(k[1]), np.uint32(c[0]), np.uint32(c[1]))
        assert [int(y0), int(y1)] == [int(x, 16) for x in v["out"]]


def test_prngkey_layout():
    np.testing.assert_array_equal(R.PRNGKey(42), np.array([0, 42], np.uint32))
    np.testing.assert_array_equal(R.PRNGKey((5 << 32) | 7), np.array([5, 7], np.uint32))


def test_split_kat():
    for v in KAT["split"]:
        np.testing.assert_array_equal(R.split(R.PRNGKey(v["seed"]), v["num"], v["impl"]), np.array(v["out"], np.uint32))


def test_erfinv32_forms_agree_to_3_ulp():
    """The two roundings of the f32 erf_inv the oracle knows (jaxrng.ERFINV32_FORM: correctly rounded log1p + fused Horner, the
    default; and libm log1pf + unfused Horner) are the same algorithm: <= 3 ulp apart, > 90 % of the draws bit-equal."""
    key = R.PRNGKey(5)
    a = R.normal(key, (400, 401), np.float32, "partitionable", form="cr-fma")
    b = R.normal(key, (400, 401), np.float32, "partitionable", form="libm")
    np.testing.assert_array_equal(a, R.normal(key, (400, 401), np.float32, "partitionable"))      # the default form
    ulp = np.spacing(np.abs(a))
    assert np.max(np.abs(a - b) / ulp) <= 3.0
    assert np.mean(a == b) > 0.9


@pytest.mark.parametrize("form", ["cr-fma", "libm"])
def test_normal_f32_doc_values_both_forms(form):
    for v in KAT["normal_f32"]:
        key = np.array(v["key"], np.uint32) if "key" in v else R.PRNGKey(v["seed"])
        got = R.normal(key, tuple(v["shape"]), np.float32, v["impl"], form=form)
        np.testing.assert_allclose(np.atleast_1d(got), v["out"], atol=KAT["normal_tol"], rtol=0)


def test_normal_f32_doc_values():
    for v in KAT["normal_f32"]:
        key = R.PRNGKey(v["seed"]) if "seed" in v else np.array(v["key"], np.uint32)
        got = np.atleast_1d(R.normal(key, tuple(v["shape"]), np.float32, v["impl"]))
        np.testing.assert_allclose(got, np.array(v["out"], np.float32), rtol=0, atol=KAT["normal_tol"])


@pytest.mark.parametrize("impl", ["partitionable", "legacy"])
def test_bits_layout_properties(impl):
    key = R.PRNGKey(11)
    # 64-bit draws: (y0 << 32 | y1) of one hash per element, in both layouts
    for m in (1, 2, 5, 64):
        b64 = R.random_bits(key, 64, (m,), impl)
        if impl == "partitionable":
            y0, y1 = R.threefry2x32(key[0], key[1], np.zeros(m, np.uint32), np.arange(m, dtype=np.uint32))
        else:
            y0, y1 = R.threefry2x32(key[0], key[1], np.arange(m, dtype=np.uint32), np.arange(m, 2 * m, dtype=np.uint32))
        np.testing.assert_array_equal(b64, (y0.astype(np.uint64) << np.uint64(32)) | y1.astype(np.uint64))
    # partitionable bits of a prefix do not depend on the array size; legacy ones do
    a, b = R.random_bits(key, 32, (10,), impl), R.random_bits(key, 32, (20,), impl)
    assert np.array_equal(a, b[:10]) == (impl == "partitionable")
    # odd sizes in the legacy layout pad the counter array with a zero
    if impl == "legacy":
        y0, y1 = R.threefry2x32(key[0], key[1], np.array([0, 1], np.uint32), np.array([2, 0], np.uint32))
        np.testing.assert_array_equal(R.random_bits(key, 32, (3,), impl), np.array([y0[0], y0[1], y1[0]], np.uint32))


def test_uniform_ranges_and_rounding():
    key = R.PRNGKey(3)
    for dt in (np.float32, np.float64):
        u = R.uniform(key, (100000,), dt, 1e-10, 1.0)
        assert u.dtype == dt and u.min() >= dt(1e-10) and u.max() < 1.0
        z = R.normal(key, (100000,), dt)
        assert np.all(np.isfinite(z)) and abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02


def test_erfinv_polynomials_match_scipy():
    x = np.linspace(-1, 1, 200001)[1:-1]
    s = erfinv(x)
    e64 = R.erf_inv(x)
    m = np.abs(x) > 1e-3
    assert np.max(np.abs(e64 - s)[m] / np.abs(s)[m]) < 1e-12          # Giles DP polynomial coefficients intact
    x32 = x.astype(np.float32)
    e32, s32 = R.erf_inv(x32), erfinv(x32.astype(np.float64))
    assert np.max(np.abs(e32 - s32)[m] / np.abs(s32)[m]) < 1e-5      # Giles SP polynomial: ~2e-7 design error + f32 rounding
    tail = 1 - np.logspace(-16, -1, 500)
    assert np.max(np.abs(R.erf_inv(tail) - erfinv(tail)) / erfinv(tail)) < 1e-9


@pytest.mark.parametrize("impl", ["partitionable", "legacy"])
def test_shuffle_rounds_and_choice(impl):
    assert [R.shuffle_rounds(n) for n in (1, 2, 1625, 1626, 100000)] == [0, 1, 1, 2, 2]
    key = R.PRNGKey(8)
    for n in (2, 5, 50, 2000):
        p = R.permutation(key, n, impl)
        assert sorted(p.tolist()) == list(range(n))
        c = R.choice_no_replace(key, n, 2, impl)
        assert c[0] != c[1] and np.array_equal(c, p[:2])
    # one round == indices of the two smallest sort keys, ties by lower index
    ks = R.split(key, 2, impl)
    sk = R.random_bits(ks[1], 32, (50,), impl)
    order = np.lexsort((np.arange(50), sk))
    np.testing.assert_array_equal(R.choice_no_replace(key, 50, 2, impl), order[:2])
