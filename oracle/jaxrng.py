"""Bit-exact NumPy restatement of the `jax.random` subset the hot path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The arithmetic lives in the un-vendored third-party dependency `jax`
(`jax>=0.4.20`, reference `pyproject.toml:16-19`, no lock file); what is
restated here is the published algorithm of `jax/_src/prng.py` (Threefry-2x32,
20 rounds; `threefry_split`; `threefry_random_bits` in both the legacy and the
`jax_threefry_partitionable` layouts) and `jax/_src/random.py` (`uniform`,
`normal`, `_shuffle`/`permutation`/`choice`).  Reference call sites whose bits
matter:

  split   : samplers/hamiltonian_ensemble.py:90,228,330 ; moves/hamiltonian/hmc_side.py:37 ;
            adaptation/reasonable_stepsize.py:44
  normal  : moves/hamiltonian/hmc_walk.py:36 ; moves/hamiltonian/hmc_side.py:54
  choice  : moves/hamiltonian/hmc_side.py:44-47
  uniform : samplers/sampler_utils.py:114

`impl` selects the counter layout: "partitionable" (JAX >= 0.5 default; the
mode the reference authors ran, see tests/test_oracle_golden.py) or "legacy".
`dtype` float32 draws 32 random bits per value, float64 (x64 mode) draws 64.
"""
from __future__ import annotations

import math

import numpy as np

U32 = np.uint32
U64 = np.uint64

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = U32(0x1BD11BDA)

DEFAULT_IMPL = "partitionable"


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al. 2011; jax/_src/prng.py `_threefry2x32_lowering`).

    All arguments are uint32 (scalars or arrays, broadcast together).
    Returns (y0, y1).
    """
    k0 = np.asarray(k0, dtype=U32)
    k1 = np.asarray(k1, dtype=U32)
    x0 = np.array(x0, dtype=U32, copy=True)
    x1 = np.array(x1, dtype=U32, copy=True)
    x0, x1 = np.broadcast_arrays(x0, x1)
    x0 = x0.copy()
    x1 = x1.copy()
    ks = (k0, k1, k0 ^ k1 ^ _PARITY)
    with np.errstate(over="ignore"):
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    """`jax.random.PRNGKey(seed)` as a raw uint32[2] array: [seed >> 32, seed & 0xffffffff]."""
    seed = int(seed)
    if seed < 0:
        seed += 1 << 64
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def _threefry_2x32_flat(key, count: np.ndarray) -> np.ndarray:
    """`prng.threefry_2x32(keypair, count)`: hash a flat counter array, pairing
    element e of the first half with element e of the second half (odd sizes padded)."""
    count = np.asarray(count, dtype=U32).ravel()
    m = count.size
    odd = m % 2
    if odd:
        count = np.concatenate([count, np.zeros(1, U32)])
    h = count.size // 2
    y0, y1 = threefry2x32(key[0], key[1], count[:h], count[h:])
    out = np.concatenate([y0, y1])
    return out[:m] if odd else out


def _iota_2x32(size: int):
    idx = np.arange(size, dtype=np.uint64)
    return (idx >> U64(32)).astype(U32), (idx & U64(0xFFFFFFFF)).astype(U32)


def split(key, num: int = 2, impl: str = DEFAULT_IMPL) -> np.ndarray:
    """`jax.random.split(key, num)` -> uint32[num, 2]."""
    key = np.asarray(key, dtype=U32)
    if impl == "legacy":
        counts = np.arange(2 * num, dtype=U32)
        return _threefry_2x32_flat(key, counts).reshape(num, 2)
    hi, lo = _iota_2x32(num)
    y0, y1 = threefry2x32(key[0], key[1], hi, lo)
    return np.stack([y0, y1], axis=-1)


def random_bits(key, bit_width: int, shape, impl: str = DEFAULT_IMPL) -> np.ndarray:
    """`jax._src.prng.threefry_random_bits(key, bit_width, shape)` for bit_width in {32, 64}."""
    key = np.asarray(key, dtype=U32)
    shape = tuple(int(s) for s in (shape if hasattr(shape, "__len__") else (shape,)))
    size = int(np.prod(shape)) if shape else 1
    if bit_width not in (32, 64):
        raise ValueError("bit_width must be 32 or 64")
    if impl == "legacy":
        nwords = size * bit_width // 32
        words = _threefry_2x32_flat(key, np.arange(nwords, dtype=U32))
        if bit_width == 64:
            hi = words[:size].astype(U64)
            lo = words[size:].astype(U64)
            return ((hi << U64(32)) | lo).reshape(shape)
        return words.reshape(shape)
    hi, lo = _iota_2x32(size)
    y0, y1 = threefry2x32(key[0], key[1], hi, lo)
    if bit_width == 64:
        return ((y0.astype(U64) << U64(32)) | y1.astype(U64)).reshape(shape)
    return (y0 ^ y1).reshape(shape)


def _bits_to_unit_float(bits: np.ndarray, dtype) -> np.ndarray:
    """Mantissa trick of `jax.random._uniform`: floats in [0, 1)."""
    if dtype == np.float32:
        fb = (bits >> U32(32 - 23)) | U32(0x3F800000)
        return fb.view(np.float32) - np.float32(1.0)
    fb = (bits >> U64(64 - 52)) | U64(0x3FF0000000000000)
    return fb.view(np.float64) - np.float64(1.0)


def uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0, impl: str = DEFAULT_IMPL):
    """`jax.random.uniform`: max(minval, f * (maxval - minval) + minval), all in `dtype`."""
    dtype = np.dtype(dtype).type
    nbits = 32 if dtype == np.float32 else 64
    bits = random_bits(key, nbits, shape, impl)
    f = _bits_to_unit_float(bits, dtype)
    lo = dtype(minval)
    hi = dtype(maxval)
    return np.maximum(lo, f * dtype(hi - lo) + lo).astype(dtype)


# Giles, "Approximating the erfinv function" (GPU Computing Gems 2011): the
# polynomial XLA uses for f32 (`ErfInv32`, xla/client/lib/math.cc) ...
_ERFINV32_LT5 = (
    2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
    -0.00125372503, -0.00417768164, 0.246640727, 1.50140941,
)
_ERFINV32_GE5 = (
    -0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
    -0.0076224613, 0.00943887047, 1.00167406, 2.83297682,
)
# ... and for f64 (`ErfInv64`): three branches on w = -log1p(-x*x).
_ERFINV64_LT6_25 = (
    -3.6444120640178196996e-21, -1.685059138182016589e-19, 1.2858480715256400167e-18,
    1.115787767802518096e-17, -1.333171662854620906e-16, 2.0972767875968561637e-17,
    6.6376381343583238325e-15, -4.0545662729752068639e-14, -8.1519341976054721522e-14,
    2.6335093153082322977e-12, -1.2975133253453532498e-11, -5.4154120542946279317e-11,
    1.051212273321532285e-09, -4.1126339803469836976e-09, -2.9070369957882005086e-08,
    4.2347877827932403518e-07, -1.3654692000834678645e-06, -1.3882523362786468719e-05,
    0.0001867342080340571352, -0.00074070253416626697512, -0.0060336708714301490533,
    0.24015818242558961693, 1.6536545626831027356,
)
_ERFINV64_LT16 = (
    2.2137376921775787049e-09, 9.0756561938885390979e-08, -2.7517406297064545428e-07,
    1.8239629214389227755e-08, 1.5027403968909827627e-06, -4.013867526981545969e-06,
    2.9234449089955446044e-06, 1.2475304481671778723e-05, -4.7318229009055733981e-05,
    6.8284851459573175448e-05, 2.4031110387097893999e-05, -0.0003550375203628474796,
    0.00095328937973738049703, -0.0016882755560235047313, 0.0024914420961078508066,
    -0.0037512085075692412107, 0.005370914553590063617, 1.0052589676941592334,
    3.0838856104922207635,
)
_ERFINV64_GE16 = (
    -2.7109920616438573243e-11, -2.5556418169965252055e-10, 1.5076572693500548083e-09,
    -3.7894654401267369937e-09, 7.6157012080783393804e-09, -1.4960026627149240478e-08,
    2.9147953450901080826e-08, -6.7711997758452339498e-08, 2.2900482228026654717e-07,
    -9.9298272942317002539e-07, 4.5260625972231537039e-06, -1.9681778105531670567e-05,
    7.5995277030017761139e-05, -0.00021503011930044477347, -0.00013871931833623122026,
    1.0103004648645343977, 4.8499064014085844221,
)


# How the f32 erf_inv is ROUNDED is not part of the algorithm and differs between XLA backends; the oracle fixes one
# well-defined, platform-independent form (ERFINV32_FORM = "cr-fma"):
#   * w = -log1p(-x*x) correctly rounded to f32 (evaluated in f64 and rounded once).  glibc's log1pf ("libm" form) is
#     not correctly rounded (it differs from this in ~1 % of the draws, which moves ~1 % of the normals by 1-3 ulp), and
#     XLA itself expands log1p as log(1 + x) with its own polynomial logf, so no libm is "the" reference;
#   * the Horner steps p*w + c are FUSED (one rounding): XLA's CPU compiler always allows FMA fusion
#     (cpu_compiler.cc CompilerTargetOptions: AllowFPOpFusion = Fast) and NVPTX contracts by default.
# The "libm" form (np.log1p in f32, unfused Horner) is kept for comparison: the two agree to <= 3 ulp, ~95 % bit-equal
# (tests/test_oracle_rng.py).  f64 has no such ambiguity at the tolerances used (1e-13).
ERFINV32_FORM = "cr-fma"


def _horner(coeffs, w, dtype, fma=False):
    p = np.full_like(w, dtype(coeffs[0]))
    for c in coeffs[1:]:
        if fma:      # f32 fma: the f64 product of two f32 is exact, the sum is rounded to f64 and then to f32
            p = (p.astype(np.float64) * w.astype(np.float64) + np.float64(dtype(c))).astype(dtype)
        else:
            p = dtype(c) + p * w
    return p


def erf_inv(x: np.ndarray, form: str = None) -> np.ndarray:
    """XLA's `erf_inv` (Giles' polynomials), evaluated in the dtype of `x`; `form` selects the f32 rounding ("cr-fma" |
    "libm", default ERFINV32_FORM)."""
    x = np.asarray(x)
    dt = x.dtype.type
    form = ERFINV32_FORM if form is None else form
    with np.errstate(divide="ignore", invalid="ignore"):
        if dt == np.float32 and form == "cr-fma":
            w = (-np.log1p(-(x * x).astype(np.float64))).astype(np.float32)
        else:
            w = -np.log1p(-x * x)
        if dt == np.float32:
            fma = form == "cr-fma"
            lt = w < dt(5.0)
            wa = w - dt(2.5)
            wb = np.sqrt(w) - dt(3.0)
            p = np.where(lt, _horner(_ERFINV32_LT5, wa, dt, fma), _horner(_ERFINV32_GE5, wb, dt, fma))
        else:
            lt1 = w < dt(6.25)
            lt2 = w < dt(16.0)
            wa = w - dt(3.125)
            wb = np.sqrt(w) - dt(3.25)
            wc = np.sqrt(w) - dt(5.0)
            p = np.where(
                lt1,
                _horner(_ERFINV64_LT6_25, wa, dt),
                np.where(lt2, _horner(_ERFINV64_LT16, wb, dt), _horner(_ERFINV64_GE16, wc, dt)),
            )
        res = p * x
        res = np.where(np.abs(x) == dt(1.0), x * dt(np.inf), res)
    return res.astype(dt)


def normal(key, shape=(), dtype=np.float32, impl: str = DEFAULT_IMPL, form: str = None) -> np.ndarray:
    """`jax.random.normal`: sqrt(2) * erf_inv(uniform(key, shape, nextafter(-1, 0), 1))."""
    dtype = np.dtype(dtype).type
    lo = np.nextafter(dtype(-1.0), dtype(0.0))
    u = uniform(key, shape, dtype, lo, dtype(1.0), impl)
    return (dtype(np.sqrt(2)) * erf_inv(u, form)).astype(dtype)


def shuffle_rounds(n: int) -> int:
    """Number of sort rounds in `jax.random._shuffle` (exponent 3)."""
    return int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))


def permutation(key, n: int, impl: str = DEFAULT_IMPL) -> np.ndarray:
    """`jax.random.permutation(key, arange(n))` via `_shuffle`: repeated stable sort by random keys."""
    x = np.arange(n)
    key = np.asarray(key, dtype=U32)
    for _ in range(shuffle_rounds(n)):
        ks = split(key, 2, impl)
        key, sub = ks[0], ks[1]
        sort_keys = random_bits(sub, 32, (n,), impl)
        order = np.argsort(sort_keys, kind="stable")
        x = x[order]
    return x


def choice_no_replace(key, n: int, k: int, impl: str = DEFAULT_IMPL) -> np.ndarray:
    """`jax.random.choice(key, arange(n), (k,), replace=False)` = permutation(key, n)[:k]."""
    return permutation(key, n, impl)[:k]
