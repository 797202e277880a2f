"""CPU model of the tensor-core operand splits (csrc/hx_tc_gemm.cuh): what each kind loses against an
f64 product of the same fp32 data.  The kernels accumulate in fp32 inside the tensor core; this model accumulates the split
products in f64, so it isolates the error of the operand representation -- the part the choice of kind controls.

  bf16 hi/lo      x ~ bf16(x) + bf16(x - bf16(x)),            products hi*hi + hi*lo + lo*hi          ~2^-16 per product
  fp16 + 2 e4m3   x ~ fp16(x) + r, r = x - fp16(x);           fp16(a) fp16(b) + [e4m3(a) e4m3(rb 2^11) + e4m3(ra 2^11) e4m3(b)] 2^-11
  fp16 hi/lo      like bf16 hi/lo with fp16 parts of x * 2^k  (power-of-two scale into fp16's normal range)  ~2^-22
"""
import numpy as np
import pytest
import torch


def _rn(x, dtype):
    return torch.from_numpy(np.asarray(x, np.float32)).to(dtype).to(torch.float32).numpy().astype(np.float64)


def _products(a, b, kind):
    """sum_k a[i,k] b[j,k] with the operands represented in `kind` (a: draws, b: scaled complement), f64 accumulation."""
    a = a.astype(np.float32); b = b.astype(np.float32)
    if kind == "bf16x3":
        ah, bh = _rn(a, torch.bfloat16), _rn(b, torch.bfloat16)
        al, bl = _rn(a - ah, torch.bfloat16), _rn(b - bh, torch.bfloat16)
        return ah @ bh.T + ah @ bl.T + al @ bh.T
    if kind == "f16f8":
        R = 2048.0
        ah, bh = _rn(a, torch.float16), _rn(b, torch.float16)
        a8, b8 = _rn(a, torch.float8_e4m3fn), _rn(b, torch.float8_e4m3fn)
        ra8, rb8 = _rn((a - ah) * R, torch.float8_e4m3fn), _rn((b - bh) * R, torch.float8_e4m3fn)
        return ah @ bh.T + (a8 @ rb8.T + ra8 @ b8.T) / R
    if kind == "f16x3":
        ah, bh = _rn(a, torch.float16), _rn(b, torch.float16)
        al, bl = _rn(a - ah, torch.float16), _rn(b - bh, torch.float16)
        return ah @ bh.T + ah @ bl.T + al @ bh.T
    raise ValueError(kind)


def _scale_pow2(x, top):
    amax = float(np.abs(x).max())
    m, e = np.frexp(amax)
    return float(np.ldexp(1.0, top - int(e)))


@pytest.fixture(scope="module")
def data():
    rng = np.random.default_rng(0)
    K = 4096
    a = rng.standard_normal((64, K)).astype(np.float32)                     # momentum draws
    sig = np.sqrt(0.1 * np.linspace(1.0, 1e4, 48))                          # kappa = 1e4 column scales of the complement
    b = ((rng.standard_normal((K, 48)) * sig) / np.sqrt(K)).T.astype(np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64).T
    return a, b, exact


def test_projection_operand_kinds(data):
    a, b, exact = data
    scale = np.abs(exact).max(axis=0, keepdims=True)                        # per output column (coordinate)
    sc = _scale_pow2(b, 6)                                                  # complement into [32, 64): fp16 / e4m3 range
    err = {"bf16x3": np.abs(_products(a, b, "bf16x3") - exact) / scale,
           "f16f8": np.abs(_products(a, b * sc, "f16f8") / sc - exact) / scale,
           "f16x3": np.abs(_products(a, b * sc, "f16x3") / sc - exact) / scale}
    # stated bounds: bf16 hi/lo and fp16 + 2 e4m3 are both fp32-class for this product, fp16 hi/lo far below
    msg = {k: float(v.max()) for k, v in err.items()}
    assert err["bf16x3"].max() < 2e-5, msg
    assert err["f16f8"].max() < 4e-5, msg
    assert err["f16x3"].max() < 1e-6, msg
    assert err["f16f8"].max() < 4 * err["bf16x3"].max() + 1e-6, msg
    assert err["f16x3"].max() < 0.1 * err["bf16x3"].max(), msg


def test_fp16_pair_needs_the_power_of_two_scale():
    """Without scaling, small operands push the lo part into fp16's subnormal range and the pair loses its 22 bits; a
    power-of-two scale from the row's max-abs restores them exactly (scaling by 2^k is exact in fp32)."""
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((8, 1024)) * 1e-3).astype(np.float32)          # e.g. q - mu of a narrow coordinate
    p = rng.standard_normal((1024, 8)).astype(np.float32).T
    exact = x.astype(np.float64) @ p.astype(np.float64).T
    raw = _products(x, p, "f16x3")
    s = np.array([[_scale_pow2(r, 12)] for r in x])
    scaled = _products((x * s).astype(np.float32), p, "f16x3") / s
    e_raw = np.abs(raw - exact).max() / np.abs(exact).max()
    e_sc = np.abs(scaled - exact).max() / np.abs(exact).max()
    assert e_sc < 2e-6 and e_raw > 5 * e_sc
