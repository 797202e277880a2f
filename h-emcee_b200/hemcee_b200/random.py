"""`jax.random` look-alikes backed by libhx (bit-exact threefry2x32 streams).

Keys are raw uint32[2] arrays like `jax.random.PRNGKey` (the reference reshapes them as such:
src/hemcee/samplers/hamiltonian_ensemble.py:228,330).  Draws are produced on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, _rt


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    if seed < 0:
        seed += 1 << 64
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def split_device(key, num: int = 2, device=None, impl=None) -> torch.Tensor:
    """split(key, num) kept on the device as an int32 [num, 2] tensor (bit pattern of uint32)."""
    dev = _rt.device(device)
    out = torch.empty((int(num), 2), dtype=torch.int32, device=dev)
    if int(num) == 0:
        return out
    _lib.check(_lib.load().hx_split(_lib.key_arg(_rt.as_key(key)), _rt.impl_code(impl), int(num), _rt.ptr(out),
                                    _rt.stream(dev)), "hx_split")
    return out


def split(key, num: int = 2, device=None, impl=None) -> np.ndarray:
    """`jax.random.split(key, num)` -> uint32[num, 2] on the host."""
    return split_device(key, num, device, impl).cpu().numpy().view(np.uint32)


def _shape(shape):
    if isinstance(shape, int):
        return (shape,)
    return tuple(int(s) for s in shape)


def random_bits(key, bit_width: int, shape, device=None, impl=None) -> torch.Tensor:
    dev = _rt.device(device)
    shape = _shape(shape)
    count = int(np.prod(shape)) if shape else 1
    out = torch.empty(shape, dtype=torch.int32 if bit_width == 32 else torch.int64, device=dev)
    _lib.check(_lib.load().hx_random_bits(_lib.key_arg(_rt.as_key(key)), _rt.impl_code(impl), int(bit_width), count,
                                          _rt.ptr(out), _rt.stream(dev)), "hx_random_bits")
    return out


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0, device=None, impl=None) -> torch.Tensor:
    dev = _rt.device(device)
    dtype = _rt.default_dtype() if dtype is None else dtype
    shape = _shape(shape)
    count = int(np.prod(shape)) if shape else 1
    out = torch.empty(shape, dtype=dtype, device=dev)
    _lib.check(_lib.load().hx_uniform(_lib.key_arg(_rt.as_key(key)), _rt.impl_code(impl), _rt.dtype_code(dtype),
                                      float(minval), float(maxval), count, _rt.ptr(out), _rt.stream(dev)), "hx_uniform")
    return out


def normal(key, shape=(), dtype=None, device=None, impl=None, exact=None) -> torch.Tensor:
    """`jax.random.normal`.  f32: `exact=True` (default under config precision "exact") evaluates erf_inv's log1p correctly
    rounded (<= 1 ulp from the oracle); otherwise the SFU log2 form of the fused kernels (<= 3 ulp)."""
    dev = _rt.device(device)
    dtype = _rt.default_dtype() if dtype is None else dtype
    shape = _shape(shape)
    count = int(np.prod(shape)) if shape else 1
    out = torch.empty(shape, dtype=dtype, device=dev)
    if exact is None:
        exact = _rt.config.precision == "exact"
    _lib.check(_lib.load().hx_normal(_lib.key_arg(_rt.as_key(key)), _rt.impl_code(impl) | (16 if exact else 0), _rt.dtype_code(dtype), count,
                                     _rt.ptr(out), _rt.stream(dev)), "hx_normal")
    return out


def choice2(keys: torch.Tensor, n: int, impl=None) -> torch.Tensor:
    """Per key: `choice(key_i, arange(n), (2,), replace=False)` -> int32 [nkeys, 2]."""
    dev = _rt.device(keys.device)
    keys = keys.contiguous()
    out = torch.empty((keys.shape[0], 2), dtype=torch.int32, device=dev)
    _lib.check(_lib.load().hx_choice2(_rt.ctx(dev), _rt.ptr(keys), keys.shape[0], _rt.impl_code(impl), int(n),
                                      _rt.ptr(out), _rt.stream(dev)), "hx_choice2")
    return out
