"""Runtime plumbing: per-device libhx contexts, dtype codes, the current CUDA stream.

PyTorch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class _Config:
    """Process-wide switches mirroring the JAX flags the reference depends on.

    enable_x64 : like `jax_enable_x64` — float64 state and 64 random bits per draw
                 (reference tests set it: src/hemcee/tests/test_proposal.py:7).
    rng_impl   : "partitionable" (JAX >= 0.5 default, the mode the reference authors ran)
                 or "legacy" (`jax_threefry_partitionable=False`).
    precision  : how the walk move's dense contractions are evaluated (include/hx.h HX_PREC_*):
                 "auto" (exact fp32/fp64 SIMT kernels; split-operand tensor cores ("f16f8") for large fp32
                 full-covariance Gaussian problems), "exact", "bf16x3", "bf16", "tf32x3", "f16f8".
    leapfrog_form : tensor-core walk path only (include/hx.h HX_LEAP_*): "stepwise" kicks on every walker,
                 "propagator" (L-step linear map of a Gaussian target applied in one GEMM), or "auto".
    """

    PRECISIONS = {"auto": 0, "exact": 1, "bf16x3": 2, "bf16": 3, "tf32x3": 4, "f16f8": 5}
    LEAPFROG_FORMS = {"auto": 0, "stepwise": 1, "propagator": 2}

    def __init__(self):
        self.enable_x64 = False
        self.rng_impl = "partitionable"
        self.precision = "auto"
        self.leapfrog_form = "auto"
        self.device_adaptation = True     # warm-up adapters resident on the device (hx_run_warmup) when applicable

    def update(self, name, value):
        if name in ("enable_x64", "jax_enable_x64"):
            self.enable_x64 = bool(value)
        elif name in ("rng_impl",):
            if value not in ("partitionable", "legacy"):
                raise ValueError("rng_impl must be 'partitionable' or 'legacy'")
            self.rng_impl = value
        elif name == "precision":
            if value not in self.PRECISIONS:
                raise ValueError(f"precision must be one of {sorted(self.PRECISIONS)}")
            self.precision = value
            for h in _ctx.values():
                _lib.check(_lib.load().hx_ctx_set_precision(h, self.PRECISIONS[value]), "hx_ctx_set_precision")
        elif name == "leapfrog_form":
            if value not in self.LEAPFROG_FORMS:
                raise ValueError(f"leapfrog_form must be one of {sorted(self.LEAPFROG_FORMS)}")
            self.leapfrog_form = value
            for h in _ctx.values():
                _lib.check(_lib.load().hx_ctx_set_leapfrog_form(h, self.LEAPFROG_FORMS[value]), "hx_ctx_set_leapfrog_form")
        elif name == "device_adaptation":
            self.device_adaptation = bool(value)
        elif name == "jax_threefry_partitionable":
            self.rng_impl = "partitionable" if value else "legacy"
        else:
            raise ValueError(f"unknown config option {name}")


config = _Config()
_ctx = {}


def default_dtype() -> torch.dtype:
    return torch.float64 if config.enable_x64 else torch.float32


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return _lib.HX_F32
    if dt == torch.float64:
        return _lib.HX_F64
    raise TypeError(f"libhx computes in float32 or float64, got {dt}")


def impl_code(impl=None) -> int:
    impl = config.rng_impl if impl is None else impl
    return _lib.HX_RNG_LEGACY if impl == "legacy" else _lib.HX_RNG_PARTITIONABLE


def device(dev=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.HxError("hemcee_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
    if dev is None:
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(dev)
    if dev.type != "cuda":
        raise _lib.HxError("hemcee_b200 runs on CUDA devices only")
    return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


def ctx(dev: torch.device) -> C.c_void_p:
    """The libhx context (private scratch workspace) of a device."""
    idx = dev.index
    if idx not in _ctx:
        h = C.c_void_p()
        _lib.check(_lib.load().hx_ctx_create(C.byref(h), idx), "hx_ctx_create")
        _lib.check(_lib.load().hx_ctx_set_precision(h, config.PRECISIONS[config.precision]), "hx_ctx_set_precision")
        _lib.check(_lib.load().hx_ctx_set_leapfrog_form(h, config.LEAPFROG_FORMS[config.leapfrog_form]),
                   "hx_ctx_set_leapfrog_form")
        _ctx[idx] = h
    return _ctx[idx]


def stream(dev: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def as_key(key) -> np.ndarray:
    if isinstance(key, torch.Tensor):
        key = key.detach().cpu().numpy()
    k = np.asarray(key)
    if k.dtype != np.uint32:
        k = k.astype(np.int64).astype(np.uint32) if k.dtype.kind in "iu" else k.view(np.uint32)
    return k.reshape(2)


# ---- row-sharded multi-GPU plumbing (include/hx.h section 5) ---------------------------------------------------------
_dist = {}


class _DevMem:
    """A libhx-owned device buffer exposed through __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr, shape, dtype):
        typestr = {torch.float32: "<f4", torch.float64: "<f8"}[dtype]
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def dist_setup(dev: torch.device, group, dtype: torch.dtype, n: int, dim: int):
    """Create (once per shape) the peer-mapped exchange region of this rank and connect it to the other ranks of `group`.
    torch.distributed only carries the 64-byte IPC handles.  Returns (X_full [2n, dim], lp_full [2n]) torch views."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    key = (dev.index, id(group))
    want = (n, dim, dtype, world)
    lib = _lib.load()
    h = ctx(dev)
    if key in _dist and _dist[key][0] != want:
        torch.cuda.synchronize(dev)
        dist.barrier(group)
        _lib.check(lib.hx_dist_destroy(h), "hx_dist_destroy")
        del _dist[key]
    if key not in _dist:
        handle = (C.c_uint8 * 64)()
        _lib.check(lib.hx_dist_init(h, dtype_code(dtype), rank, world, n, dim, handle), "hx_dist_init")
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=dev)
        allh = torch.empty(world * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        buf = (C.c_uint8 * (world * 64)).from_buffer_copy(bytes(allh.cpu().numpy().tobytes()))
        _lib.check(lib.hx_dist_connect(h, buf), "hx_dist_connect")
        px, pl = C.c_void_p(), C.c_void_p()
        _lib.check(lib.hx_dist_buffers(h, C.byref(px), C.byref(pl)), "hx_dist_buffers")
        X = torch.as_tensor(_DevMem(px.value, (2 * n, dim), dtype), device=dev)
        lp = torch.as_tensor(_DevMem(pl.value, (2 * n,), dtype), device=dev)
        _dist[key] = (want, X, lp)
    return _dist[key][1], _dist[key][2]
