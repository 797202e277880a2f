"""ctypes binding of libhx.so (include/hx.h).

The CUDA extension is the only compute path of this package: if the shared
library is missing or a call fails, an exception is raised — there is no CPU or
PyTorch fallback (BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes as C
import os

HX_F32, HX_F64 = 0, 1
HX_RNG_PARTITIONABLE, HX_RNG_LEGACY = 0, 1
HX_TARGET_GAUSS_ISO, HX_TARGET_GAUSS_FULL, HX_TARGET_ROSENBROCK, HX_TARGET_FUNNEL = 0, 1, 2, 3
HX_MOVE_FROZEN_GRAD = 1
HX_PREC_AUTO, HX_PREC_EXACT, HX_PREC_BF16X3, HX_PREC_BF16 = 0, 1, 2, 3

LIB_PATH = os.environ.get("HX_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libhx.so")   # HX_LIB: diagnostic builds


class HxError(RuntimeError):
    """A libhx call returned a non-zero status."""


class hx_target(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("dim", C.c_int32), ("ld", C.c_int32), ("has_lower0", C.c_int32),
        ("mean", C.c_void_p), ("precision", C.c_void_p),
        ("a", C.c_double), ("b", C.c_double), ("lower0", C.c_double),
    ]


class hx_adapt_params(C.Structure):
    _fields_ = [("use_da", C.c_int32), ("use_chees", C.c_int32), ("agg_harmonic", C.c_int32), ("pass_L", C.c_int32),
                ("da_target_accept", C.c_double), ("da_t0", C.c_double), ("da_gamma", C.c_double), ("da_kappa", C.c_double),
                ("da_init_step", C.c_double),
                ("ch_T_min", C.c_double), ("ch_T_max", C.c_double), ("ch_T_interp", C.c_double), ("ch_jitter", C.c_double),
                ("ch_lr", C.c_double), ("ch_beta1", C.c_double), ("ch_beta2", C.c_double), ("ch_reg", C.c_double),
                ("pass_step", C.c_double)]


class hx_adapt_state(C.Structure):
    _fields_ = [("da_iteration", C.c_int64), ("da_log_step", C.c_double), ("da_log_step_bar", C.c_double), ("da_H_bar", C.c_double),
                ("ch_log_T", C.c_double), ("ch_log_T_bar", C.c_double), ("ch_m1", C.c_double), ("ch_m2", C.c_double),
                ("ch_iteration", C.c_double), ("ch_halton", C.c_double)]


class hx_chain_stats(C.Structure):
    _fields_ = [("nwalkers", C.c_int32), ("dim", C.c_int32),
                ("n_track", C.c_int32), ("walker_stride", C.c_int32), ("n_dims", C.c_int32), ("max_lag", C.c_int32),
                ("dims_dev", C.c_void_p), ("count_dev", C.c_void_p),
                ("sum_dev", C.c_void_p), ("sumsq_dev", C.c_void_p), ("partial_dev", C.c_void_p),
                ("lagsum_dev", C.c_void_p), ("ring_dev", C.c_void_p), ("first_dev", C.c_void_p), ("total_dev", C.c_void_p)]


_P = C.c_void_p
_KEY = C.POINTER(C.c_uint32)
_PROTOS = {
    "hx_abi_version": (C.c_int, []),
    "hx_last_error": (C.c_char_p, []),
    "hx_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "hx_ctx_destroy": (C.c_int, [_P]),
    "hx_ctx_workspace_bytes": (C.c_size_t, [_P]),
    "hx_ctx_set_precision": (C.c_int, [_P, C.c_int]),
    "hx_ctx_set_leapfrog_form": (C.c_int, [_P, C.c_int]),
    "hx_walk_prefetch": (C.c_int, [_P, C.c_int, C.c_int32, C.c_int32, C.c_int32, _KEY]),
    "hx_tc_microbench": (C.c_int, [_P, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), _P]),
    "hx_ctx_set_tuning": (C.c_int, [_P, C.c_int, C.c_int]),
    "hx_ctx_last_path": (C.c_int, [_P]),
    "hx_ctx_timing_enable": (C.c_int, [_P, C.c_int]),
    "hx_ctx_timing_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "hx_ctx_timing_phases": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "hx_split": (C.c_int, [_KEY, C.c_int, C.c_int64, _P, _P]),
    "hx_random_bits": (C.c_int, [_KEY, C.c_int, C.c_int, C.c_int64, _P, _P]),
    "hx_uniform": (C.c_int, [_KEY, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int64, _P, _P]),
    "hx_normal": (C.c_int, [_KEY, C.c_int, C.c_int, C.c_int64, _P, _P]),
    "hx_choice2": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int32, _P, _P]),
    "hx_log_prob": (C.c_int, [_P, C.c_int, C.POINTER(hx_target), _P, C.c_int64, _P, _P, _P]),
    "hx_walk_move": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(hx_target), _P, _P, C.c_int32, C.c_int32, C.c_int32,
                               C.c_double, _KEY, C.c_int32, _P, _P, _P, _P, _P, C.c_int, _P]),
    "hx_side_move": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(hx_target), _P, _P, C.c_int32, C.c_int32, C.c_int32,
                               C.c_double, _KEY, C.c_int32, _P, _P, _P, _P, _P, C.c_int, _P]),
    "hx_leapfrog_walk": (C.c_int, [_P, C.c_int, C.POINTER(hx_target), _P, _P, _P, C.c_int32, C.c_int32, C.c_double,
                                   C.c_int32, _P, _P]),
    "hx_leapfrog_side": (C.c_int, [_P, C.c_int, C.POINTER(hx_target), _P, _P, _P, C.c_int32, C.c_double, C.c_int32, _P]),
    "hx_accept": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _KEY, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                            _P, _P]),
    "hx_run_iterations": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(hx_target), _P, _P, C.c_int32, C.c_double,
                                    C.c_int32, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "hx_ctx_set_chain_thinning": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "hx_stats_row_blocks": (C.c_int32, []),
    "hx_stats_feed": (C.c_int, [_P, C.c_int, C.POINTER(hx_chain_stats), _P, _P]),
    "hx_ctx_set_chain_stats": (C.c_int, [_P, C.POINTER(hx_chain_stats)]),
    "hx_spd_inverse": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_double, _P, C.c_int32, _P]),
    "hx_run_warmup": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(hx_target), _P, _P, C.c_int32, C.POINTER(hx_adapt_params),
                                C.POINTER(hx_adapt_state), _P, C.c_int64, _P, _P, _P, _P, _P]),
    "hx_vanilla_move": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(hx_target), _P, _P, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_double, _KEY, _P, _P, _P, _P, _P]),
    "hx_run_vanilla_iterations": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(hx_target), _P, _P, C.c_int32, C.c_double, _P,
                                            C.c_int64, _P, _P, _P, _P, _P]),
    "hx_dist_init": (C.c_int, [_P, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "hx_dist_connect": (C.c_int, [_P, _P]),
    "hx_dist_buffers": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "hx_dist_destroy": (C.c_int, [_P]),
    "hx_run_iterations_sharded": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(hx_target), C.c_double, C.c_int32, _P,
                                            C.c_int64, _P, _P, _P, _P, _P]),
}

EXPORTS = tuple(_PROTOS)
_lib = None


def load() -> C.CDLL:
    """Load libhx.so (once) and attach prototypes.  Raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HxError(
            f"libhx.so not found at {LIB_PATH}: build it with `make -C h-emcee_b200 -j8` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.hx_abi_version() != 1:
        raise HxError(f"libhx ABI version {lib.hx_abi_version()} != 1")
    _lib = lib
    return lib


def check(status: int, what: str = "libhx"):
    if status != 0:
        msg = load().hx_last_error().decode("utf-8", "replace")
        raise HxError(f"{what} failed with status {status}: {msg}")


def key_arg(key):
    """uint32[2] host key -> ctypes array."""
    import numpy as np
    k = np.asarray(key, dtype=np.uint32).reshape(2)
    return (C.c_uint32 * 2)(int(k[0]), int(k[1]))
