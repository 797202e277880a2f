"""The C-ABI shared library loads and exports every symbol include/hx.h declares (no compute, no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hx.h")).read()
    return sorted(set(re.findall(r"HX_API\s+[\w\s\*]+?\b(hx_\w+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = _declared()
    for must in ("hx_walk_move", "hx_side_move", "hx_accept", "hx_run_iterations", "hx_split", "hx_normal",
                 "hx_uniform", "hx_choice2", "hx_log_prob", "hx_leapfrog_walk", "hx_leapfrog_side", "hx_ctx_create"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from hemcee_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libhx.so not built: run `python -c 'import __graft_entry__ as g; g.build()'`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"libhx.so does not export {name}"
    assert sorted(_lib.EXPORTS) == _declared()          # the ctypes binding covers the whole header
    assert _lib.load().hx_abi_version() == 1


def test_argument_validation_without_gpu():
    """Entry points validate pointers / enums before touching the device."""
    from hemcee_b200 import _lib
    lib = _lib.load()
    key = (ctypes.c_uint32 * 2)(0, 1)
    assert lib.hx_split(key, 0, 4, None, None) == -1                      # HX_E_NULL
    assert lib.hx_random_bits(key, 0, 16, 4, ctypes.c_void_p(8), None) == -3   # bad bit width
    assert lib.hx_normal(key, 7, 0, 4, ctypes.c_void_p(8), None) == -3          # bad impl
    assert lib.hx_uniform(key, 0, 9, 0.0, 1.0, 4, ctypes.c_void_p(8), None) == -3
    assert b"hx_uniform" in lib.hx_last_error()
    assert lib.hx_walk_move(None, 0, 0, None, None, None, 4, 0, 4, 0.1, key, 1, None, None, None, None, None, 0, None) == -1


def test_round2_entry_points_validate_arguments_without_gpu():
    """hx_stats_feed / hx_ctx_set_chain_stats / hx_spd_inverse reject NULL handles before touching the device; the Python mirror of
    hx_chain_stats has the header's field order."""
    from hemcee_b200 import _lib
    lib = _lib.load()
    st = _lib.hx_chain_stats()
    assert lib.hx_stats_feed(None, 0, ctypes.byref(st), ctypes.c_void_p(8), None) == -1
    assert lib.hx_ctx_set_chain_stats(None, ctypes.byref(st)) == -1
    assert lib.hx_spd_inverse(None, ctypes.c_void_p(8), 4, 4, ctypes.c_double(1.0), ctypes.c_void_p(8), 4, None) == -1
    assert lib.hx_stats_row_blocks() >= 1
    src = open(os.path.join(ROOT, "include", "hx.h")).read()
    body = src[src.index("typedef struct hx_chain_stats {"):src.index("} hx_chain_stats;")]
    names = re.findall(r"[\*\s,](\w+_dev|nwalkers|dim|n_track|walker_stride|n_dims|max_lag)\b(?=[,;])", body)
    assert names == [f[0] for f in _lib.hx_chain_stats._fields_], names


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "h-emcee_b200", "hemcee_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'"""[\s\S]*?"""', "", src), fn


def test_precision_and_leapfrog_enums_match_the_python_mirror():
    """hemcee_b200.config's precision / leapfrog-form names carry the values of the HX_PREC_* / HX_LEAP_* enums of the header."""
    from hemcee_b200 import _rt
    src = open(os.path.join(ROOT, "include", "hx.h")).read()
    enums = dict((k, int(v)) for k, v in re.findall(r"\b(HX_(?:PREC|LEAP)_\w+)\s*=\s*(\d+)", src))
    prec = {k[len("HX_PREC_"):].lower(): v for k, v in enums.items() if k.startswith("HX_PREC_")}
    leap = {k[len("HX_LEAP_"):].lower(): v for k, v in enums.items() if k.startswith("HX_LEAP_")}
    assert prec == _rt.config.PRECISIONS
    assert leap == _rt.config.LEAPFROG_FORMS
