"""ctypes binding of ``libsleekit_b200.so`` (the C ABI in include/sleekit_b200.h).

There is no CPU fallback anywhere in this package: if the library cannot be
loaded, or no CUDA device is present, every hot-path entry point raises.
"""

from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "sleekit_b200.h")
LIB_PATH = os.path.join(HERE, "lib", "libsleekit_b200.so")


class SlkCodebook(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("size", C.c_int32),
        ("lo", C.c_double),
        ("hi", C.c_double),
        ("step", C.c_double),
        ("values", C.c_void_p),
        ("limits", C.c_void_p),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_INT = C.c_int
_D = C.c_double
_SZ = C.c_size_t
_CB = C.POINTER(SlkCodebook)

# name -> (restype, argtypes); mirrors include/sleekit_b200.h declaration by declaration
SIGNATURES = {
    "slk_abi_version": (_INT, []),
    "slk_last_error": (C.c_char_p, []),
    "slk_launch_count": (_I64, []),
    "slk_device_info": (_INT, [C.POINTER(_INT)] * 3),
    "slk_round_f32": (_INT, [_P, _I64, _CB, _INT, _P, _P, _P]),
    "slk_round_f64": (_INT, [_P, _I64, _CB, _INT, _P, _P, _P]),
    "slk_scale_axis_f32": (_INT, [_P, _I64, _I64, _I64, _P, _INT, _P, _P]),
    "slk_scale_axis_f64": (_INT, [_P, _I64, _I64, _I64, _P, _INT, _P, _P]),
    "slk_row_noclip_scale_f32": (_INT, [_P, _I64, _I64, _D, _D, _P, _P]),
    "slk_row_noclip_scale_f64": (_INT, [_P, _I64, _I64, _D, _D, _P, _P]),
    "slk_row_rms_scale_f32": (_INT, [_P, _I64, _I64, _P, _P]),
    "slk_row_rms_scale_f64": (_INT, [_P, _I64, _I64, _P, _P]),
    "slk_scale_search_f32": (_INT, [_P, _I64, _I64, _CB, _P, _I32, _P, _I32, _P, _P, _P, _P]),
    "slk_hweighted_error_ws_bytes": (_SZ, [_I64, _I64, _I32]),
    "slk_hweighted_error_f32": (_INT, [_P, _P, _P, _I64, _I64, _P, _SZ, _P, _P]),
    "slk_hweighted_error_f64": (_INT, [_P, _P, _P, _I64, _I64, _P, _SZ, _P, _P]),
    "slk_mean_f32": (_INT, [_P, _I64, _P, _P]),
    "slk_mean_f64": (_INT, [_P, _I64, _P, _P]),
    "slk_gain_f32": (_INT, [_P, _P, _P, _P, _I64, _I64, _P, _P]),
    "slk_gain_f64": (_INT, [_P, _P, _P, _P, _I64, _I64, _P, _P]),
    "slk_scale_search_fullh_ws_bytes": (_SZ, [_I64, _I64, _I32, _I32]),
    "slk_scale_search_fullh_f32": (_INT, [_P, _I64, _I64, _CB, _P, _I32, _P, _I32, _P, _SZ, _P, _P, _P]),
    "slk_scale_search_fullh_checked_f32": (_INT, [_P, _I64, _I64, _CB, _P, _I32, _P, _I32, _P, _SZ, _P, _P, _P, _P]),
    "slk_hessian_accum_ws_bytes": (_SZ, [_I64, _I64]),
    "slk_hessian_accum_f32": (_INT, [_P, _I64, _I64, _I64, _P, _P, _D, _D, _P, _SZ, _P]),
    "slk_remove_input_bias_f32": (_INT, [_P, _P, _I64, _P, _P]),
    "slk_remove_input_bias_f64": (_INT, [_P, _P, _I64, _P, _P]),
    "slk_damp_value_f32": (_INT, [_P, _I64, _D, _P, _P]),
    "slk_col_resid_sums_f32": (_INT, [_P, _I64, _I64, _CB, _INT, _P, _P]),
    "slk_order_keys": (_INT, [_P, _I64, _P, _P, _P, _P]),
    "slk_argsort_f64": (_INT, [_P, _I64, _P, _P]),
    "slk_pivot_order_ws_bytes": (_SZ, [_I64]),
    "slk_pivot_order_f64": (_INT, [_P, _I64, _P, _SZ, _P, _P]),
    "slk_permute_cols_f32": (_INT, [_P, _I64, _I64, _P, _INT, _P, _P]),
    "slk_scale_permute_cols_f32": (_INT, [_P, _I64, _I64, _P, _P, _INT, _P, _P]),
    "slk_hinv_ws_bytes": (_SZ, [_I64]),
    "slk_hinv_from_f32": (_INT, [_P, _I64, _P, _P, _P, _SZ, _P, _P, _P, _P]),
    "slk_hinv_from_f64": (_INT, [_P, _I64, _P, _SZ, _P, _P, _P, _P]),
    "slk_chol_factor_ws_bytes": (_SZ, [_I64]),
    "slk_chol_factor_f32": (_INT, [_P, _I64, _P, _P, _P, _SZ, _P, _P, _P, _P, _P, _P]),
    "slk_chol_factor_batched_f32": (_INT, [_I32, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "slk_chol_dist_gather_f32": (_INT, [_P, _I64, _P, _P, _P, _SZ, _P, _P]),
    "slk_chol_dist_factor": (_INT, [_I64, _P, _I32, _I32, _P, _P, _P]),
    "slk_chol_dist_export_f32": (_INT, [_I64, _P, _P, _P, _P, _P, _P]),
    "slk_peer_alloc": (_INT, [_SZ, _P, _P]),
    "slk_peer_open": (_INT, [_P, _P]),
    "slk_peer_close": (_INT, [_P]),
    "slk_peer_free": (_INT, [_P]),
    "slk_peer_allreduce_f32": (_INT, [_P, _I32, _I32, _I64, _P]),
    "slk_sym_pack_f32": (_INT, [_P, _I64, _I64, C.c_float, _P, _P]),
    "slk_sym_unpack_f32": (_INT, [_P, _I64, _I64, C.c_float, _P, _P]),
    "slk_gptq_sweep_r_ws_bytes": (_SZ, [_I64, _I64]),
    "slk_codebook_breaks_host": (_INT, [_CB, _P]),
    "slk_stream_create": (_INT, [_INT, _P]),
    "slk_stream_destroy": (_INT, [_P]),
    "slk_set_option": (_INT, [C.c_char_p, _I64]),
    "slk_debug_timestamp": (_INT, [_P, _P]),
    "slk_debug_chol_trace": (_INT, [_P]),
    "slk_debug_sweep_trace": (_INT, [_P]),
    "slk_debug_scale_search_direct": (_INT, [_INT]),
    "slk_gptq_sweep_r_f32": (_INT, [_P, _P, _I64, _I64, _P, _P, _P, _P, _CB, _P, _SZ, _P]),
    "slk_gptq_sweep_r_err_f32": (_INT, [_P, _P, _I64, _I64, _P, _P, _P, _P, _CB, _P, _SZ, _P, _P]),
    "slk_sweep_error_f32": (_INT, [_P, _P, _P, _I64, _P, _P, _P]),
    "slk_upload_symmetric_f32": (_INT, [_P, _P, _I64, _I64, _P]),
    "slk_upload_symmetric_bytes": (_SZ, [_I64, _I64]),
    "slk_upload_symmetric_copy_f32": (_INT, [_P, _P, _I64, _I64, _P]),
    "slk_mirror_symmetric_f32": (_INT, [_P, _I64, _I64, _P]),
    "slk_gptq_sweep_f32": (_INT, [_P, _P, _I64, _I64, _P, _P, _CB, _I32, _I32, _I32, _P]),
    "slk_local_search_ws_bytes": (_SZ, [_I64, _I64]),
    "slk_local_search_f32": (_INT, [_P, _P, _P, _I64, _I64, _CB, _I32, _P, _SZ, _P]),
    "slk_local_search_err_f32": (_INT, [_P, _P, _P, _I64, _I64, _CB, _I32, _P, _SZ, _P, _P]),
    "slk_local_search_step_f32": (_INT, [_P, _P, _P, _I64, _I64, _CB, _I32, _P, _SZ, _I32, _P]),
    "slk_row_wsq_f32": (_INT, [_P, _P, _I64, _I64, _P, _P]),
    "slk_row_wsq_f64": (_INT, [_P, _P, _I64, _I64, _P, _P]),
    "slk_row_wsq_f32_h64": (_INT, [_P, _P, _I64, _I64, _P, _P]),
    "slk_bias_delta_f32": (_INT, [_P, _P, _P, _I64, _I64, _P, _P]),
    "slk_selftest_fastdiv_f32": (_INT, [_P, _I32, _P, _P]),
    "slk_tc_gemm_ws_bytes": (_SZ, [_I64, _I64, _I64]),
    "slk_tc_gemm_f32": (_INT, [_I32, _P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I64, _I64, _I64, C.c_float, C.c_float,
                               C.c_float, _P, _SZ, _P, _P]),
}


def header_symbols():
    """Every function name declared in include/sleekit_b200.h."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slk_[a-z0-9_]+)\s*\(", text)))


class SleekitLibError(RuntimeError):
    pass


_lib = None


def load(build_if_missing=True):
    """Load (building first if the .so is absent) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise SleekitLibError(f"{LIB_PATH} is missing; run `python -m sleekit_b200.build`")
        from . import build as _build

        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.slk_abi_version() != 1:
        raise SleekitLibError("libsleekit_b200.so ABI version mismatch; rebuild with --force")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().slk_last_error()
        raise SleekitLibError(f"sleekit_b200 C-ABI call failed ({rc}): {msg.decode() if msg else ''}")


def call(name, *args):
    """Invoke an int-returning entry point and raise on a non-zero status."""
    check(getattr(load(), name)(*args))
