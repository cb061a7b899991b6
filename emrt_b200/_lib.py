"""ctypes binding of libemrt_b200.so (the C ABI declared in include/emrt_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libemrt_b200.so")

# enums (mirror include/emrt_b200.h)
F32, BF16, F16, I32, U8 = 0, 1, 2, 3, 4
LOC_NORMALIZED, LOC_PIXEL_OFFSET, VALUE_HEAD_MAJOR, QUERY_PIXEL_GRID = 0, 1, 2, 4
EPI_NONE, EPI_ROW_MASK, EPI_RELU, EPI_RESIDUAL_LN, EPI_MSDA_QPROJ, EPI_HEAD_MAJOR = 0, 1, 2, 4, 8, 16
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2
ERR_UNSUPPORTED = -3


class EmrtError(RuntimeError):
    pass


class GnBranch(C.Structure):
    _fields_ = [("conv", C.c_void_p), ("skip", C.c_void_p), ("stats", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("L", C.c_int32), ("groups", C.c_int32), ("Lv", C.c_int32), ("eps", C.c_float), ("shapes_hw", C.c_int32 * 16)]


class LinearArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("y", C.c_void_p),
        ("rows", C.c_int64), ("K", C.c_int32), ("N", C.c_int32),
        ("x_dtype", C.c_int32), ("w_dtype", C.c_int32), ("y_dtype", C.c_int32), ("w_transposed", C.c_int32),
        ("epilogue", C.c_int32),
        ("row_scale", C.c_void_p),
        ("residual", C.c_void_p), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float),
        ("y2", C.c_void_p), ("qproj_group", C.c_int32),
        ("impl", C.c_int32),
        ("hm_rows", C.c_int32), ("hm_D", C.c_int32),
        ("x2", C.c_void_p), ("x2_period", C.c_int32),
        ("row_bias", C.c_void_p), ("row_bias_period", C.c_int32),
        ("gn", C.POINTER(GnBranch)),
        ("x_nchw_hw", C.c_int32),
    ]


class FfnArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float),
        ("y", C.c_void_p), ("rows", C.c_int64), ("d_model", C.c_int32), ("d_ff", C.c_int32),
        ("gn", C.POINTER(GnBranch)),
    ]


class MsdaArgs(C.Structure):
    _fields_ = [
        ("query", C.c_void_p), ("value", C.c_void_p), ("ref", C.c_void_p), ("ref_batches", C.c_int32),
        ("query_pos", C.c_void_p), ("query_pos_rows", C.c_int32), ("query_scratch", C.c_void_p), ("query_eff", C.c_void_p),
        ("value_mask", C.c_void_p),
        ("w_value", C.c_void_p), ("b_value", C.c_void_p), ("w_offsets", C.c_void_p), ("b_offsets", C.c_void_p),
        ("w_attn", C.c_void_p), ("b_attn", C.c_void_p), ("w_out", C.c_void_p), ("b_out", C.c_void_p),
        ("wv_packed", C.c_void_p), ("wq_packed", C.c_void_p), ("wo_packed", C.c_void_p), ("b_query", C.c_void_p),
        ("row_bias", C.c_void_p),
        ("residual", C.c_void_p), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float),
        ("out", C.c_void_p), ("workspace", C.c_void_p),
        ("B", C.c_int32), ("Lq", C.c_int32), ("Lv", C.c_int32), ("C", C.c_int32), ("M", C.c_int32), ("L", C.c_int32), ("P", C.c_int32),
        ("shapes_hw", C.c_int32 * 16),
        ("dtype", C.c_int32), ("flags", C.c_int32), ("keep_pixel_major", C.c_int32),
        ("window_center", C.POINTER(C.c_int32)),
        ("timing_events", C.c_void_p * 8),
        ("gather_start_event", C.c_void_p),
    ]


class MsdaGrads(C.Structure):
    _fields_ = [
        ("d_out", C.c_void_p), ("d_query", C.c_void_p), ("d_value", C.c_void_p), ("d_ref", C.c_void_p),
        ("dw_query", C.c_void_p), ("db_query", C.c_void_p), ("dw_value", C.c_void_p), ("db_value", C.c_void_p),
        ("dw_out", C.c_void_p), ("db_out", C.c_void_p),
        ("wq_cat", C.c_void_p), ("w_value_cast", C.c_void_p), ("w_out_cast", C.c_void_p),
        ("workspace", C.c_void_p),
    ]


_P, _I, _L = C.c_void_p, C.c_int, C.c_int64
_I32P = C.POINTER(C.c_int32)

# name -> (restype, argtypes); exactly the symbols include/emrt_b200.h declares
SIGNATURES = {
    "emrt_version": (C.c_int, []),
    "emrt_last_error": (C.c_char_p, []),
    "emrt_device_check": (C.c_int, []),
    "emrt_launch_count": (C.c_int64, []),
    "emrt_reset_launch_count": (None, []),
    "emrt_msda_gather_fwd": (C.c_int, [_P, _P, _P, _P, _L, _P, _I, _I, _I, _I, _I, _I, _I, _I32P, _I32P, _I, _I, _I, _P]),
    "emrt_msda_gather_fwd_hint": (C.c_int, [_P, _P, _P, _P, _L, _P, _I, _I, _I, _I, _I, _I, _I, _I32P, _I32P, _I, _I, _I, _I32P, _P]),
    "emrt_msda_gather_bwd": (C.c_int, [_P, _P, _P, _P, _P, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I32P, _I32P,
                                       _I, _I, _I, _P]),
    "emrt_msda_gather_bwd_hint": (C.c_int, [_P, _P, _P, _P, _P, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I32P, _I32P,
                                            _I, _I, _I, _I32P, _P]),
    "emrt_linear_fwd": (C.c_int, [C.POINTER(LinearArgs), _P]),
    "emrt_ffn_fused_fwd": (C.c_int, [C.POINTER(FfnArgs), _P]),
    "emrt_pack_weight": (C.c_int, [_P, _I, _P, _I, _I, _I, _P]),
    "emrt_linear_bwd_weight": (C.c_int, [_P, _P, _P, _P, _L, _I, _I, _I, _I, _P]),
    "emrt_msda_qproj_bwd": (C.c_int, [_P, _P, _P, _P, _L, _I, _I, _I, _I32P, _I, _I, _I, _P]),
    "emrt_msda_ref_bwd": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I32P, _I, _P]),
    "emrt_scale_rows_cast": (C.c_int, [_P, _P, _P, _L, _I, _I, _P]),
    "emrt_msda_softmax_loc": (C.c_int, [_P, _L, _P, _L, _P, _L, _P, _P, _I, _I, _I, _I, _I, _I32P, _I, _I, _P]),
    "emrt_add_layernorm": (C.c_int, [_P, _P, _P, _P, _P, _L, _I, C.c_float, _I, _P]),
    "emrt_residual_layernorm": (C.c_int, [_P, _P, _P, _P, _P, _P, _L, _I, C.c_float, _I, _P]),
    "emrt_pack_conv3x3_weight": (C.c_int, [_P, _P, _I, _I, _I, _P]),
    "emrt_conv3x3_tokens_fwd": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I32P, _I, _I, _I, _P]),
    "emrt_conv3x3_stats_workspace_floats": (C.c_longlong, [_I, _I, _I]),
    "emrt_conv3x3_tokens_stats_fwd": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I32P, _I, _P]),
    "emrt_conv3x3_tokens_stats_part_fwd": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I32P, _I, _I, _P]),
    "emrt_groupnorm_gelu_residual": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, C.c_float, _I32P, _I, _P]),
    "emrt_groupnorm_stats": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I32P, _I, _P]),
    "emrt_groupnorm_workspace_floats": (C.c_longlong, [_I, _I, _I]),
    "emrt_residual_layernorm_gn": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, C.c_float, C.c_float,
                                             _I32P, _I, _P]),
    "emrt_nchw_to_tokens": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "emrt_groupnorm_tokens": (C.c_int, [_P, _P, _P, _P, _L, _P, _I, _I, _I, _I, C.c_float, _I, _P]),
    "emrt_mha_small": (C.c_int, [_P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _I, _I, C.c_float, _I, _P]),
    "emrt_add_bcast": (C.c_int, [_P, _P, _P, _L, _L, _I, _P]),
    "emrt_upsample2x": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "emrt_window_accumulate": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "emrt_finalize_argmax": (C.c_int, [_P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "emrt_stitch_argmax_fused": (C.c_int, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "emrt_stitch_argmax_eval": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "emrt_calculate_area": (C.c_int, [_P, _P, _L, _I, _I, _P, _P]),
    "emrt_msda_fused_workspace_bytes": (C.c_int64, [_I, _I, _I, _I, _I, _I, _I, _I]),
    "emrt_msda_fused_fwd": (C.c_int, [C.POINTER(MsdaArgs), _P]),
    "emrt_msda_fused_bwd_workspace_bytes": (C.c_int64, [_I, _I, _I, _I, _I, _I, _I, _I]),
    "emrt_msda_fused_bwd": (C.c_int, [C.POINTER(MsdaArgs), C.POINTER(MsdaGrads), _P]),
    "emrt_layernorm_bwd_workspace_floats": (C.c_int64, [_L, _I]),
    "emrt_layernorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, C.c_float, _I, _P]),
    "emrt_groupnorm_bwd_workspace_floats": (C.c_int64, [_I, _I, _I, _I]),
    "emrt_groupnorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, C.c_float, _I32P, _I, _I, _P]),
    "emrt_relu_bwd": (C.c_int, [_P, _P, _P, _L, _I, _P]),
    "emrt_batch_sum": (C.c_int, [_P, _P, _I, _L, _I, _P]),
    "emrt_column_sum": (C.c_int, [_P, _P, _L, _I, _I, _P]),
    "emrt_sigmoid_fwd": (C.c_int, [_P, _P, _L, _P]),
    "emrt_sigmoid_bwd": (C.c_int, [_P, _P, _P, _L, _P]),
    "emrt_mha_small_bwd": (C.c_int, [_P, _L, _P, _L, _P, _L, _P, _P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _I, C.c_float, _I, _P]),
    "emrt_conv3x3_tokens_bwd_weight": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I32P, _I, _I, _P]),
}

_lib = None


def load():
    """Load the shared library (once) and bind every declared symbol.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EmrtError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(emrt_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        msg = load().emrt_last_error()
        err = EmrtError(f"emrt_b200 error {status}: {msg.decode() if msg else ''}")
        err.status = int(status)
        raise err


def i32_array(values):
    arr = (C.c_int32 * len(values))(*[int(v) for v in values])
    return arr
