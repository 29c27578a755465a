"""ctypes binding of libtfswa_b200.so (the C ABI declared in include/tfswa_b200.h).

The library is loaded lazily; a missing library or a failing call raises ``RuntimeError`` - there is
deliberately no fallback implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtfswa_b200.so")

F32, BF16 = 0, 1
PRO_NONE, PRO_LNHAT, PRO_GELU, PRO_AFFINE = 0, 1, 2, 4
EPI_NONE, EPI_GELU, EPI_MUL_DGELU = 0, 1, 2
GEOM_TSA, GEOM_FSA, GEOM_SWA = 0, 1, 2

_i64, _i32, _p, _f = C.c_int64, C.c_int32, C.c_void_p, C.c_float


class LinearArgs(C.Structure):
    _fields_ = [("x", _p), ("ldx", _i64), ("x_bs", _i64),
                ("w", _p), ("w_bs", _i64),
                ("bias", _p), ("bias_bs", _i64),
                ("row_stats", _p), ("rs_bs", _i64),
                ("in_scale", _p), ("in_shift", _p),
                ("r1", _p), ("ldr1", _i64), ("r1_bs", _i64),
                ("r2", _p), ("ldr2", _i64), ("r2_bs", _i64),
                ("y", _p), ("ldy", _i64), ("y_bs", _i64),
                ("pre", _p), ("ldpre", _i64), ("pre_bs", _i64),
                ("col_stats", _p),
                ("M", _i64), ("N", _i32), ("K", _i32),
                ("prologue", _i32), ("epilogue", _i32), ("batch", _i32), ("dtype", _i32)]


class TailArgs(C.Structure):
    _fields_ = [("att", _p), ("lda", _i64), ("att_bs", _i64),
                ("res", _p), ("ldr", _i64), ("res_bs", _i64),
                ("wp", _p), ("w1", _p), ("w2", _p),
                ("bp", _p), ("b1", _p), ("b2", _p),
                ("out", _p), ("ldo", _i64), ("out_bs", _i64),
                ("M", _i64), ("C", _i32), ("hidden", _i32), ("batch", _i32), ("eps", C.c_float)]


class HeadArgs(C.Structure):
    _fields_ = [("x", _p), ("ldx", _i64), ("wi", _p), ("wq", _p), ("bi", _p), ("bq", _p),
                ("x1", _p), ("ld1", _i64), ("qkv", _p), ("ldq", _i64), ("M", _i64), ("C", _i32), ("eps", C.c_float)]


class AttnArgs(C.Structure):
    _fields_ = [("qkv", _p), ("ldq", _i64), ("out", _p), ("ldo", _i64), ("lse", _p),
                ("pad_kv", _p), ("rel_bias", _p),
                ("B", _i32), ("H", _i32), ("W", _i32), ("C", _i32), ("heads", _i32),
                ("geom", _i32), ("ws", _i32), ("shift", _i32), ("use_shift_mask", _i32), ("dtype", _i32),
                ("flags", _i32)]


class ConvArgs(C.Structure):
    _fields_ = [("x", _p), ("y", _p), ("pre", _p), ("w", _p), ("bias", _p), ("col_stats", _p),
                ("B", _i32), ("Hin", _i32), ("Win", _i32), ("Cin", _i32), ("Hout", _i32), ("Wout", _i32), ("Cout", _i32),
                ("kind", _i32), ("epilogue", _i32), ("dtype", _i32)]


# name -> (restype, argtypes); every symbol include/tfswa_b200.h declares
SIGNATURES = {
    "tfswa_last_error": (C.c_char_p, []),
    "tfswa_version": (C.c_char_p, []),
    "tfswa_device_supported": (C.c_int, []),
    "tfswa_linear_fwd": (C.c_int, [C.POINTER(LinearArgs), _p]),
    "tfswa_linear_tc_fwd": (C.c_int, [C.POINTER(LinearArgs), _p, _p, _p]),
    "tfswa_branch_tail_tc_fwd": (C.c_int, [C.POINTER(TailArgs), _p]),
    "tfswa_block_head_tc_fwd": (C.c_int, [C.POINTER(HeadArgs), _p]),
    "tfswa_row_stats": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _i32, _i32, _i32, _p]),
    "tfswa_attn_fwd": (C.c_int, [C.POINTER(AttnArgs), _p]),
    "tfswa_attn_tc_scratch_bytes": (C.c_int64, [C.POINTER(AttnArgs)]),
    "tfswa_attn_tc_fwd": (C.c_int, [C.POINTER(AttnArgs), _p, _i64, _p]),
    "tfswa_attn_win_tc_fwd": (C.c_int, [C.POINTER(AttnArgs), _p]),
    "tfswa_attn_win_tc_interior_launches": (C.c_longlong, []),
    "tfswa_conv_fwd": (C.c_int, [C.POINTER(ConvArgs), _p]),
    "tfswa_conv_tc_fwd": (C.c_int, [C.POINTER(ConvArgs), _p, _p]),
    "tfswa_stem_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "tfswa_head_tail_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "tfswa_bn_finalize": (C.c_int, [_p, _i64, _p, _p, _p, _p, _f, _f, _p, _p, _p, _i32, _p]),
    "tfswa_affine_act": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "tfswa_bilinear_fwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    # backward
    "tfswa_linear_wgrad": (C.c_int, [C.POINTER(LinearArgs), _p, _i64, _i64, _p, _p, _p]),
    "tfswa_conv_wgrad": (C.c_int, [C.POINTER(ConvArgs), _p, _p, _p, _p]),
    "tfswa_act_bwd": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "tfswa_affine_act_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "tfswa_lnhat_bwd": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _p, _i64, _i64, _i64, _i32, _i32, _i32, _p]),
    "tfswa_sum_batch": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p]),
    "tfswa_bilinear_bwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "tfswa_attn_bwd": (C.c_int, [C.POINTER(AttnArgs), _p, _p, _p, _p, _p]),
    "tfswa_stem_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "tfswa_head_tail_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    # spectrogram steps around the model (overlap-add separation)
    "tfswa_spec_pack_norm": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _f, _i32, _p]),
    "tfswa_spec_mask_apply": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "tfswa_ola_add": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _i32, _i32, _i64, _i64, _i64, _p]),
    "tfswa_mrstft_mag_loss": (C.c_int, [_p, _p, _i64, _f, _f, _f, _p, _p, _p]),
    # optimiser step over the flat arena
    "tfswa_grad_sumsq": (C.c_int, [_p, _i64, _p, _p]),
    "tfswa_adamw_clip_step": (C.c_int, [_p, _p, _p, _p, _i64, _p, _p, _f, _f, _f, _f, _f, _f, _f, _i64, _p, _p]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). tfswa_unet_b200 has no CPU or PyTorch fallback.")
        # TFSWA_B200_LIB: developer A/B switch (an alternative build of the SAME library, e.g. another tile constant)
        l = C.CDLL(os.environ.get("TFSWA_B200_LIB") or LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)   # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().tfswa_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"tfswa_b200 {what} failed (rc={rc}): {msg}")
