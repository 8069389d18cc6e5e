"""ctypes binding of libclipn.so (C ABI declared in include/clipn.h).

The product path FAILS LOUDLY when the CUDA library is missing or a call errors: there is no
CPU / PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclipn.so")

EPI_STORE, EPI_BIAS_GELU, EPI_BIAS_RESID, EPI_DGELU, EPI_ACCUM_F32, EPI_STORE_F32, EPI_LSE, EPI_CLIP_DLOGITS, \
    EPI_SIGLIP, EPI_BIAS_GELU_GRAD, EPI_MUL_AUX = range(11)


class GemmDesc(C.Structure):
    """Mirror of `struct clipn_gemm_desc` (include/clipn.h)."""
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_int64), ("a_mn_major", C.c_int32),
        ("b", C.c_void_p), ("ldb", C.c_int64), ("b_mn_major", C.c_int32),
        ("c", C.c_void_p), ("ldc", C.c_int64),
        ("c2", C.c_void_p), ("ldc2", C.c_int64),
        ("bias", C.c_void_p),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("epilogue", C.c_int32),
        ("alpha", C.c_float),
        ("splits", C.c_int32),
        ("row_lse", C.c_void_p), ("col_lse", C.c_void_p),
        ("part_max", C.c_void_p), ("part_sum", C.c_void_p),
        ("pos", C.c_void_p), ("scalar_acc", C.c_void_p),
        ("logit_bias", C.c_float), ("gscale", C.c_float), ("col_w", C.c_float),
        ("label_offset", C.c_int32), ("negative_only", C.c_int32),
        ("alpha_dev", C.c_void_p), ("logit_bias_dev", C.c_void_p),
        ("col_sum", C.c_void_p),
    ]


class AdamTensor(C.Structure):
    """Mirror of `struct clipn_adamw_tensor` (include/clipn.h)."""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64), ("lr", C.c_float), ("weight_decay", C.c_float), ("is_bf16", C.c_int32)]


_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/clipn.h declares
SIGNATURES = {
    "clipn_version": (C.c_int, []),
    "clipn_last_error": (C.c_char_p, []),
    "clipn_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "clipn_gemm": (C.c_int, [C.POINTER(GemmDesc), _P]),
    "clipn_gemm_ref": (C.c_int, [C.POINTER(GemmDesc), _P]),
    "clipn_gemm_tile_n": (C.c_int, [C.c_int]),
    "clipn_layernorm_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _F, _P]),
    "clipn_layernorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "clipn_attention_fwd": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _F, _P]),
    "clipn_attention_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, _P]),
    "clipn_patchify": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    "clipn_patchify_padded": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _P]),
    "clipn_accum_rows_f32": (C.c_int, [_P, _I64, _P, _I64, _I64, _I32, _P]),
    "clipn_vision_embed_fwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _P]),
    "clipn_vision_embed_bwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _P]),
    "clipn_text_embed_fwd": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "clipn_text_embed_bwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "clipn_gather_rows": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P]),
    "clipn_scatter_rows": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P]),
    "clipn_l2norm_fwd": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "clipn_l2norm_bwd": (C.c_int, [_P, _I32, _P, _P, _P, _I64, _I32, _P]),
    "clipn_colsum": (C.c_int, [_P, _I64, _P, _I64, _I32, _P]),
    "clipn_cast_f32_to_bf16": (C.c_int, [_P, _P, _I64, _P]),
    "clipn_adamw_multi": (C.c_int, [C.POINTER(AdamTensor), _I32, _F, _F, _F, _F, _F, _P]),
    "clipn_peer_gemm_tile_n": (_I32, [_I32, _I32, _I32]),
    "clipn_peer_gather": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _I32, _I32, _I32, _P, _P, _P]),
    "clipn_clip_fwd_fused_workspace": (C.c_int64, [_I32, _I32, _I32]),
    "clipn_stage_timing": (C.c_int, [_I32]),
    "clipn_stage_times": (_I32, [C.POINTER(C.c_float)]),
    "clipn_clip_fwd_fused": (C.c_int, [_P, _P, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _I32, _I32, _I32, _I32, _F,
                                       _P, _P, _P, _P, _P, _P, _P]),
    "clipn_siglip_fwd_fused": (C.c_int, [_P, _P, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _I32, _I32, _I32, _I32,
                                         _P, _P, _F, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "clipn_clip_lse_workspace": (C.c_int64, [_I32, _I32]),
    "clipn_clip_lse_fwd": (C.c_int, [_P, _P, _I32, _I32, _I32, _F, _P, _I32, _P, _P, _P, _P]),
    "clipn_clip_dlogits": (C.c_int, [_P, _P, _I32, _I32, _I32, _F, _P, _I32, _P, _P, _F, _F, _P, _I64, _P, _P, _P]),
    "clipn_clip_dfeat": (C.c_int, [_P, _I64, _P, _I32, _I32, _I32, _F, _P, _P, _I32, _P]),
}

_lib = None


class ClipnError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libclipn.so (once). Raises ClipnError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ClipnError(
                f"{LIB_PATH} is missing: build it with `python -m open_clip_b200.build` "
                "(or __graft_entry__.build()). open_clip_b200 has no CPU/PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise ClipnError("libclipn call failed (%d): %s" % (rc, lib().clipn_last_error().decode()))
