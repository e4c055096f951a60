"""ctypes binding of libxnv2_b200.so (the C ABI declared in include/xnv2_b200.h).

The library is the product; there is no Python/CPU fallback.  ``load()`` raises if the
shared object is missing (build it with ``python -m on_device_image_captioning_b200.build``
or ``__graft_entry__.build()``), and ``xn_create`` fails on a machine without a B200.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libxnv2_b200.so")

XN_PREC_FP32, XN_PREC_BF16, XN_PREC_FP16 = 0, 1, 2
XN_DTYPE_F32, XN_DTYPE_I64 = 0, 1


class XnConfig(C.Structure):
    _fields_ = [
        ("has_swin", C.c_int32), ("img_size", C.c_int32), ("patch_size", C.c_int32), ("in_chans", C.c_int32),
        ("embed_dim", C.c_int32), ("n_stages", C.c_int32), ("depths", C.c_int32 * 4), ("swin_heads", C.c_int32 * 4),
        ("window_size", C.c_int32), ("mlp_ratio", C.c_float), ("feat_dim", C.c_int32), ("d_model", C.c_int32),
        ("n_enc", C.c_int32), ("n_dec", C.c_int32), ("ff", C.c_int32), ("num_heads", C.c_int32),
        ("n_exp_groups", C.c_int32), ("exp_groups", C.c_int32 * 8), ("num_exp_dec", C.c_int32),
        ("vocab", C.c_int32), ("max_seq_len", C.c_int32), ("enc_len", C.c_int32),
    ]


_P = C.c_void_p
_I = C.c_int
_SIGNATURES = {
    "xn_create": (C.c_int, [C.POINTER(XnConfig), _I, C.POINTER(_P)]),
    "xn_destroy": (C.c_int, [_P]),
    "xn_last_error": (C.c_char_p, [_P]),
    "xn_load_tensor": (C.c_int, [_P, C.c_char_p, _P, _I, C.POINTER(C.c_int64), _I]),
    "xn_finalize_weights": (C.c_int, [_P, _I]),
    "xn_forward_swin": (C.c_int, [_P, _P, _I, _P, _P]),
    "xn_forward_enc": (C.c_int, [_P, _P, _I, _P, _P, _P]),
    "xn_forward_dec": (C.c_int, [_P, _P, _I, _P, _P, _I, _P, _I, _P, _P]),
    "xn_beam_search": (C.c_int, [_P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "xn_beam_search_from_enc": (C.c_int, [_P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "xn_ensemble_beam_search": (C.c_int, [C.POINTER(_P), _I, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "xn_caption_host": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "xn_preprocess_rgb8": (C.c_int, [_P, _P, _I, _I, _I, _P, _I, _P]),
    "xn_preprocess_rgb8_batch": (C.c_int, [_P, C.POINTER(_P), _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _I, _P, _I, _P]),
    "xn_jpeg_available": (C.c_int, []),
    "xn_preprocess_jpeg_batch": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64), _I, _P, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "xn_caption_host_begin": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "xn_caption_host_end": (C.c_int, [_P, _I]),
    "xn_beam_search_sample": (C.c_int, [_P, _P, _I, _P, _I, _I, _I, _I, _I, C.c_uint64, _P, _P, _P, _P]),
    "xn_sample": (C.c_int, [_P, _P, _I, _P, _I, _I, _I, _I, C.c_uint64, _P, _P, _P, _P]),
    "xn_overflow_flag": (C.c_int, [_P, C.POINTER(C.c_int), _I]),
    "xn_kernel_launches": (C.c_int64, [_P]),
    "xn_workspace_bytes": (C.c_int64, [_P]),
    "xn_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "xn_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "xn_profile_read_min": (C.c_int, [_P, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "xn_profile_kernels": (C.c_int, [_P, C.c_char_p, _I]),
    "xn_mega_timeline": (C.c_int, [_P, C.POINTER(C.c_uint64), _I]),
    "xn_op_layernorm": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _P]),
    "xn_op_linear": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "xn_op_linear_skinny": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "xn_op_gemm_raw": (C.c_int, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "xn_op_window_attention": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "xn_op_logsoftmax_topk": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library has not been built "
                "(run `python -m on_device_image_captioning_b200.build`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
