"""ctypes binding of libwipa.so (the C ABI declared in include/wipa.h).

There is no fallback: if the shared library is missing or a call fails, this raises.  PyTorch is used by
the callers only to own device memory and streams; no torch type crosses this boundary.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwipa.so")

DTYPE_F32, DTYPE_BF16 = 0, 1
INFO_WORKSPACE_BYTES, INFO_CROSSKV_BYTES, INFO_DECODE_STEPS, INFO_XATTN_LATENT = 0, 1, 2, 3


class WipaError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str):
        super().__init__(f"libwipa: {what} failed with {code} ({detail})")
        self.code = code


class Arch(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("d_model", "enc_layers", "dec_layers", "heads", "ffn", "n_mels", "vocab", "dtype")]


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


class DecodeOpts(C.Structure):
    _fields_ = [("prompt", C.POINTER(C.c_int32)), ("prompt_len", C.c_int32), ("max_new", C.c_int32),
                ("eot", C.c_int32), ("suppress", C.POINTER(C.c_int32)), ("n_suppress", C.c_int32),
                ("begin_suppress", C.POINTER(C.c_int32)), ("n_begin_suppress", C.c_int32)]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
# name -> (restype, argtypes); must list every symbol include/wipa.h declares (tests/test_abi.py checks it)
PROTOTYPES = {
    "wipa_ctx_create": (_i, [C.POINTER(Arch), _i, _i, C.POINTER(_vp)]),
    "wipa_ctx_load_weights": (_i, [_vp, C.POINTER(TensorDesc), _i, _vp]),
    "wipa_ctx_destroy": (_i, [_vp]),
    "wipa_logmel": (_i, [_vp, _i, _i, _vp, _vp]),
    "wipa_encode": (_i, [_vp, _vp, _i, _vp, _vp]),
    "wipa_set_audio_features": (_i, [_vp, _vp, _i, _vp]),
    "wipa_decode_greedy": (_i, [_vp, _i, C.POINTER(DecodeOpts), _vp, _vp, _vp]),
    "wipa_decode_beam": (_i, [_vp, _i, _i, _f, C.POINTER(DecodeOpts), _vp, _vp, _vp]),
    "wipa_decode_logits": (_i, [_vp, _i, C.POINTER(C.c_int32), _i, _vp, _vp]),
    "wipa_per_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "wipa_pfer_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "wipa_strerror": (C.c_char_p, [_i]),
    "wipa_last_error": (C.c_char_p, []),
    "wipa_launch_count": (_i64, [_i]),
    "wipa_ctx_get_info": (_i, [_vp, _i, C.POINTER(_i64)]),
    "wipa_test_gemm_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wipa_test_gemm_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "wipa_test_gemm_epilogue": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "wipa_test_cross_attn_latent": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "wipa_test_gemm_rows": (_i, [_vp, _i, C.c_longlong, _i, C.c_longlong, _i, _vp, _vp, _i, _i, _i, _vp]),
    "wipa_test_cross_attn": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "wipa_test_enc_attention": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wipa_test_self_attn": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "wipa_test_enc_attention_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libwipa.so (built in-tree by __graft_entry__.build()).  Raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        l = lib()
        raise WipaError(rc, what, f"{l.wipa_strerror(rc).decode()}: {l.wipa_last_error().decode()}")


def launch_count(reset: bool = False) -> int:
    return int(lib().wipa_launch_count(1 if reset else 0))
