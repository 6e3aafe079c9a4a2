"""ctypes binding of libwipa.so (the C ABI declared in include/wipa.h).

There is no fallback: if the shared library is missing or a call fails, this raises.  PyTorch is used by
the callers only to own device memory and streams; no torch type crosses this boundary.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwipa.so")                 # 16-bit path in IEEE fp16 (and the fp32 path)
LIB_PATH_BF16 = os.path.join(_HERE, "libwipa_bf16.so")        # the same sources compiled with bfloat16 as the 16-bit type

DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
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
    "wipa_resample_pcm16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp]),
    "wipa_encode": (_i, [_vp, _vp, _i, _vp, _vp]),
    "wipa_set_audio_features": (_i, [_vp, _vp, _i, _vp]),
    "wipa_decode_greedy": (_i, [_vp, _i, C.POINTER(DecodeOpts), _vp, _vp, _vp]),
    "wipa_decode_beam": (_i, [_vp, _i, _i, _f, C.POINTER(DecodeOpts), _vp, _vp, _vp]),
    "wipa_decode_begin": (_i, [_vp, _i, C.POINTER(C.c_int32), _i, _vp, _vp]),
    "wipa_decode_next": (_i, [_vp, _i, _vp, _vp, _vp]),
    "wipa_decode_logits": (_i, [_vp, _i, C.POINTER(C.c_int32), _i, _vp, _vp]),
    "wipa_per_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "wipa_pfer_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "wipa_strerror": (C.c_char_p, [_i]),
    "wipa_last_error": (C.c_char_p, []),
    "wipa_launch_count": (_i64, [_i]),
    "wipa_h16_dtype": (_i, []),
    "wipa_ctx_get_info": (_i, [_vp, _i, C.POINTER(_i64)]),
    "wipa_test_gemm_h16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wipa_test_gemm_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "wipa_test_gemm_epilogue": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "wipa_test_cross_attn_latent": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "wipa_test_lat_tiled_elems": (C.c_longlong, [_i, _i]),
    "wipa_test_lat_tile": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "wipa_test_xlq_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "wipa_test_gemm_rows": (_i, [_vp, _i, C.c_longlong, _i, C.c_longlong, _i, _vp, _vp, _i, _i, _i, _vp]),
    "wipa_test_logits_argmax": (_i, [_vp, _i, _vp]),
    "wipa_test_cross_attn": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "wipa_test_enc_attention": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wipa_test_self_attn": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "wipa_test_enc_attention_h16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
}

_libs = {}


def lib(h16: str = "f16") -> C.CDLL:
    """Load libwipa.so (h16="f16", the default build) or libwipa_bf16.so (h16="bf16"), both built in-tree by
    __graft_entry__.build().  Raises if the file is absent: there is no CPU fallback."""
    if h16 not in ("f16", "bf16"):
        raise ValueError(f"h16 must be 'f16' or 'bf16', got {h16!r}")
    handle = _libs.get(h16)
    if handle is None:
        path = LIB_PATH if h16 == "f16" else LIB_PATH_BF16
        if h16 == "f16" and os.environ.get("WIPA_LIBWIPA"):          # development: an experimental build of the same ABI
            path = os.environ["WIPA_LIBWIPA"]
        if not os.path.exists(path):
            raise ImportError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        handle = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        want = DTYPE_F16 if h16 == "f16" else DTYPE_BF16
        if handle.wipa_h16_dtype() != want:
            raise ImportError(f"{path} was built for 16-bit dtype {handle.wipa_h16_dtype()}, expected {want}")
        _libs[h16] = handle
    return handle


def lib_for_dtype(dtype_code: int) -> C.CDLL:
    return lib("bf16" if dtype_code == DTYPE_BF16 else "f16")


def torch_h16(h16: str = "f16"):
    """torch dtype of the 16-bit element type the named build computes in (for the standalone kernel entry points)."""
    import torch
    return torch.float16 if h16 == "f16" else torch.bfloat16


def check(rc: int, what: str, handle: C.CDLL = None) -> None:
    if rc != 0:
        l = handle if handle is not None else lib()
        raise WipaError(rc, what, f"{l.wipa_strerror(rc).decode()}: {l.wipa_last_error().decode()}")


def launch_count(reset: bool = False) -> int:
    """Kernels launched by every loaded build of the library since the last reset."""
    if not _libs:
        lib()
    return sum(int(h.wipa_launch_count(1 if reset else 0)) for h in _libs.values())
