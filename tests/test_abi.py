"""The C-ABI shared library loads (no GPU needed) and exports every symbol include/wipa.h declares."""
import ctypes
import os
import re

from whisper_ipa_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wipa.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wipa_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    names = declared_symbols()
    assert len(names) >= 19
    for path in (_lib.LIB_PATH, _lib.LIB_PATH_BF16):               # both builds: fp16 (default) and bfloat16
        handle = ctypes.CDLL(path)
        for n in names:
            assert hasattr(handle, n), f"{n} declared in include/wipa.h but not exported by {os.path.basename(path)}"


def test_each_build_reports_its_16_bit_type(built_lib):
    assert _lib.lib("f16").wipa_h16_dtype() == _lib.DTYPE_F16
    assert _lib.lib("bf16").wipa_h16_dtype() == _lib.DTYPE_BF16


def test_binding_covers_header(built_lib):
    assert sorted(_lib.PROTOTYPES) == declared_symbols()


def test_error_strings(built_lib):
    assert built_lib.wipa_strerror(0) == b"ok"
    assert built_lib.wipa_strerror(-4) == b"call out of order"
    assert built_lib.wipa_strerror(-99) == b"unknown error"
    assert isinstance(_lib.launch_count(), int)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Arch) == 8 * 4
    assert ctypes.sizeof(_lib.TensorDesc) == 24
    assert ctypes.sizeof(_lib.DecodeOpts) == 56      # ptr, 3 x i32 (+pad), ptr, i32 (+pad), ptr, i32 (+pad)
    assert _lib.DecodeOpts.suppress.offset == 24 and _lib.DecodeOpts.begin_suppress.offset == 40


def test_missing_library_is_loud(monkeypatch, tmp_path):
    import importlib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_libs", {})
    try:
        _lib.lib()
    except ImportError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")
