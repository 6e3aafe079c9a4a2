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
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/wipa.h but not exported by libwipa.so"


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
    monkeypatch.setattr(_lib, "_lib", None)
    try:
        _lib.lib()
    except ImportError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")
