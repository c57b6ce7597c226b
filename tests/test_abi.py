"""The C-ABI library must load without a GPU and export every symbol include/tvts_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tvts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tvts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tvts_b200 import _lib
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.tvts_version() == 100


def test_argument_errors_are_reported_not_crashing():
    from tvts_b200 import _lib
    lib = _lib.lib()
    g = _lib.GemmArgs()
    assert lib.tvts_gemm(ctypes.byref(g), None) < 0
    assert b"empty problem" in lib.tvts_last_error()
