"""The C-ABI shared library loads (no GPU needed) and exports every function include/drpo_b200.h declares; the ctypes binding
table names exactly those functions."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "drpo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drpo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from drpo_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/drpo_b200.h but not exported by {_lib.LIB_PATH}"
    assert sorted(s[0] for s in _lib.SYMBOLS) == names
    assert _lib.load().drpo_abi_version() == 2


def test_missing_library_fails_loudly(monkeypatch):
    from drpo_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdrpo_sm100.so")
    import pytest
    with pytest.raises(RuntimeError, match="no CPU or eager fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected():
    import pytest
    import torch
    from drpo_b200 import _lib
    with pytest.raises(RuntimeError, match="no CPU path"):
        _lib.ptr(torch.zeros(4))
