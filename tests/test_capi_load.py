"""CPU-only: the C-ABI library builds, loads and exports every symbol include/pfc.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pfc.h")).read()
    return sorted(set(re.findall(r"\b(pfc_[a-z0-9_]+)\s*\(", txt)) - {"pfc_ctx"})


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from pfc_b200 import capi
    L = ctypes.CDLL(capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} declared in pfc.h but not exported"
    assert sorted(capi.SYMBOLS) == declared
    assert b"sm_100a" in capi.lib().pfc_version()


def test_error_convention_without_gpu():
    from pfc_b200 import capi
    L = capi.lib()
    assert L.pfc_sync(None) < 0
    assert b"NULL" in L.pfc_last_error()
    assert L.pfc_set_debug(None, 1) < 0


def test_missing_library_fails_loudly(monkeypatch):
    from pfc_b200 import capi
    monkeypatch.setattr(capi, "_LIB", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/libpfc_b200.so")
    with pytest.raises(ImportError):
        capi.lib()
