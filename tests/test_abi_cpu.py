"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol include/vc_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vc_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from viewcrafter_b200 import _lib
    assert _declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_loads_and_exports_every_symbol():
    from viewcrafter_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    assert lib.vc_abi_version() == _lib.ABI_VERSION
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_errors_are_reported_not_swallowed():
    """Argument validation happens on the host before any launch, so it is checkable without a GPU."""
    import ctypes as C
    from viewcrafter_b200 import _lib
    lib = _lib.load()
    d = _lib.GemmDesc()
    rc = lib.vc_gemm_tap(C.byref(d), None)
    assert rc != 0 and b"null" in lib.vc_last_error()
    with pytest.raises(_lib.VcError):
        _lib.check(rc, "vc_gemm_tap")
