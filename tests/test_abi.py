"""CPU tests: the C-ABI library builds, loads and exports every symbol include/crisper.h declares (no compute)."""
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "crisper.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from crisperwhisper_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libcrisper.so does not export {n}"
        assert n in _lib.EXPORTS, f"{n} has no ctypes signature in crisperwhisper_b200/_lib.py"
    assert lib.cw_abi_version() == 1


def test_product_never_imports_oracle():
    """The product package must not route through the oracle or any CPU fallback."""
    pkg = os.path.join(ROOT, "crisperwhisper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "from oracle" not in txt and "import oracle" not in txt, f


def test_missing_library_fails_loudly(monkeypatch):
    from crisperwhisper_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcrisper.so")
    try:
        _lib.load()
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must raise when the CUDA library is missing")
