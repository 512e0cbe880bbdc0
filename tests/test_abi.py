"""The C-ABI library loads and exports every symbol declared in include/xclim_b200.h (no GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "xclim_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from xclim_b200 import _build, _lib
    if not os.path.exists(_build.LIB_PATH):
        _build.build()
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    from xclim_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/xclim_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in xclim_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(syms)


def test_version_and_error_text(lib):
    assert lib.xc_version() == 100
    assert isinstance(lib.xc_last_error(), bytes)


def test_operator_validation_matches_reference_messages():
    # indices/generic.py:285, 296
    from xclim_b200 import _lib
    with pytest.raises(ValueError, match="not recognized"):
        _lib.op_code("=>")
    with pytest.raises(ValueError, match="not permitted"):
        _lib.op_code("==", constrain=(">", "<"))
    assert _lib.op_code("lt") == _lib.op_code("<")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from xclim_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.XclimB200Error, match="no CPU fallback"):
        _lib.load()
