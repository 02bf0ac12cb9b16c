"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares (no GPU needed)."""
import ctypes, glob, os, re
from conftest import ROOT


def declared_in_headers():
    names = set()
    for h in glob.glob(os.path.join(ROOT, 'include', '*.h')):
        src = re.sub(r'/\*.*?\*/', '', open(h).read(), flags=re.S)
        names.update(re.findall(r'\b(gn_[a-z0-9_]+)\s*\(', src))
    names.discard('gn_stream_t')
    return sorted(names)


def test_library_builds_and_exports_all_declared_symbols():
    from gridnext_b200.build import build
    path = build()
    lib = ctypes.CDLL(path)
    names = declared_in_headers()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    from gridnext_b200 import _lib
    assert set(_lib.declared_symbols()) <= set(names) | {'gn_last_error'}, set(_lib.declared_symbols()) - set(names)
    assert lib.gn_version() >= 100


def test_missing_library_fails_loudly(monkeypatch):
    from gridnext_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libgridnext_b200.so')
    import pytest
    with pytest.raises(RuntimeError):
        _lib.load()


def test_cpu_tensors_are_rejected_not_silently_computed():
    import pytest, torch
    from gridnext_b200.hexagdly import Conv2d
    m = Conv2d(3, 4)
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 3, 8, 8))
