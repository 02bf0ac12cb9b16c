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


def test_host_pattern_recognition_without_a_gpu():
    """Host logic that decides which module trees run on the fused kernels (no device work): the count-MLP pattern of
    Tutorial_visium_count cell 12, the hexagonal corrector of gridnet_models.py:128-148 and the Cartesian one of :51-66."""
    import torch.nn as nn
    from gridnext_b200.count_mlp import compile_count_mlp
    from gridnext_b200.corrector import parse_corrector
    from gridnext_b200.gridnet_models import GridNet, GridNetHexOddr
    mlp = nn.Sequential(nn.Linear(40, 500), nn.Linear(500, 100), nn.BatchNorm1d(100), nn.ReLU(), nn.Linear(100, 100), nn.Linear(100, 50),
                        nn.BatchNorm1d(50), nn.ReLU(), nn.Linear(50, 7))
    c = compile_count_mlp(mlp)
    assert c is not None and [s.lin.out_features for s in c.stages] == [500, 100, 100, 50, 7]
    assert [(s.bn is not None, s.relu) for s in c.stages] == [(False, False), (True, True), (False, False), (True, True), (False, False)]
    assert compile_count_mlp(nn.Sequential(nn.Linear(8, 8), nn.BatchNorm1d(8), nn.Linear(8, 3))) is None       # BN without ReLU
    assert compile_count_mlp(nn.Sequential(nn.Linear(8, 8), nn.ReLU(), nn.Linear(9, 3))) is None                # width mismatch
    hexnet = GridNetHexOddr(nn.Linear(4, 7), (4,), (78, 64), 7, use_bn=True)
    st = parse_corrector(hexnet.corrector)
    assert [(type(m).__name__, bn is not None, relu) for m, bn, relu in st] == [('Conv2d', False, False), ('Conv2d', False, False),
                                                                               ('Conv2d', True, True), ('Conv2d', False, False), ('Conv2d', True, True)]
    cart = GridNet(nn.Linear(4, 6), (4,), (9, 11), 5, use_bn=True, f_dim=6)
    st = parse_corrector(cart.corrector)
    assert [m.kernel_size for m, _, _ in st] == [(3, 3), (5, 5), (5, 5), (3, 3)] and [bn is not None for _, bn, _ in st] == [False, True, True, True]
    assert parse_corrector(GridNet(nn.Linear(4, 6), (4,), (9, 11), 5, use_bn=False, f_dim=6).corrector) is not None
    assert parse_corrector(nn.Sequential(nn.Conv2d(3, 3, 3, padding=1, stride=2))) is None                       # strided: module-by-module path
    assert parse_corrector(nn.Sequential(nn.Conv2d(3, 3, 7, padding=3))) is None                                 # 7x7: not a corrector shape
