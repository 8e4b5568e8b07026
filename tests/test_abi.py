"""The C-ABI library loads and exports every symbol include/dbg_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from dbg_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'dbg_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dbg_[a-z_]+)\s*\(', src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_native.SYMBOLS)


def test_cuda_library_exports_every_symbol():
    """The nvcc-built library (cross-compiled for sm_100a; loadable without a GPU)."""
    path = _native.build()
    dll = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(dll, name), name
    dll.dbg_version.restype = ctypes.c_int
    assert dll.dbg_version() == 1


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_native.Contig) == 32
    assert ctypes.sizeof(_native.Segment) == 8
    assert ctypes.sizeof(_native.Stats) == 19 * 8 + 4 * 4 + 18 * 4      # dbg_stats_t: 19 u64, 4 u32, 18 float


def test_product_refuses_to_run_without_a_gpu():
    """No CPU fallback: on a box without CUDA the product raises instead of computing."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import debruijn as db
    _native._lib = None
    with pytest.raises(Exception) as ei:
        db.construct_graph(['ACGTACGT'], 3, 2)
    assert 'CUDA' in str(ei.value) or 'cuda' in str(ei.value)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, '_lib', None)
    monkeypatch.setattr(_native, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _native.get()
