"""pytest wiring.

Markers: ``gpu`` = needs a real B200 and the nvcc-built ``libdbg_b200.so`` (run with ``-m gpu``).
Everything else runs on CPU.  The ``backend`` fixture runs the same parity tests two ways:

* ``gpu``: the product, unchanged (``dbg_b200._native.get()`` -> nvcc-built library).
* ``emu``: the SAME host code and kernel bodies compiled by g++ against ``tests/emu/cuda_emu.h``
  (a serial CUDA emulator) and injected for the duration of the test.  It exists because the
  build container has no GPU; it is test infrastructure, never a product path.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'py-debruijn_b200')
for p in (ROOT, PKG, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200) and the built CUDA library')


@pytest.fixture(scope='session')
def emu_library():
    from emu import build_emu
    from dbg_b200 import _native
    return _native.Library(build_emu.build())


@pytest.fixture(params=[pytest.param('emu'), pytest.param('gpu', marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    from dbg_b200 import _native
    if request.param == 'emu':
        monkeypatch.setattr(_native, '_lib', request.getfixturevalue('emu_library'))
    else:
        monkeypatch.setattr(_native, '_lib', None)
        _native.get()                       # fails loudly if the CUDA library is not built
    return request.param


@pytest.fixture
def backend_gpu(monkeypatch):
    """The product library on a real GPU, nothing else (tests marked ``gpu``)."""
    from dbg_b200 import _native
    monkeypatch.setattr(_native, '_lib', None)
    _native.get()                           # fails loudly if the CUDA library is not built
    return 'gpu'
