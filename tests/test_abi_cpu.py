"""CPU-side checks of the boundary: the shared object builds for sm_100a, loads, exports
every symbol include/mimo_b200.h declares, and refuses to compute without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.join(os.path.dirname(__file__), '..')


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from mimo_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, 'include', 'mimo_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', ' ', header, flags=re.S)
    declared = set(re.findall(r'\b(mimo_\w+)\s*\(', header))
    assert len(declared) >= 25
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(so, name), name
    assert declared == set(lib.SIGNATURES)


def test_version_and_error_string(lib):
    so = lib.load()
    assert so.mimo_version() >= 100
    assert isinstance(lib.last_error(), str)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(lib.MimoCudaError):
        lib.require_device()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'mimo_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f
