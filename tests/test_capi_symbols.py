"""The C-ABI library loads on a machine without a GPU and exports every
symbol include/aslsd_b200.h declares; no compute call is made."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'aslsd_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(aslsd_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported_and_bound():
    from aslsd_b200 import _build, capi
    _build.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), 'library does not export %s' % name
        assert name in capi.SYMBOLS, 'capi.py does not bind %s' % name
    assert sorted(capi.SYMBOLS) == declared


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path fails loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import aslsd_b200
    with pytest.raises(aslsd_b200.EngineError):
        aslsd_b200.Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'aslsd_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
