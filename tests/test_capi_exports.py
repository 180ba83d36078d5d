"""CPU: libcrown_b200.so loads and exports every symbol include/crown_b200.h declares (no compute
calls without a GPU), and the Python binding's list matches the header."""
import ctypes
import os
import re

from neuralsat_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'crown_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(cb_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_what_the_binding_uses():
    syms = header_symbols()
    assert syms, 'no cb_* declarations found'
    assert sorted(capi.EXPORTS) == syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), 'build with __graft_entry__.build() first'
    L = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(L, s)]
    assert not missing, missing
    L.cb_version.restype = ctypes.c_int
    assert L.cb_version() >= 1
    L.cb_last_error.restype = ctypes.c_char_p
    assert L.cb_last_error() is not None


def test_no_fallback_when_library_is_missing(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, '_lib', None)
    monkeypatch.setattr(capi, 'LIB_PATH', str(tmp_path / 'libcrown_b200.so'))
    try:
        capi.lib()
    except RuntimeError as e:
        assert 'no CPU fallback' in str(e)
    else:
        raise AssertionError('capi.lib() must fail loudly without the CUDA library')


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'neuralsat_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn
