"""The C-ABI boundary: every function include/dyncore.h declares is exported by the CUDA
libraries (production and strict build) and by the host-emulation build of the same sources.
Loading needs no GPU (no compute call is made here)."""
import ctypes
import os
import re

import pytest

from helpers import CUDA_LIB, CUDA_LIB_STRICT, build_emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'dyncore.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = set(re.findall(r'\b(dc_[a-z0-9_]+)\s*\(', src))
    assert len(names) >= 30, names
    return sorted(names)


@pytest.mark.parametrize('which', ['production', 'strict', 'emulation'])
def test_library_exports_every_declared_symbol(which):
    path = {'production': CUDA_LIB, 'strict': CUDA_LIB_STRICT}.get(which) or build_emu()
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(os.path.abspath(path))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, '%s does not export %s' % (path, missing)
    lib.dc_is_cuda.restype = ctypes.c_int
    assert lib.dc_is_cuda() == (0 if which == 'emulation' else 1)
    lib.dc_num_fields.restype = ctypes.c_int
    lib.dc_field_id.argtypes = [ctypes.c_char_p]
    assert lib.dc_num_fields() >= 49 and lib.dc_field_id(b'PGCOL') >= 0


def test_product_binding_refuses_to_run_without_the_cuda_library(tmp_path, monkeypatch):
    """no CPU fallback: a missing libdyncore.so is an ImportError, not a silent emulation"""
    from climate_model_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'DEFAULT_LIBRARY', str(tmp_path / 'libdyncore.so'))
    with pytest.raises(ImportError, match='no CPU fallback'):
        _lib.lib()
