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


def test_grid_descriptor_layout_matches_the_header(tmp_path):
    """the ctypes mirror of dc_grid_desc (climate_model_b200/_lib.py) against the C compiler's
    view of include/dyncore.h: size and the offset of every member"""
    import subprocess
    from climate_model_b200._lib import GridDesc
    members = [n for n, _ in GridDesc._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "dyncore.h"', 'int main(void) {',
            '  printf("%zu\\n", sizeof(dc_grid_desc));']
    prog += ['  printf("%%zu\\n", offsetof(dc_grid_desc, %s));' % n for n in members]
    prog += ['  return 0; }']
    src, exe = tmp_path / 'layout.c', tmp_path / 'layout'
    src.write_text('\n'.join(prog))
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), '-o', str(exe), str(src)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(GridDesc)
    for n, off in zip(members, out[1:]):
        assert getattr(GridDesc, n).offset == off, n
