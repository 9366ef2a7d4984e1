"""Static checks of the shipped CUDA library (no GPU): resource usage and the instructions
that prove the design -- TMA loads and mbarriers in the tile kernels, no local-memory spills in
the kernels of the fused step, and the asynchronous-copy POTT ring of the diagnostics sweep (a
register prefetch there once shared its scoreboard with a per-level load and never ran ahead:
DESIGN.md section 4)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBS = [os.path.join(ROOT, 'climate_model_b200', n) for n in ('libdyncore.so',
                                                               'libdyncore_strict.so')]
pytestmark = pytest.mark.skipif(shutil.which('cuobjdump') is None, reason='needs cuobjdump')

STEP_KERNELS = ('k_stage3', 'k_moist3', 'k_diagINS_15PrimaryDiagBodyILi1',
                'k_blocksINS_18ContinuityTileBodyILi0')


def _built(lib):
    if not os.path.exists(lib):
        import __graft_entry__ as ge
        ge.build()
    return lib


def _resources(lib):
    out = subprocess.run(['cuobjdump', '--dump-resource-usage', lib], capture_output=True,
                         text=True, check=True).stdout
    res, name = {}, None
    for line in out.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            name = m.group(1)
        m = re.search(r'REG:(\d+) STACK:(\d+) SHARED:(\d+)', line)
        if m and name:
            res[name] = tuple(int(x) for x in m.groups())
    return res


def _sass(lib, pattern):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True,
                         check=True).stdout
    on, text = False, []
    for line in out.splitlines():
        if 'Function :' in line:
            on = pattern in line
        elif on:
            text.append(line)
    return '\n'.join(text)


def test_kernels_of_the_fused_step_do_not_spill():
    # production build (the strict build's IEEE divisions push its continuity kernel over the
    # 128-register budget: 200 B of stack, accepted -- it is the bit-exact build, not the fast one)
    res = _resources(_built(LIBS[0]))
    for k in STEP_KERNELS:
        hit = {n: r for n, r in res.items() if k in n}
        assert hit, k
        for n, (regs, stack, _) in hit.items():
            assert stack == 0, (n, regs, stack)
            assert regs <= 255


@pytest.mark.parametrize('lib', LIBS)
def test_tile_kernels_load_through_tma_and_mbarriers(lib):
    for k in ('k_stage3', 'k_moist3'):
        s = _sass(_built(lib), k)
        assert 'UTMALDG.3D' in s and 'SYNCS.ARRIVE.TRANS64' in s, k
        assert 'sm_100' in subprocess.run(['cuobjdump', '-lelf', lib], capture_output=True,
                                          text=True).stdout


@pytest.mark.parametrize('lib', LIBS)
def test_diagnostics_sweep_requests_pott_through_the_async_copy_ring(lib):
    s = _sass(_built(lib), 'k_diagINS_15PrimaryDiagBodyILi1')
    assert 'LDGSTS.E.64' in s and 'LDGDEPBAR' in s
    # the wait inside the level loop leaves DIAG_PF - 1 = 2 copies in flight
    assert re.search(r'DEPBAR\.LE SB0, 0x2', s)
