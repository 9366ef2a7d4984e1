"""Shared test helpers: golden fixtures, the oracle, comparison metrics."""
import os

import numpy as np

from oracle.oracle import GRID_FIELDS, Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
STATE = ['UWIND', 'VWIND', 'POTT', 'COLP', 'QV', 'QC']

# parity tolerances, metric max|a-b|/max|b| over the interior (testsuite.py:48-54 of the
# reference); SURVEY.md section 8(c): >= 20x the measured 1-ulp-perturbation floor at 5 deg / 1 deg.
#
# 1-ulp-perturbation floor of the reference algorithm ITSELF on the benchmarked shapes (the
# oracle run twice, U, V, POTT perturbed by +-1 ulp; tools/ulp_floor.py, this round):
#   shape                        steps  UWIND    VWIND    POTT     COLP     QV
#   1 deg x 32 (360x168x32)        10   2.3e-11  4.1e-12  1.8e-15  6.4e-16  1.0e-14
#   1 deg x 32                     50   3.0e-11  8.3e-12  4.3e-15  1.7e-15  1.1e-13
#   0.25 deg x 64 band 1440x84     10   1.2e-11  1.7e-11  1.9e-15  9.5e-16  3.2e-15
#   0.25 deg x 64 band 1440x84     50   1.0e-10  8.0e-11  6.5e-15  2.3e-15  1.5e-14
#   0.1 deg x 96 band 3600x32      10   4.3e-11  4.0e-11  4.2e-15  1.1e-15  8.0e-15
# and the CUDA builds against the oracle on the same shapes (tests/test_gpu_bench_shapes.py,
# profiles/r2_parity_bench_shapes.json), 0.25 deg band after 50 steps: production UWIND 1.1e-10,
# VWIND 9.6e-11, POTT 7.7e-15, COLP 4.8e-15; strict 7.6e-11, 8.2e-11, 4.4e-15, 2.2e-15 -- both
# builds sit AT the floor: no implementation of this scheme can be closer to the reference.  At
# 0.25 deg the wind tolerance is therefore 10x the floor (not 20x as on the coarse grids); it is
# kept at 1e-9 because every case passes it with that margin.
TOL = {'UWIND': 1e-9, 'VWIND': 1e-9, 'POTT': 1e-12, 'COLP': 1e-12, 'QV': 1e-11, 'QC': 1e-11}


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_dims(g):
    nx, ny, nz, nb, dt = [int(x) for x in g['dims']]
    assert nb == 1
    return nx, ny, nz, dt


COUPLING = ['KMOM', 'KHEAT', 'SMOMXFLX', 'SMOMYFLX', 'SSHFLX', 'SLHFLX', 'dPOTTdt_RAD']


def oracle_from_golden(g, i_moist=True):
    """the oracle on a golden fixture's grid and inputs; a fixture that carries the physics
    coupling fields (ref_10deg_coupled.npz) switches their terms on"""
    nx, ny, nz, dt = golden_dims(g)
    coupled = 'IN_KMOM' in g
    turb = 'T1_KMOM' in g        # fixture made with the reference's turbulence module
    O = Oracle(nx, ny, nz, dt, {n: g['GR_' + n] for n in GRID_FIELDS}, i_moist=i_moist,
               i_coupling=coupled or turb)
    O.set(**{n: g['IN_' + n] for n in ['HSURF'] + STATE + (COUPLING if coupled else [])})
    if turb:
        for n in COUPLING:
            O.F[n][:] = 0.
    return O


def interior(name, nx, ny):
    """index box of the cells a field owns (halo cells excluded)"""
    from oracle.oracle import FIELDS
    stgx, stgy, _ = FIELDS[name]
    return (slice(1, nx + 1 + stgx), slice(1, ny + 1 + stgy), slice(None))


def rel_err(a, b, scale=None):
    """the reference testsuite's metric: max|a-b| / max|b|"""
    den = np.max(np.abs(b)) if scale is None else scale
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def state_err(n, got, ref):
    """parity metric of prognostic field n; got / ref are {name: array}.
    QC starts at exactly 0 and only picks up the 1e-7 clamp artefact of the reference's
    logarithmic interface interpolation (dyn_functions.py:70-95), i.e. values ~1e-12 that are
    differences of nearly equal fluxes; max|QC| is then not a meaningful scale, so QC errors
    are measured against the water-vapour scale max|QV| (same units, same equations)."""
    if n == 'QC':
        scale = max(np.max(np.abs(ref['QC'])), np.max(np.abs(ref['QV'])))
        return rel_err(got['QC'], ref['QC'], scale)
    return rel_err(got[n], ref[n])


# ---------------------------------------------------------------------------------------
# product-side helpers (climate_model_b200 on a given libdyncore build)
# ---------------------------------------------------------------------------------------
EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu')
CUDA_LIB = os.path.join(os.path.dirname(EMU_DIR), '..', 'climate_model_b200', 'libdyncore.so')
CUDA_LIB_STRICT = os.path.join(os.path.dirname(EMU_DIR), '..', 'climate_model_b200',
                               'libdyncore_strict.so')


def build_emu(fast=False):
    """compile tests/emu/libdyncore_emu[_fast].so (host emulation of the kernel bodies);
    same tiling as the CUDA build; `fast` = the production arithmetic mode (DC_FAST_MATH +
    FMA contraction), default = strict (IEEE divisions, no FMA).  DC_EMU_FLAGS / DC_EMU_TAG in
    the environment build a tagged variant with extra -D switches (development only)."""
    import subprocess
    tag = os.environ.get('DC_EMU_TAG', '')
    extra = os.environ.get('DC_EMU_FLAGS', '').split()
    so = os.path.join(EMU_DIR, 'libdyncore_emu%s%s.so' % ('_fast' if fast else '', tag))
    srcs = [os.path.join(EMU_DIR, 'emu_dyncore.cpp')]
    csrc = os.path.join(os.path.dirname(EMU_DIR), '..', 'climate_model_b200', 'csrc')
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith('.h')]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        mode = (['-DDC_FAST_MATH', '-mfma', '-ffp-contract=fast'] if fast
                else ['-ffp-contract=off'])
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared'] + mode + extra +
                              ['-o', so, srcs[0]])
    return so


def grid_from_golden(g, band=(0, 1), **kw):
    from climate_model_b200.main_grid import Grid
    nx, ny, nz, dt = golden_dims(g)
    arrays = {n: g['GR_' + n] for n in GRID_FIELDS}
    arrays.update(nx=nx, ny=ny, nz=nz, dt=dt)
    if 'IN_KMOM' in g or 'T1_KMOM' in g:    # fixture with non-zero physics coupling fields
        kw.setdefault('i_coupling', 1)
    return Grid(band=band, from_arrays=arrays, **kw)


def fields_from_golden(GR, g, prefix='IN_', names=('HSURF',) + tuple(STATE)):
    """ModelFields whose host state is the golden initial state; coupling fields zero;
    WWIND / POTTVB zero as io_initial_conditions.py:45-46 leaves them"""
    from climate_model_b200.main_fields import ModelFields
    F = ModelFields(GR, gpu_enable=True, initialize=False)
    if GR.i_coupling and prefix + 'KMOM' in g:
        names = tuple(names) + tuple(COUPLING)
    for n in names:
        F.host[n][...] = g[prefix + n]
    F.host['WWIND'][...] = 0.
    F.host['POTTVB'][...] = 0.
    F.copy_host_to_device(GR, F.ALL_FIELDS)
    return F
