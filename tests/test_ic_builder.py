"""the grid and initial-condition builders against the arrays the REAL
reference produced (tests/golden: GR_* grid fields and IN_* initial state of two reference runs).
Grid geometry and the sigma levels must be identical; the initial state within 1e-13 (numpy's
vectorised exp / pow vs the reference's scalar math.exp loops)."""
import numpy as np
import pytest

from helpers import build_emu, golden_dims, load_golden
from oracle.oracle import GRID_FIELDS

CASES = {
    'ref_10deg_rand.npz': dict(
        grid=dict(nz=6, lat0_deg=-80, lat1_deg=80, dlat_deg=10, dlon_deg=10, i_out_nth_hour=8),
        ic=dict(UWIND_random_pert=2.0, VWIND_random_pert=2.0, POTT_random_pert=1.0,
                QV_random_pert=0.0005, COLP_random_pert=100.)),
    'ref_5deg.npz': dict(
        grid=dict(nz=8, lat0_deg=-80, lat1_deg=80, dlat_deg=5, dlon_deg=5, i_out_nth_hour=8),
        ic=dict()),
}


@pytest.fixture(scope='module', autouse=True)
def emu_library():
    from climate_model_b200 import _lib
    prev = _lib.library_path()
    _lib.use_library(build_emu())
    yield
    if prev:
        _lib.use_library(prev)


@pytest.mark.parametrize('fixture', sorted(CASES))
def test_grid_matches_reference(fixture):
    from climate_model_b200.main_grid import Grid
    g = load_golden(fixture)
    GR = Grid(**CASES[fixture]['grid'])
    nx, ny, nz, dt = golden_dims(g)
    assert (int(GR.nx), int(GR.ny), int(GR.nz), int(GR.dt)) == (nx, ny, nz, dt)
    for n in GRID_FIELDS:
        a, b = np.asarray(GR.GRF['CPU'][n]), g['GR_' + n]
        assert a.shape == b.shape, n
        assert np.array_equal(a, b, equal_nan=True), '%s: max|diff| %g' % (
            n, np.nanmax(np.abs(a - b)))


@pytest.mark.parametrize('fixture', sorted(CASES))
def test_initial_state_matches_reference(fixture):
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    g = load_golden(fixture)
    GR = Grid(**CASES[fixture]['grid'])
    F = ModelFields(GR, gpu_enable=False, device='cpu', **CASES[fixture]['ic'])
    for n in ['HSURF', 'COLP', 'UWIND', 'VWIND', 'POTT', 'QV', 'QC']:
        a, b = F.host[n], g['IN_' + n]
        assert np.array_equal(np.isnan(a), np.isnan(b)), n + ': NaN pattern'
        m = ~np.isnan(b)
        scale = max(np.max(np.abs(b[m])), 1e-300)
        assert np.max(np.abs(a[m] - b[m])) / scale <= 1e-13, '%s: %g' % (
            n, np.max(np.abs(a[m] - b[m])) / scale)


@pytest.mark.parametrize('kw', [dict(), dict(i_use_topo=0),
                                dict(UWIND_random_pert=2.0, VWIND_random_pert=2.0,
                                     POTT_random_pert=1.0, COLP_random_pert=100.,
                                     POTT_gaussian_pert=3.0, COLP_gaussian_pert=500.)])
def test_band_local_builder_equals_whole_grid_builder(kw):
    """ModelFields(band_local=True) builds, on every rank, only the rows the rank holds
    (io_initial_conditions.initialize_fields_band): bit-identical to the same rows of the
    whole-grid builder, for 1 and for 3 latitude bands (host side only, no stepping)"""
    from helpers import build_emu
    from climate_model_b200 import _lib
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    _lib.use_library(build_emu())
    grid = dict(nz=10, lat0_deg=-78, lat1_deg=78, dlat_deg=3.0, dlon_deg=3.0)
    G0 = Grid(**grid)
    F0 = ModelFields(G0, gpu_enable=False, device='cpu', **kw)
    names = ['HSURF', 'COLP', 'UWIND', 'VWIND', 'POTT', 'QV', 'QC', 'PVTF', 'PVTFVB', 'POTTVB',
             'WWIND']
    for world in (1, 3):
        for rank in range(world):
            G = Grid(band=(rank, world), **grid)
            F = ModelFields(G, gpu_enable=False, device='cpu', band_local=True, **kw)
            for n in names:
                ja, jb = F._rows_of(G)(n)
                assert F.host[n].shape[1] == jb - ja + 1
                assert np.array_equal(F.host[n], F0.host[n][:, ja:jb + 1], equal_nan=True), \
                    (world, rank, n)


def test_band_local_fields_step_like_whole_grid_fields():
    """band-local host arrays through dc_import_rows / dc_export_rows (host emulation, one
    band = the whole grid): same device state and same result after a step"""
    from helpers import STATE, build_emu
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    import torch
    _lib.use_library(build_emu())
    grid = dict(nz=6, lat0_deg=-80, lat1_deg=80, dlat_deg=10, dlon_deg=10)
    out = []
    for band_local in (False, True):
        G = Grid(**grid)
        F = ModelFields(G, band_local=band_local, UWIND_random_pert=1.0)
        Diagnostics.primary_diag(G.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
        step_matsuno(G, F, 2)
        F.copy_device_to_host(G, F.PROGNOSTIC_FIELDS)
        out.append({n: F.host[n].copy() for n in STATE})
    for n in STATE:
        assert np.array_equal(out[0][n], out[1][n], equal_nan=True), n
