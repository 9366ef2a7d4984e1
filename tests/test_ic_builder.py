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
