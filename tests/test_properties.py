"""Property tests (hypothesis) of the host-side pieces every path goes through: the layout
transposes of dc_import_field / dc_export_field, the exchange_BC entry against the reference
rule (misc_boundaries.py:22-42), and the latitude-band row split -- on random grid shapes,
through the host emulation of the library."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from helpers import build_emu


@pytest.fixture(scope='module', autouse=True)
def emu_library():
    from climate_model_b200 import _lib
    prev = _lib.library_path()
    _lib.use_library(build_emu())
    yield
    if prev:
        _lib.use_library(prev)


def _grid(nlon, nlat, nz):
    from climate_model_b200.main_grid import Grid
    dlon = 360. / nlon
    return Grid(nz=nz, lat0_deg=-nlat * 2.5, lat1_deg=nlat * 2.5, dlat_deg=5.0, dlon_deg=dlon,
                UVFLX_dif_coef=1.0)


def _reference_exchange_BC(a, nx, ny, stgx, stgy):
    """misc_boundaries.py:22-42 restated with numpy slices"""
    a = a.copy()
    nxs, nys = nx + 1, ny + 1
    if stgx:
        a[0] = a[nxs - 1]
        a[nxs] = a[1]
        a[nxs + 1] = a[2]
    else:
        a[0] = a[nx]
        a[nx + 1] = a[1]
    if stgy:
        for j in (0, 1, nys, nys + 1):
            a[:, j] = 0.
    else:
        a[:, 0] = a[:, 1]
        a[:, ny + 1] = a[:, ny]
    return a


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(nlon=st.sampled_from([8, 9, 12, 15, 16, 30, 36, 45]), nlat=st.integers(3, 14),
       nz=st.integers(3, 11), seed=st.integers(0, 2 ** 16))
def test_layout_roundtrip_and_exchange_bc(nlon, nlat, nz, seed):
    from climate_model_b200 import _lib
    from climate_model_b200.main_fields import ModelFields
    GR = _grid(nlon, nlat, nz)
    assert (int(GR.nx), int(GR.ny)) == (nlon, nlat)
    F = ModelFields(GR, initialize=False)
    rng = np.random.default_rng(seed)
    L = _lib.lib()
    for n, (stgx, stgy) in (('POTT', (0, 0)), ('UWIND', (1, 0)), ('VWIND', (0, 1)),
                            ('COLP', (0, 0)), ('WWIND', (0, 0))):
        a = rng.standard_normal(F.host[n].shape)
        F.host[n][...] = a
        F.to_device(GR, n)
        F.host[n][...] = -7.
        F.to_host(GR, n)
        assert np.array_equal(F.host[n], a), n                        # lossless round trip
        _lib.check(L.dc_exchange_bc(GR.dyncore(), F.table[n][0], 0))
        F.to_host(GR, n)
        want = _reference_exchange_BC(a, nlon, nlat, stgx, stgy)
        assert np.array_equal(F.host[n], want), n
        _lib.check(L.dc_exchange_bc(GR.dyncore(), F.table[n][0], 0))  # idempotent
        F.to_host(GR, n)
        assert np.array_equal(F.host[n], want), n
    GR.close()


@given(ny=st.integers(1, 4000), nranks=st.integers(1, 64))
def test_band_rows_partition_the_latitude_range(ny, nranks):
    from climate_model_b200.main_grid import band_rows
    if nranks > ny:
        nranks = ny
    rows = [band_rows(ny, r, nranks) for r in range(nranks)]
    assert rows[0][0] == 1 and rows[-1][1] == ny
    sizes = [j1 - j0 + 1 for j0, j1 in rows]
    assert all(b[0] == a[1] + 1 for a, b in zip(rows, rows[1:]))      # contiguous, no overlap
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_table_driven_power_and_logarithm_are_within_two_ulp():
    """pow_kappa_tab (Exner function) and log_tab (tracer interface values) of the production
    build (csrc/dc_point.h) against extended-precision references over the arguments the model
    produces: pressures 3 kPa .. 130 kPa, mixing ratios 1e-7 .. 0.1"""
    import ctypes
    import numpy as np
    from helpers import build_emu
    lib = ctypes.CDLL(build_emu(fast=True))
    lib.emu_pow_kappa_tab.restype = ctypes.c_double
    lib.emu_pow_kappa_tab.argtypes = [ctypes.c_double]
    lib.emu_log_tab.restype = ctypes.c_double
    lib.emu_log_tab.argtypes = [ctypes.c_double]
    rng = np.random.default_rng(7)
    kappa = np.longdouble(287.058) / np.longdouble(1005.)
    kappa = np.longdouble(np.float64(287.058 / 1005.))         # the double the kernels use
    worst_pow = worst_log = 0.
    for x in rng.uniform(0.03, 1.3, size=20000):
        ref = np.power(np.longdouble(x), kappa)
        got = lib.emu_pow_kappa_tab(float(x))
        worst_pow = max(worst_pow, float(abs(np.longdouble(got) - ref) / np.spacing(np.float64(ref))))
    for x in np.exp(rng.uniform(np.log(1e-7), np.log(0.1), size=20000)):
        ref = np.log(np.longdouble(x))
        got = lib.emu_log_tab(float(x))
        worst_log = max(worst_log, float(abs(np.longdouble(got) - ref) / np.spacing(np.float64(abs(ref)))))
    assert worst_pow <= 2.0 and worst_log <= 2.0, (worst_pow, worst_log)
    # outside the tabulated binades the power falls back to the library pow
    for x in (1e-3, 7.5):
        assert abs(lib.emu_pow_kappa_tab(x) - x ** (287.058 / 1005.)) <= 4 * np.spacing(x ** 0.2856)
