"""CPU checks of the PRODUCT's host logic and kernel bodies without a GPU.

tests/emu/libdyncore_emu.so is built from the same dc_api_impl.h / dc_kernels.h as the CUDA
library, with "device" memory = host memory and a launch = a loop over the thread box (in
descending order).  Loaded through the same ctypes binding and the same Python field /
factory API, it is compared with the oracle on the reference's golden inputs.  On the
host the libm is glibc's, and the emulation is compiled without FMA contraction, so the
comparison is BIT-EXACT -- any index, boundary-image or ordering mistake shows up here.
(The CUDA build itself is tested by the -m gpu tests.)
"""
import numpy as np
import pytest

from helpers import (STATE, build_emu, fields_from_golden, golden_dims, grid_from_golden,
                     interior, load_golden, oracle_from_golden)


@pytest.fixture(scope='module', autouse=True)
def emu_library():
    from climate_model_b200 import _lib
    prev = _lib.library_path()
    _lib.use_library(build_emu())
    assert not _lib.is_cuda()
    yield
    if prev:
        _lib.use_library(prev)


@pytest.fixture(scope='module')
def g10():
    return load_golden('ref_10deg_rand.npz')


def _eq(a, b, what):
    assert np.array_equal(a, b), '%s: max|diff| = %g' % (what, np.nanmax(np.abs(a - b)))


def test_layout_roundtrip(g10):
    GR = grid_from_golden(g10)
    F = fields_from_golden(GR, g10)
    before = {n: F.host[n].copy() for n in ['UWIND', 'VWIND', 'COLP', 'POTT']}
    for n in before:
        F.host[n][...] = -1.
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    for n, a in before.items():
        assert np.array_equal(a, F.host[n], equal_nan=True), n
    # longitude is the fastest axis of the device layout
    t = F.device['UWIND']
    assert t.shape == (GR.nz, GR.NJ, GR.NI) and t.stride(2) == 1 and GR.NI % 16 == 0


def test_factories_stage1_bit_exact(g10):
    """every factory (= fine-grained C entry) against the oracle after one compute_tendencies,
    Euler step and diagnostics"""
    from climate_model_b200.dyn_matsuno import Diagnostics, Prognostics
    from climate_model_b200.dyn_tendencies import compute_tendencies
    from climate_model_b200.io_read_namelist import B200
    nx, ny, nz, _ = golden_dims(g10)
    GR = grid_from_golden(g10)
    F = fields_from_golden(GR, g10)
    O = oracle_from_golden(g10)
    O.primary_diag()
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    for n in ['PVTF', 'PVTFVB', 'PHI', 'PHIVB', 'POTTVB']:
        _eq(F.host[n], O.F[n], n)

    for n in STATE:                          # dyn_matsuno.py:34-49
        O.F[n + '_OLD'][:] = O.F[n]
        F.device[n + '_OLD'].copy_(F.device[n])
    O.compute_tendencies()
    compute_tendencies(GR, F)
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    full = (slice(None),) * 3
    box = lambda i1, j1: (slice(1, i1 + 1), slice(1, j1 + 1), slice(None))
    ranges = {
        'UFLX': full, 'VFLX': full, 'WWIND': full, 'COLP_NEW': full,
        'FLXDIV': box(nx, ny), 'dCOLPdt': box(nx, ny),
        'WWIND_UWIND': box(nx + 1, ny), 'WWIND_VWIND': box(nx, ny + 1),
        'BFLX': box(nx, ny), 'RFLX': box(nx, ny), 'CFLX': box(nx + 1, ny + 1),
        'QFLX': box(nx + 1, ny + 1), 'DFLX': box(nx, ny + 1), 'EFLX': box(nx, ny + 1),
        'SFLX': box(nx + 1, ny), 'TFLX': box(nx + 1, ny), 'dUFLXdt': box(nx, ny),
        'dVFLXdt': (slice(1, nx + 1), slice(2, ny + 1), slice(None)),
        'dPOTTdt': box(nx, ny), 'dQVdt': box(nx, ny), 'dQCdt': box(nx, ny),
    }
    for n, sl in ranges.items():
        _eq(F.host[n][sl], O.F[n][sl], n)

    O.F['COLP'][:] = O.F['COLP_NEW']
    F.device['COLP'].copy_(F.device['COLP_NEW'])
    O.euler_forward()
    Prognostics.euler_forward(GR, GR.GRF[B200], **F.get(Prognostics.fields_prognostic, target=B200))
    O.secondary_diag()
    Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    for n in STATE:
        _eq(F.host[n], O.F[n], n)            # whole arrays: the fused BC images too
    for n in ['TAIR', 'PAIR', 'RHO', 'TAIRVB', 'PAIRVB', 'RHOVB', 'WINDX', 'WINDY', 'WIND']:
        _eq(F.host[n][interior(n, nx, ny)], O.F[n][interior(n, nx, ny)], n)


@pytest.mark.parametrize('fixture,steps', [('ref_10deg_rand.npz', [1, 2, 10]),
                                           ('ref_5deg.npz', [10])])
def test_step_matsuno_against_reference_golden(fixture, steps):
    """the coarse entry (dc_step_matsuno) against the REAL reference's outputs"""
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    g = load_golden(fixture)
    GR = grid_from_golden(g)
    F = fields_from_golden(GR, g)
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    done = 0
    for s in steps:
        step_matsuno(GR, F, s - done)
        done = s
        F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
        for n in STATE:
            _eq(F.host[n], g['N%d_%s' % (s, n)], 'N%d %s' % (s, n))


@pytest.mark.parametrize('moist', [1, 0])
def test_fused_mode_equals_kernel_mode(g10, moist):
    """dc_step_matsuno: the fused stage kernel (shared-memory tiles, no intermediate fields,
    *_OLD used as second state buffer) against the one-kernel-per-reference-kernel mode"""
    from climate_model_b200.dyn_matsuno import Diagnostics, set_mode, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    out = {}
    for mode in ('fused', 'kernels'):
        GR = grid_from_golden(g10, i_moist_main_switch=moist)
        F = fields_from_golden(GR, g10)
        set_mode(GR, mode)
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        step_matsuno(GR, F, 3)
        # the fused stages keep PHI / POTTVB / PGCOL only: the export itself brings PVTF,
        # PVTFVB and PHIVB up to date (dc_export_field), and evaluates the tendencies of the
        # current state for a flux / tendency field (ModelFields._refresh_for_export)
        F.copy_device_to_host(GR, F.ALL_FIELDS)
        out[mode] = {n: F.host[n].copy() for n in STATE + ['PHI', 'POTTVB', 'PGCOL', 'WWIND',
                                                           'dUFLXdt', 'dQVdt', 'PVTF', 'PVTFVB',
                                                           'PHIVB'] if n in F.host}
        if mode == 'fused':
            # reference for the on-demand tendencies: the kernel decomposition on the same state
            from climate_model_b200.dyn_tendencies import compute_tendencies
            F.device['COLP_OLD'].copy_(F.device['COLP'])
            set_mode(GR, 'kernels')
            compute_tendencies(GR, F)
            F.copy_device_to_host(GR, F.ALL_FIELDS)
            tend = {n: F.host[n].copy() for n in ('dUFLXdt', 'dQVdt')}
    for n in STATE[:4] + (STATE[4:] if moist else []) + ['PHI', 'POTTVB', 'WWIND', 'PVTF',
                                                         'PVTFVB', 'PHIVB']:
        _eq(out['fused'][n], out['kernels'][n], n)
    # the fused path does not keep the tendencies; an export evaluates them at the current state
    assert np.any(out['fused']['dUFLXdt'] != 0.)
    _eq(out['fused']['dUFLXdt'], tend['dUFLXdt'], 'on-demand dUFLXdt')
    if moist:
        assert np.nanmax(np.abs(out['fused']['dQVdt'])) > 0.
        _eq(out['fused']['dQVdt'], tend['dQVdt'], 'on-demand dQVdt')


def test_two_field_sets_on_one_grid(g10):
    """the bindings live on the handle: stepping F1, F2, F1 on ONE grid must advance the set
    that was passed, and a factory call on F2 must not leave F2's PGCOL behind"""
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    GR = grid_from_golden(g10)
    F1 = fields_from_golden(GR, g10)
    F2 = fields_from_golden(GR, g10)
    F2.device['POTT'].add_(0.5)
    ref = {}
    for tag, F, n in (('a', F1, 2), ('b', F2, 1)):      # each set alone on a fresh grid
        G = grid_from_golden(g10)
        X = fields_from_golden(G, g10)
        if tag == 'b':
            X.device['POTT'].add_(0.5)
        Diagnostics.primary_diag(G.GRF[B200], **X.get(Diagnostics.fields_primary_diag, target=B200))
        step_matsuno(G, X, n)
        ref[tag] = {m: X.device[m].clone() for m in STATE[:4]}
    Diagnostics.primary_diag(GR.GRF[B200], **F1.get(Diagnostics.fields_primary_diag, target=B200))
    step_matsuno(GR, F1, 1)
    Diagnostics.primary_diag(GR.GRF[B200], **F2.get(Diagnostics.fields_primary_diag, target=B200))
    step_matsuno(GR, F2, 1)
    step_matsuno(GR, F1, 1)
    import torch
    for m in STATE[:4]:
        assert torch.equal(F1.device[m], ref['a'][m]), m
        assert torch.equal(F2.device[m], ref['b'][m]), m


@pytest.mark.parametrize('kchunks', [2, 3])
def test_sigma_column_chunks_are_bit_identical(g10, kchunks, monkeypatch):
    """the stage kernel with the sigma column cut into chunks (one block each, a warm-up level
    per chunk; used for small latitude bands) against the unchunked march"""
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    out = {}
    for n in (1, kchunks):
        monkeypatch.setenv('DC_STAGE_KCHUNKS', str(n))     # read by dc_create
        GR = grid_from_golden(g10)
        F = fields_from_golden(GR, g10)
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        step_matsuno(GR, F, 3)
        F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
        out[n] = {m: F.host[m].copy() for m in STATE}
    for m in STATE:
        assert np.array_equal(out[1][m], out[kchunks][m], equal_nan=True), m


@pytest.mark.parametrize('fixture,steps', [('ref_10deg_rand.npz', [10]), ('ref_5deg.npz', [10, 50])])
def test_fast_math_mode_within_tolerance(fixture, steps):
    """the PRODUCTION arithmetic mode (reciprocal multiplications, FMA contraction; dc_point.h
    DC_FAST_MATH) against the reference's golden outputs, with the parity tolerances"""
    from helpers import TOL, state_err
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    strict = _lib.library_path()
    _lib.use_library(build_emu(fast=True))
    try:
        g = load_golden(fixture)
        GR = grid_from_golden(g)
        F = fields_from_golden(GR, g)
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        done = 0
        for s in steps:
            step_matsuno(GR, F, s - done)
            done = s
            F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
            ref = {n: g['N%d_%s' % (s, n)] for n in STATE}
            for n in STATE:
                e = state_err(n, F.host, ref)
                assert e <= TOL[n], 'N%d %s: %.3e > %.0e' % (s, n, e, TOL[n])
        GR.close()
    finally:
        _lib.use_library(strict)


def test_factory_path_equals_coarse_entry(g10):
    from climate_model_b200.dyn_matsuno import (Diagnostics, step_matsuno,
                                                 step_matsuno_factories)
    from climate_model_b200.io_read_namelist import B200
    out = []
    for stepper in (step_matsuno, step_matsuno_factories):
        GR = grid_from_golden(g10)
        F = fields_from_golden(GR, g10)
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        for _ in range(2):
            stepper(GR, F)
        F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
        out.append({n: F.host[n].copy() for n in STATE})
    for n in STATE:
        _eq(out[0][n], out[1][n], n)


def test_exchange_bc_entry_matches_reference_rule(g10):
    import ctypes
    from climate_model_b200 import _lib
    GR = grid_from_golden(g10)
    F = fields_from_golden(GR, g10)
    O = oracle_from_golden(g10)
    rng = np.random.default_rng(0)
    L = _lib.lib()
    for n in ['POTT', 'UWIND', 'VWIND', 'COLP']:
        a = rng.standard_normal(F.host[n].shape)
        F.host[n][...] = a
        F.to_device(GR, n)
        _lib.check(L.dc_exchange_bc(GR.dyncore(), F.table[n][0], 0))
        F.to_host(GR, n)
        ref = a.copy()
        _lib_o = __import__('oracle.oracle', fromlist=['lib']).lib()
        _lib_o.orc_exchange_BC.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 3
        _lib_o.orc_exchange_BC(ctypes.byref(O._g), ref.ctypes.data, *ref.shape)
        _eq(F.host[n], ref, n)


def test_errors_are_loud(g10):
    from climate_model_b200 import _lib
    from climate_model_b200.main_grid import Grid
    GR = grid_from_golden(g10)
    L = _lib.lib()
    h = GR.dyncore()
    with pytest.raises(_lib.DyncoreError, match='not bound'):
        _lib.check(L.dc_step_matsuno(h, 1, 0))
    with pytest.raises(_lib.DyncoreError, match='bad field id'):
        _lib.check(L.dc_bind_field(h, 999, 0, 0))
    F = fields_from_golden(GR, g10)
    with pytest.raises(_lib.DyncoreError, match='needs'):
        _lib.check(L.dc_bind_field(h, F.table['UWIND'][0], F.device['UWIND'].data_ptr(), 8))
    # a longitude-dependent grid field is rejected
    arrays = {n: g10['GR_' + n].copy() for n in
              ['corf', 'corf_is', 'A', 'sigma_vb', 'dsigma', 'dxjs', 'dyis', 'lat_rad',
               'lat_is_rad', 'dlat_rad', 'dlon_rad', 'POTT_dif_coef', 'UVFLX_dif_coef',
               'moist_dif_coef']}
    nx, ny, nz, dt = golden_dims(g10)
    arrays.update(nx=nx, ny=ny, nz=nz, dt=dt)
    arrays['A'][3, 4, 0] *= 1.5
    bad = Grid(from_arrays=arrays)
    with pytest.raises(_lib.DyncoreError, match='varies with longitude'):
        bad.dyncore()


def test_solver_entry_point_runs_and_stays_finite(capsys):
    """solver.run: Grid + ModelFields from namelist-style overrides, primary/secondary diag,
    step loop, print-diagnostics with the crash check"""
    from climate_model_b200 import solver
    GR, F = solver.run(nsteps=3, nz=6, lat0_deg=-80, lat1_deg=80, dlat_deg=10, dlon_deg=10,
                       i_out_nth_hour=8)
    assert GR.ts == 3 and GR.sim_time_sec == 3 * GR.dt
    F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
    assert np.isfinite(F.host['UWIND'][1:-2, 1:-1]).all()
    assert 'vmax' in capsys.readouterr().out


@pytest.mark.parametrize('fixture', ['ref_10deg_rand.npz', 'ref_5deg.npz'])
def test_boundary_interior_split_equals_whole_stage(fixture):
    """the band entries with the stage kernel split into boundary and interior tile rows (what
    the NCCL run overlaps with the halo exchange) against dc_step_matsuno"""
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import Diagnostics, _bind_all, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    g = load_golden(fixture)
    out = []
    for split in (False, True):
        GR = grid_from_golden(g)
        F = fields_from_golden(GR, g)
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        if not split:
            step_matsuno(GR, F, 2)
        else:
            L, h = _lib.lib(), GR.dyncore()
            _bind_all(GR, F)
            for _ in range(2):
                _lib.check(L.dc_step_begin(h, 0))
                for stage in (0, 1):
                    _lib.check(L.dc_stage_compute(h, stage, _lib.DC_PART_CONT, 0))
                    # INTERIOR before BOUNDARY: the two are independent
                    _lib.check(L.dc_stage_compute(h, stage, _lib.DC_PART_INTERIOR, 0))
                    _lib.check(L.dc_stage_compute(h, stage, _lib.DC_PART_BOUNDARY, 0))
                    _lib.check(L.dc_halo_pack(h, stage, None, None, 0))
                    _lib.check(L.dc_stage_compute(h, stage, _lib.DC_PART_COLP, 0))
                    _lib.check(L.dc_halo_unpack(h, stage, None, None, 0))
                    _lib.check(L.dc_stage_diag(h, stage, 0))
        F.copy_device_to_host(GR, F.ALL_FIELDS)
        out.append({n: F.host[n].copy() for n in STATE + ['PHI', 'WWIND']})
    for n in STATE + ['PHI', 'WWIND']:
        _eq(out[0][n], out[1][n], n)


# the last three: sigma columns of BASELINE configs[4] (96 levels), a level count that is not a
# multiple of the continuity kernel's 16 levels per thread, and the maximum (128)
@pytest.mark.parametrize('dlon,dlat,nz', [(8.0, 7.0, 5), (2.4, 5.0, 9), (1.25, 3.0, 7),
                                          (10.0, 11.0, 96), (10.0, 11.0, 100), (12.0, 14.0, 128)])
def test_ragged_grids_fused_equals_kernel_mode(dlon, dlat, nz):
    """odd nx (the second column of the last thread pair is masked), tile rows and columns that
    do not divide the grid, odd ny: fused path against the one-kernel-per-reference-kernel mode"""
    from climate_model_b200.dyn_matsuno import Diagnostics, set_mode, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    out = {}
    for mode in ('fused', 'kernels'):
        GR = Grid(nz=nz, lat0_deg=-77, lat1_deg=77, dlat_deg=dlat, dlon_deg=dlon,
                  i_moist_main_switch=1)
        F = ModelFields(GR, UWIND_random_pert=2.0, VWIND_random_pert=2.0, POTT_random_pert=1.0,
                        COLP_random_pert=100.)
        set_mode(GR, mode)
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        step_matsuno(GR, F, 3)
        F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
        out[mode] = {n: F.host[n][interior(n, int(GR.nx), int(GR.ny))].copy() for n in STATE}
    for n in STATE:
        _eq(out['fused'][n], out['kernels'][n], n)


def test_run_diagnostics_on_device_match_reference_formulas(g10):
    """dc_run_diag (two deterministic reduction passes on the device) against the reference's
    diagnose_print_diag_fields (io_functions.py:70-93) restated with numpy on the host state"""
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_functions import diagnose_print_diag_fields, print_ts_info
    from climate_model_b200.io_read_namelist import B200
    nx, ny, nz, _ = golden_dims(g10)
    GR = grid_from_golden(g10)
    F = fields_from_golden(GR, g10)
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    step_matsuno(GR, F, 2)
    got = diagnose_print_diag_fields(GR, F)
    F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
    U, V, T, C = (F.host[n] for n in ('UWIND', 'VWIND', 'POTT', 'COLP'))
    ii, jj = slice(1, nx + 1), slice(1, ny + 1)
    wx = (U[1:nx + 1, jj] + U[2:nx + 2, jj]) / 2.
    wy = (V[ii, 1:ny + 1] + V[ii, 2:ny + 2]) / 2.
    WIND = np.sqrt(wx * wx + wy * wy)
    A = np.asarray(GR.A)[ii, jj, 0]
    ca = C[ii, jj, 0] * A
    mean_wind = sum(np.sum(WIND[:, :, k] * ca) / np.sum(ca) for k in range(nz)) / nz
    mean_temp = sum(np.sum(T[ii, jj, k] * ca) / np.sum(ca) for k in range(nz)) / nz
    want = (np.max(WIND), mean_wind, mean_temp, np.sum(ca) / np.sum(A),
            np.max(U[1:nx + 2, jj]), 0.)
    assert got[0] == want[0] and got[4] == want[4] and got[5] == 0.
    for g_, w_ in zip(got[1:4], want[1:4]):
        assert abs(g_ - w_) <= 1e-13 * abs(w_), (g_, w_)
    # the crash check fires on a NaN in UWIND
    F.device['UWIND'][1, GR.jshift + 3, 5] = float('nan')
    with pytest.raises(ValueError, match='MODEL CRASH'):
        print_ts_info(GR, F, force=True)


# ---------------------------------------------------------------------------------------
# physics coupling terms: NON-ZERO KMOM / KHEAT / surface fluxes / dPOTTdt_RAD,
# compared with the REAL reference's outputs (tests/golden/ref_10deg_coupled.npz)
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def gc():
    return load_golden('ref_10deg_coupled.npz')


def _coupled_setup(gc):
    from climate_model_b200.dyn_matsuno import Diagnostics
    from climate_model_b200.io_read_namelist import B200
    GR = grid_from_golden(gc)
    assert GR.i_coupling == 1
    F = fields_from_golden(GR, gc)
    assert 'KMOM' in F.device and 'dPOTTdt_RAD' in F.device
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    return GR, F


def test_coupled_factories_stage1_bit_exact(gc):
    """compute_tendencies through the factories (= dc_continuity / dc_momentum /
    dc_temperature / dc_moisture) with the coupling fields bound"""
    from climate_model_b200.dyn_matsuno import Diagnostics
    from climate_model_b200.dyn_tendencies import compute_tendencies
    from climate_model_b200.io_read_namelist import B200
    nx, ny, nz, _ = golden_dims(gc)
    GR, F = _coupled_setup(gc)
    Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
    F.device['COLP_OLD'].copy_(F.device['COLP'])
    compute_tendencies(GR, F)
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    box = lambda i1, j1, j0=1: (slice(1, i1 + 1), slice(j0, j1 + 1), slice(None))
    ranges = {
        'KMOM_dUWINDdz': box(nx + 1, ny), 'KMOM_dVWINDdz': box(nx, ny + 1),
        'dUFLXdt_TURB': box(nx, ny), 'dVFLXdt_TURB': box(nx, ny, 2),
        'dPOTTdt_TURB': box(nx, ny), 'dQVdt_TURB': box(nx, ny),
        'dUFLXdt': box(nx, ny), 'dVFLXdt': box(nx, ny, 2), 'dPOTTdt': box(nx, ny),
        'dQVdt': box(nx, ny), 'dQCdt': box(nx, ny),
    }
    for n, sl in ranges.items():
        _eq(F.host[n][sl], gc['S1_' + n][sl], n)
    # the exchange_BC images of the factory (dyn_org_discretizations.py:121-249)
    for n in ['KMOM', 'SMOMXFLX', 'SMOMYFLX']:
        a = F.host[n]
        _eq(a[0], a[nx], n + ' x image')
        _eq(a[:, 0], a[:, 1], n + ' y image')
    GR.close()


@pytest.mark.parametrize('mode', ['fused', 'kernels'])
def test_coupled_step_matsuno_against_reference_golden(gc, mode):
    """secondary_diag + step_matsuno as the reference's time loop (solver.py:99-101, :70-73);
    a handle with i_coupling steps through the kernel decomposition whatever the mode"""
    from climate_model_b200.dyn_matsuno import Diagnostics, set_mode, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    GR, F = _coupled_setup(gc)
    set_mode(GR, mode)
    for ts in range(1, 11):
        Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
        step_matsuno(GR, F)
        if ts in (1, 2, 10):
            F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
            for n in STATE:
                _eq(F.host[n], gc['N%d_%s' % (ts, n)], 'N%d %s' % (ts, n))
    GR.close()


def test_coupled_fast_math_mode_within_tolerance(gc):
    from helpers import TOL, state_err
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    strict = _lib.library_path()
    _lib.use_library(build_emu(fast=True))
    try:
        GR, F = _coupled_setup(gc)
        for ts in range(1, 11):
            Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
            step_matsuno(GR, F)
        F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
        ref = {n: gc['N10_' + n] for n in STATE}
        for n in STATE:
            e = state_err(n, F.host, ref)
            assert e <= TOL[n], 'N10 %s: %.3e > %.0e' % (n, e, TOL[n])
        GR.close()
    finally:
        _lib.use_library(strict)


def test_coupled_zero_fields_equal_the_dry_path(g10):
    """i_coupling with all coupling fields zero reproduces the dry configuration bit for bit
    (what the reference computes when its physics modules are off)"""
    from climate_model_b200.dyn_matsuno import Diagnostics, set_mode, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    out = []
    for cpl in (0, 1):
        GR = grid_from_golden(g10, i_coupling=cpl)
        F = fields_from_golden(GR, g10)
        set_mode(GR, 'kernels')
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        for _ in range(2):
            Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
            step_matsuno(GR, F)
        F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
        out.append({n: F.host[n].copy() for n in STATE})
        GR.close()
    for n in STATE:
        _eq(out[0][n], out[1][n], n)


def test_coupled_errors_are_loud(gc):
    from climate_model_b200 import _lib
    from climate_model_b200.main_grid import Grid
    GR, F = _coupled_setup(gc)
    L = _lib.lib()
    with pytest.raises(_lib.DyncoreError, match='kernel decomposition'):
        _lib.check(L.dc_stage_compute(GR.dyncore(), 0, _lib.DC_PART_ALL, 0))
    # a coupling field that was never bound
    _lib.check(L.dc_bind_field(GR.dyncore(), F.table['KHEAT'][0], 0, 0))
    with pytest.raises(_lib.DyncoreError, match='KHEAT is not bound'):
        _lib.check(L.dc_temperature(GR.dyncore(), 0))
    with pytest.raises(NotImplementedError, match='latitude bands'):
        Grid(band=(0, 2), i_coupling=1)
    GR.close()


# ---------------------------------------------------------------------------------------
# turbulence module (turb_main.py / turb_compute.py) in the time loop, against the real
# reference run with its own turbulence module (tests/golden/ref_10deg_turb.npz)
# ---------------------------------------------------------------------------------------
def test_turbulence_module_against_reference_golden():
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.turb_main import Turbulence
    gt = load_golden('ref_10deg_turb.npz')
    GR = grid_from_golden(gt)
    F = fields_from_golden(GR, gt)
    TURB = Turbulence(GR, target=B200)
    assert TURB.fields_main == ['KMOM', 'KHEAT', 'PHIVB', 'HSURF', 'PHI', 'QV', 'WINDX', 'WINDY',
                                'POTTVB', 'POTT']               # turb_main.py:41-42
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    for ts in range(1, 11):
        Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
        TURB.compute_turbulence(GR, **F.get(TURB.fields_main, target=B200))
        if ts == 1:
            for n in ('KMOM', 'KHEAT'):
                F.to_host(GR, n)
                assert np.array_equal(F.host[n][:, :, 1:-1], gt['T1_' + n][:, :, 1:-1],
                                      equal_nan=True), 'T1 ' + n
        step_matsuno(GR, F)
        if ts in (1, 2, 10):
            F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
            for n in STATE:
                _eq(F.host[n], gt['N%d_%s' % (ts, n)], 'N%d %s' % (ts, n))
    GR.close()


def test_turbulence_kernel_known_answers_bit_exact():
    """the reference kernel's outputs on seeded synthetic inputs that leave KMOM between its
    clamps (tests/golden/ref_10deg_turb.npz KAT_*)"""
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.turb_main import Turbulence
    gt = load_golden('ref_10deg_turb.npz')
    GR = grid_from_golden(gt)
    F = fields_from_golden(GR, gt)
    TURB = Turbulence(GR, target=B200)
    for n in TURB.fields_main[2:]:
        F.host[n][...] = gt['KAT_' + n]
        F.to_device(GR, n)
    TURB.compute_turbulence(GR, **F.get(TURB.fields_main, target=B200))
    for n in ('KMOM', 'KHEAT'):
        F.to_host(GR, n)
        _eq(F.host[n][:, :, 1:-1], gt['KAT_' + n][:, :, 1:-1], 'KAT ' + n)
    GR.close()


def test_solver_with_turbulence_runs_and_stays_finite(capsys):
    from climate_model_b200 import solver
    GR, F = solver.run(nsteps=3, verbose=False, i_turbulence=1, nz=6, lat0_deg=-80, lat1_deg=80,
                       dlat_deg=10, dlon_deg=10)
    assert GR.i_coupling == 1
    F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
    F.to_host(GR, 'KMOM')
    nx, ny = int(GR.nx), int(GR.ny)
    for n in STATE:
        assert np.isfinite(F.host[n][interior(n, nx, ny)]).all(), n
    k = F.host['KMOM'][1:-1, 1:-1, 1:-1]
    assert k.min() >= 1e-6 and k.max() <= 0.01
    GR.close()


def test_member_stream_equals_one_member_at_a_time(g10):
    """ensemble_stream.MemberStream (upload / step / download pipelined over members) against
    the plain to_device -> primary_diag -> step_matsuno -> to_host sequence, bit for bit"""
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.ensemble_stream import MemberStream, pinned_member
    from climate_model_b200.io_read_namelist import B200
    GR = grid_from_golden(g10)
    F = fields_from_golden(GR, g10)
    rng = np.random.default_rng(5)
    names = STATE
    members, want = [], []
    for m in range(5):
        mem = pinned_member(F, names)
        for n in names:
            mem[n][...] = g10['IN_' + n]
        # distinct members: perturb the interior temperature, keep the boundary images
        mem['POTT'][1:-1, 1:-1, :] += rng.uniform(-0.5, 0.5, size=mem['POTT'][1:-1, 1:-1, :].shape)
        mem['POTT'][...] = GR.exchange_BC(mem['POTT'])
        members.append(mem)
    for mem in members:                                   # one member at a time
        for n in names:
            F.host[n][...] = mem[n]
            F.to_device(GR, n)
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        step_matsuno(GR, F, 2)
        F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
        want.append({n: F.host[n].copy() for n in names})
    ms = MemberStream(GR, F, names=names, depth=2)
    assert ms.advance(members, nsteps=2) == 5
    for mem, w in zip(members, want):
        for n in names:
            _eq(mem[n], w[n], n)
    assert not np.array_equal(want[0]['POTT'], want[1]['POTT'])
    GR.close()


# ---------------------------------------------------------------------------------------
# experimental: coupled terms beside the fused dry stage kernel (DC_COUPLED_IMPL=2)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize('fast', [False, True])
@pytest.mark.parametrize('fixture', ['ref_10deg_coupled.npz', 'ref_10deg_turb.npz'])
def test_coupled_increments_beside_the_fused_stage_kernel(fixture, fast, monkeypatch):
    """fused dry stage kernel + TurbPrepBody / TurbApplyBody (X += dt * dX_turb / C after the
    Euler step instead of inside the tendency sum): not the reference's summation order, so the
    comparison with the real reference's outputs uses the parity tolerances, in both arithmetic
    modes; the kernel decomposition stays the default"""
    from helpers import TOL, state_err
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.turb_main import Turbulence
    monkeypatch.setenv('DC_COUPLED_IMPL', '2')
    strict = _lib.library_path()
    _lib.use_library(build_emu(fast=fast))
    try:
        g = load_golden(fixture)
        turb = 'T1_KMOM' in g
        GR = grid_from_golden(g)
        F = fields_from_golden(GR, g)
        TURB = Turbulence(GR, target=B200)
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        launches = _lib.lib().dc_launch_count(GR.dyncore())
        for ts in range(1, 11):
            Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
            if turb:
                TURB.compute_turbulence(GR, **F.get(TURB.fields_main, target=B200))
            step_matsuno(GR, F)
            if ts == 1:   # 3 BC + 2 x (continuity, moisture, prep, stage, apply, diagnostics) + xhalo
                n = _lib.lib().dc_launch_count(GR.dyncore()) - launches - 1 - int(turb)
                assert n == 3 + 2 * 6 + 1, n
            if ts in (1, 2, 10):
                F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
                ref = {n: g['N%d_%s' % (ts, n)] for n in STATE}
                for n in STATE:
                    e = state_err(n, F.host, ref)
                    assert e <= TOL[n], 'N%d %s: %.3e > %.0e' % (ts, n, e, TOL[n])
                    assert e <= 1e-2 * TOL[n], (n, e)    # actual level: 1e-15 .. 1e-12
        GR.close()
    finally:
        _lib.use_library(strict)
