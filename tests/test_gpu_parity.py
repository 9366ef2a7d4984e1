"""GPU parity tests: the CUDA library (libdyncore.so, sm_100a) through the C ABI / Python
factory API against the oracle and the reference's golden vectors.

Tolerances (metric max|a-b|/max|b| over whole arrays, the reference testsuite's metric,
testsuite.py:48-54): UWIND, VWIND <= 1e-9; POTT, COLP <= 1e-12; QV, QC <= 1e-11
(>= 20x the oracle's own 1-ulp-perturbation floor).  The kernels are built
without FMA contraction and keep the reference's evaluation order, so everything that does
not pass through pow/log is compared BIT-EXACTLY at kernel level.
"""
import numpy as np
import pytest

from helpers import (STATE, TOL, fields_from_golden, golden_dims, grid_from_golden, interior,
                     load_golden, oracle_from_golden, rel_err, state_err)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def cuda_library():
    import torch
    from climate_model_b200 import _lib
    assert torch.cuda.is_available()
    _lib.use_library(_lib.DEFAULT_LIBRARY)      # fails loudly if the .so is missing
    assert _lib.is_cuda(), 'the GPU tests must run the CUDA build, not the host emulation'
    yield


@pytest.fixture(scope='module')
def g10():
    return load_golden('ref_10deg_rand.npz')


def _eq(a, b, what):
    assert np.array_equal(a, b), '%s: max|diff| = %g' % (what, np.nanmax(np.abs(a - b)))


def _diag(GR, F):
    from climate_model_b200.dyn_matsuno import Diagnostics
    from climate_model_b200.io_read_namelist import B200
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))


def test_device_buffers_are_cuda_and_lon_fastest(g10):
    GR = grid_from_golden(g10)
    F = fields_from_golden(GR, g10)
    t = F.device['POTT']
    assert t.is_cuda and t.stride(2) == 1 and t.shape == (GR.nz, GR.NJ, GR.NI)


def test_primary_diag_close_to_oracle(g10):
    """pow() differs between the device libm and glibc by <= 2 ulp"""
    GR = grid_from_golden(g10)
    F = fields_from_golden(GR, g10)
    O = oracle_from_golden(g10)
    O.primary_diag()
    _diag(GR, F)
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    for n in ['PVTF', 'PVTFVB']:
        assert rel_err(F.host[n], O.F[n]) <= 1e-13, n
    for n in ['PHI', 'PHIVB', 'POTTVB']:
        assert rel_err(F.host[n], O.F[n]) <= 1e-13, n


@pytest.fixture()
def strict_library():
    """the strict CUDA build (IEEE divisions, -fmad=false) for the bit-exactness tests"""
    from helpers import CUDA_LIB_STRICT
    from climate_model_b200 import _lib
    _lib.use_library(CUDA_LIB_STRICT)
    assert _lib.is_cuda()
    yield
    _lib.use_library(_lib.DEFAULT_LIBRARY)


def test_kernels_bit_exact_given_oracle_diagnostics(g10, strict_library):
    """STRICT build.  Feed the oracle's PVTF/PHI/... to the device: every tendency kernel, the
    continuity and the Euler step must then reproduce the oracle bit for bit (no pow/log
    involved, except moisture's log interpolation)"""
    from climate_model_b200.dyn_matsuno import Prognostics
    from climate_model_b200.dyn_tendencies import compute_tendencies
    from climate_model_b200.io_read_namelist import B200
    nx, ny, nz, _ = golden_dims(g10)
    GR = grid_from_golden(g10)
    F = fields_from_golden(GR, g10)
    O = oracle_from_golden(g10)
    O.primary_diag()
    for n in ['PVTF', 'PVTFVB', 'PHI', 'PHIVB', 'POTTVB']:
        F.host[n][...] = O.F[n]
        F.to_device(GR, n)
    for n in STATE:
        O.F[n + '_OLD'][:] = O.F[n]
        F.device[n + '_OLD'].copy_(F.device[n])
    O.compute_tendencies()
    compute_tendencies(GR, F)
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    full = (slice(None),) * 3
    box = lambda i1, j1: (slice(1, i1 + 1), slice(1, j1 + 1), slice(None))
    exact = {
        'UFLX': full, 'VFLX': full, 'WWIND': full, 'COLP_NEW': full,
        'FLXDIV': box(nx, ny), 'dCOLPdt': box(nx, ny),
        'WWIND_UWIND': box(nx + 1, ny), 'WWIND_VWIND': box(nx, ny + 1),
        'BFLX': box(nx, ny), 'RFLX': box(nx, ny), 'CFLX': box(nx + 1, ny + 1),
        'QFLX': box(nx + 1, ny + 1), 'DFLX': box(nx, ny + 1), 'EFLX': box(nx, ny + 1),
        'SFLX': box(nx + 1, ny), 'TFLX': box(nx + 1, ny), 'dUFLXdt': box(nx, ny),
        'dVFLXdt': (slice(1, nx + 1), slice(2, ny + 1), slice(None)),
        'dPOTTdt': box(nx, ny),
    }
    for n, sl in exact.items():
        _eq(F.host[n][sl], O.F[n][sl], n)
    for n in ['dQVdt', 'dQCdt']:       # device log() vs glibc log()
        assert rel_err(F.host[n][box(nx, ny)], O.F[n][box(nx, ny)]) <= 1e-12, n
    # Euler step from identical tendencies: bit exact incl. the fused BC images
    for n in ['dQVdt', 'dQCdt']:
        F.host[n][...] = O.F[n]
        F.to_device(GR, n)
    O.F['COLP'][:] = O.F['COLP_NEW']
    F.device['COLP'].copy_(F.device['COLP_NEW'])
    O.euler_forward()
    Prognostics.euler_forward(GR, GR.GRF[B200], **F.get(Prognostics.fields_prognostic, target=B200))
    F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
    for n in STATE:
        _eq(F.host[n], O.F[n], n)


@pytest.mark.parametrize('build', ['production', 'strict'])
@pytest.mark.parametrize('fixture,steps', [('ref_10deg_rand.npz', [1, 2, 10]),
                                           ('ref_5deg.npz', [10, 50])])
def test_step_matsuno_against_reference_golden(fixture, steps, build, request):
    """N Matsuno steps against the REAL reference's numba-CPU outputs, for the production
    build (DC_FAST_MATH) and the strict build"""
    from climate_model_b200.dyn_matsuno import step_matsuno
    if build == 'strict':
        request.getfixturevalue('strict_library')
    g = load_golden(fixture)
    GR = grid_from_golden(g)
    F = fields_from_golden(GR, g)
    _diag(GR, F)
    done = 0
    for s in steps:
        step_matsuno(GR, F, s - done)
        done = s
        F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
        ref = {n: g['N%d_%s' % (s, n)] for n in STATE}
        for n in STATE:
            e = state_err(n, F.host, ref)
            assert e <= TOL[n], 'N%d %s: %.3e > %.0e' % (s, n, e, TOL[n])


def test_factory_path_equals_coarse_entry(g10, strict_library):
    """STRICT build: the factory-by-factory step and the fused coarse entry agree bitwise"""
    from climate_model_b200.dyn_matsuno import step_matsuno, step_matsuno_factories
    out = []
    for stepper in (step_matsuno, step_matsuno_factories):
        GR = grid_from_golden(g10)
        F = fields_from_golden(GR, g10)
        _diag(GR, F)
        for _ in range(2):
            stepper(GR, F)
        F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
        out.append({n: F.host[n].copy() for n in STATE})
    for n in STATE:
        _eq(out[0][n], out[1][n], n)


@pytest.mark.parametrize('moist', [1, 0])
def test_fused_mode_equals_kernel_mode_bitwise(moist, strict_library):
    """STRICT build: the fused stage kernel against the one-kernel-per-reference-kernel mode on
    the device, 3 deg x 12 levels with topography (tiles cut by the domain edge in both
    directions).  (In the production build FMA contraction differs between the two code paths;
    there both are checked against the reference with the parity tolerances.)"""
    from climate_model_b200.dyn_matsuno import set_mode, step_matsuno
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    out = {}
    for mode in ('fused', 'kernels'):
        GR = Grid(nz=12, lat0_deg=-84, lat1_deg=84, dlat_deg=3.0, dlon_deg=3.0,
                  i_moist_main_switch=moist)
        F = ModelFields(GR, UWIND_random_pert=2.0, VWIND_random_pert=2.0, POTT_random_pert=1.0,
                        COLP_random_pert=100.)
        set_mode(GR, mode)
        _diag(GR, F)
        step_matsuno(GR, F, 4)
        F.copy_device_to_host(GR, F.ALL_FIELDS)
        out[mode] = {n: F.host[n].copy() for n in STATE + ['PHI', 'WWIND']}
    for n in STATE[:4] + (STATE[4:] if moist else []) + ['PHI', 'WWIND']:
        _eq(out['fused'][n], out['kernels'][n], n)


@pytest.mark.parametrize('build', ['strict', 'production'])
def test_sigma_column_chunks_are_bit_identical(build, request, monkeypatch):
    """the stage kernel with the sigma column cut into 2 / 4 chunks (what a small latitude band
    runs) against the unchunked march: same arithmetic per cell, bitwise equal in both builds"""
    import torch
    from climate_model_b200.dyn_matsuno import step_matsuno
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    if build == 'strict':
        request.getfixturevalue('strict_library')
    out = {}
    for n in (1, 2, 4):
        monkeypatch.setenv('DC_STAGE_KCHUNKS', str(n))     # read by dc_create
        GR = Grid(nz=32, lat0_deg=-60, lat1_deg=60, dlat_deg=1.0, dlon_deg=1.0)
        F = ModelFields(GR, UWIND_random_pert=2.0, VWIND_random_pert=2.0, POTT_random_pert=1.0,
                        COLP_random_pert=100.)
        _diag(GR, F)
        step_matsuno(GR, F, 3)
        torch.cuda.synchronize()
        out[n] = {m: F.device[m].clone() for m in STATE[:4]}
    for n in (2, 4):
        for m in STATE[:4]:
            assert torch.equal(out[1][m], out[n][m]), (n, m)


def test_production_kernel_mode_within_tolerance(g10):
    """PRODUCTION build, one-kernel-per-reference-kernel mode (what the factories run),
    against the reference's golden outputs"""
    from climate_model_b200.dyn_matsuno import set_mode, step_matsuno
    GR = grid_from_golden(g10)
    F = fields_from_golden(GR, g10)
    set_mode(GR, 'kernels')
    _diag(GR, F)
    step_matsuno(GR, F, 10)
    F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
    ref = {n: g10['N10_' + n] for n in STATE}
    for n in STATE:
        e = state_err(n, F.host, ref)
        assert e <= TOL[n], '%s: %.3e > %.0e' % (n, e, TOL[n])


def test_config2_1deg_32lev_against_oracle():
    """BASELINE.json configs[1]/[2]: 1 deg x 32 levels, elev.1-deg topography, moist tracers on;
    own initial-condition builder feeds both the oracle and the device; 10 steps"""
    from climate_model_b200.dyn_matsuno import step_matsuno
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    from oracle.oracle import GRID_FIELDS, Oracle
    GR = Grid(nz=32, lat0_deg=-84, lat1_deg=84, dlat_deg=1.0, dlon_deg=1.0)
    assert (GR.nx, GR.ny, GR.nz, GR.dt) == (360, 168, 32, 20)
    F = ModelFields(GR)
    O = Oracle(GR.nx, GR.ny, GR.nz, GR.dt, {n: GR.GRF['CPU'][n] for n in GRID_FIELDS})
    O.set(**{n: F.host[n] for n in ['HSURF'] + STATE})
    O.primary_diag()
    _diag(GR, F)
    O.step_matsuno(10)
    step_matsuno(GR, F, 10)
    F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
    for n in STATE:
        e = state_err(n, F.host, O.F)
        assert e <= TOL[n], '%s: %.3e > %.0e' % (n, e, TOL[n])
    assert np.max(np.abs(F.host['UWIND'][interior('UWIND', GR.nx, GR.ny)])) > 5.


@pytest.mark.parametrize('build', ['strict', 'production'])
def test_full_size_properties_quarter_degree_64_levels(build, request):
    """BASELINE.json configs[3] size (1440 x 672 x 64), where the oracle is too slow:
    size-independent properties of the scheme.
      - a state shifted by m cells in longitude must give the shifted result (periodic
        domain): BITWISE in the strict build (identical arithmetic per cell; m is odd, so
        the two columns a thread of the stage kernel owns swap roles), to rounding in the
        production build (FMA contraction may differ between the two columns of a pair and
        between edge and interior tiles);
      - boundary invariants: periodic duplicates, zero meridional wind on the walls;
      - the column-pressure update conserves total mass sum(COLP*A) to rounding."""
    import torch
    from climate_model_b200.dyn_matsuno import step_matsuno
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    if build == 'strict':
        request.getfixturevalue('strict_library')
    GR = Grid(nz=64, lat0_deg=-84, lat1_deg=84, dlat_deg=0.25, dlon_deg=0.25, i_out_nth_hour=1.0)
    assert (GR.nx, GR.ny, GR.nz) == (1440, 672, 64)
    nx, ny = int(GR.nx), int(GR.ny)
    F = ModelFields(GR, i_use_topo=0)
    m = 37
    G = ModelFields(GR, initialize=False)
    js = GR.jshift
    for n in ['HSURF'] + STATE:
        src = F.device[n]
        dst = G.device[n]
        sx = F.fdict[n]['stgx']
        # interior columns 1..nx rolled by m; halos rebuilt by the periodic rule
        dst[:, :, 1:nx + 1] = torch.roll(src[:, :, 1:nx + 1], shifts=m, dims=2)
        dst[:, :, 0] = dst[:, :, nx]
        dst[:, :, nx + 1] = dst[:, :, 1]
        if sx:
            dst[:, :, nx + 2] = dst[:, :, 2]
    a0 = torch.tensor(np.asarray(GR.A[1, :, 0]), device=F.device['COLP'].device)
    mass0 = (F.device['COLP'][0, js + 1:js + ny + 1, 1:nx + 1] * a0[1:ny + 1, None]).sum().item()
    for X in (F, G):
        _diag(GR, X)
        step_matsuno(GR, X, 2)
    torch.cuda.synchronize()
    for n in STATE:
        a = F.device[n][:, :, 1:nx + 1]
        b = G.device[n][:, :, 1:nx + 1]
        assert torch.isfinite(a[:, js + 1:js + ny + 1]).all(), n
        r = torch.roll(a, shifts=m, dims=2)
        if build == 'strict':
            assert torch.equal(r, b), 'shift invariance: ' + n
        else:
            # QC is measured on the water-vapour scale (helpers.state_err)
            scale = G.device['QV'][:, :, 1:nx + 1].abs().max() if n == 'QC' else b.abs().max()
            e = ((r - b).abs().max() / scale.clamp_min(1e-300)).item()
            assert e <= 0.1 * TOL[n], 'shift invariance (rounding): %s %.3e' % (n, e)
    U, V, C = F.device['UWIND'], F.device['VWIND'], F.device['COLP']
    assert torch.equal(U[:, :, nx + 1], U[:, :, 1]) and torch.equal(U[:, :, 0], U[:, :, nx])
    assert (V[:, js + 1] == 0).all() and (V[:, js + ny + 1] == 0).all()
    mass1 = (C[0, js + 1:js + ny + 1, 1:nx + 1] * a0[1:ny + 1, None]).sum().item()
    assert abs(mass1 - mass0) / mass0 < 1e-13


def test_run_diagnostics_on_device(g10):
    """dc_run_diag on the GPU against torch reductions of the same device arrays
    (io_functions.py:70-114: vmax, mass-weighted means, area-weighted COLP, crash check)"""
    import torch
    from climate_model_b200.dyn_matsuno import step_matsuno
    from climate_model_b200.io_functions import diagnose_print_diag_fields, print_ts_info
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    GR = Grid(nz=32, lat0_deg=-84, lat1_deg=84, dlat_deg=1.0, dlon_deg=1.0)
    F = ModelFields(GR)
    _diag(GR, F)
    step_matsuno(GR, F, 2)
    got = diagnose_print_diag_fields(GR, F)
    nx, ny, nz, js = int(GR.nx), int(GR.ny), int(GR.nz), int(GR.jshift)
    U, V, T, C = (F.device[n] for n in ('UWIND', 'VWIND', 'POTT', 'COLP'))
    rows = slice(js + 1, js + ny + 1)
    wx = (U[:, rows, 1:nx + 1] + U[:, rows, 2:nx + 2]) / 2.
    wy = (V[:, rows, 1:nx + 1] + V[:, js + 2:js + ny + 2, 1:nx + 1]) / 2.
    W = torch.sqrt(wx * wx + wy * wy)
    A = torch.tensor(np.asarray(GR.A[1, 1:ny + 1, 0]), device=U.device)[:, None]
    ca = C[0, rows, 1:nx + 1] * A
    want = (W.max().item(), ((W * ca).sum() / ca.sum() / nz).item(),
            ((T[:, rows, 1:nx + 1] * ca).sum() / ca.sum() / nz).item(),
            (ca.sum() / (A.sum() * nx)).item(), U[:, rows, 1:nx + 2].max().item(), 0.)
    assert got[0] == want[0] and got[4] == want[4] and got[5] == 0.
    for g_, w_ in zip(got[1:4], want[1:4]):
        assert abs(g_ - w_) <= 1e-12 * abs(w_), (g_, w_)
    F.device['UWIND'][3, js + 7, 11] = float('nan')
    with pytest.raises(ValueError, match='MODEL CRASH'):
        print_ts_info(GR, F, force=True)


# ---------------------------------------------------------------------------------------
# physics coupling terms with NON-ZERO KMOM / KHEAT / surface fluxes /
# dPOTTdt_RAD against the REAL reference's outputs (tests/golden/ref_10deg_coupled.npz)
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def gc():
    return load_golden('ref_10deg_coupled.npz')


def test_coupled_kernels_bit_exact_given_oracle_diagnostics(gc, strict_library):
    """STRICT build: with the oracle's primary and secondary diagnostics on the device (no
    pow involved), the turbulence / surface-flux / radiation terms and the tendencies that
    contain them reproduce the reference's stage-1 fields bit for bit"""
    from climate_model_b200.dyn_tendencies import compute_tendencies
    nx, ny, nz, _ = golden_dims(gc)
    GR = grid_from_golden(gc)
    F = fields_from_golden(GR, gc)
    O = oracle_from_golden(gc)
    O.primary_diag()
    O.secondary_diag()
    for n in ['PVTF', 'PVTFVB', 'PHI', 'PHIVB', 'POTTVB', 'RHO', 'RHOVB']:
        F.host[n][...] = O.F[n]
        F.to_device(GR, n)
    F.device['COLP_OLD'].copy_(F.device['COLP'])
    compute_tendencies(GR, F)
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    box = lambda i1, j1, j0=1: (slice(1, i1 + 1), slice(j0, j1 + 1), slice(None))
    exact = {
        'KMOM_dUWINDdz': box(nx + 1, ny), 'KMOM_dVWINDdz': box(nx, ny + 1),
        'dUFLXdt_TURB': box(nx, ny), 'dVFLXdt_TURB': box(nx, ny, 2),
        'dPOTTdt_TURB': box(nx, ny), 'dUFLXdt': box(nx, ny), 'dVFLXdt': box(nx, ny, 2),
        'dPOTTdt': box(nx, ny),
    }
    for n, sl in exact.items():
        _eq(F.host[n][sl], gc['S1_' + n][sl], n)
    for n in ['dQVdt_TURB', 'dQVdt', 'dQCdt']:     # device log() vs glibc log() in dQ*dt
        assert rel_err(F.host[n][box(nx, ny)], gc['S1_' + n][box(nx, ny)]) <= 1e-12, n
    assert np.nanmax(np.abs(gc['S1_dUFLXdt_TURB'])) > 1e-2 * np.nanmax(np.abs(gc['S1_dUFLXdt']))


@pytest.mark.parametrize('build', ['production', 'strict'])
def test_coupled_step_matsuno_against_reference_golden(gc, build, request):
    """secondary_diag + step_matsuno as the reference's time loop (solver.py:99-101, :70-73),
    10 steps, both arithmetic modes, parity tolerances"""
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    if build == 'strict':
        request.getfixturevalue('strict_library')
    GR = grid_from_golden(gc)
    F = fields_from_golden(GR, gc)
    _diag(GR, F)
    for ts in range(1, 11):
        Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
        step_matsuno(GR, F)
        if ts in (1, 2, 10):
            F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
            ref = {n: gc['N%d_%s' % (ts, n)] for n in STATE}
            for n in STATE:
                e = state_err(n, F.host, ref)
                assert e <= TOL[n], 'N%d %s: %.3e > %.0e' % (ts, n, e, TOL[n])


# ---------------------------------------------------------------------------------------
# turbulence module (turb_main.py / turb_compute.py; tests/golden/ref_10deg_turb.npz)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize('build', ['production', 'strict'])
def test_turbulence_kernel_known_answers(build, request):
    """the reference kernel's outputs on seeded synthetic inputs that leave KMOM between its
    clamps; the device log() of the humidity mean differs from glibc's by <= 2 ulp"""
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.turb_main import Turbulence
    if build == 'strict':
        request.getfixturevalue('strict_library')
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
        a, b = F.host[n][:, :, 1:-1], gt['KAT_' + n][:, :, 1:-1]
        assert np.max(np.abs(a - b) / np.abs(b)) <= 1e-12, n


@pytest.mark.parametrize('build', ['production', 'strict'])
def test_turbulence_in_the_time_loop_against_reference_golden(build, request):
    """secondary_diag -> turbulence -> step_matsuno for 10 steps against the real reference
    run with its own turbulence module (solver.py:99-112)"""
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.turb_main import Turbulence
    if build == 'strict':
        request.getfixturevalue('strict_library')
    gt = load_golden('ref_10deg_turb.npz')
    GR = grid_from_golden(gt)
    F = fields_from_golden(GR, gt)
    TURB = Turbulence(GR, target=B200)
    _diag(GR, F)
    for ts in range(1, 11):
        Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
        TURB.compute_turbulence(GR, **F.get(TURB.fields_main, target=B200))
        step_matsuno(GR, F)
        if ts in (1, 2, 10):
            F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
            ref = {n: gt['N%d_%s' % (ts, n)] for n in STATE}
            for n in STATE:
                e = state_err(n, F.host, ref)
                assert e <= TOL[n], 'N%d %s: %.3e > %.0e' % (ts, n, e, TOL[n])
    F.to_host(GR, 'KMOM')
    k, kr = F.host['KMOM'][1:-1, 1:-1, 1:-1], gt['N10_KMOM'][1:-1, 1:-1, 1:-1]
    assert np.array_equal(k, kr)          # every cell on the same clamp as the reference


def test_member_stream_equals_one_member_at_a_time(g10):
    """ensemble_stream.MemberStream: upload / step / download of consecutive host-resident
    states overlapped on three CUDA streams, bitwise equal to the plain to_device ->
    primary_diag -> step_matsuno -> to_host sequence; host buffers reused while in flight"""
    from climate_model_b200.dyn_matsuno import step_matsuno
    from climate_model_b200.ensemble_stream import MemberStream, pinned_member
    GR = grid_from_golden(g10)
    F = fields_from_golden(GR, g10)
    rng = np.random.default_rng(5)
    members, want = [], []
    for m in range(3):
        mem = pinned_member(F, STATE)
        for n in STATE:
            mem[n][...] = g10['IN_' + n]
        mem['POTT'][1:-1, 1:-1, :] += rng.uniform(-0.5, 0.5, size=mem['POTT'][1:-1, 1:-1, :].shape)
        mem['POTT'][...] = GR.exchange_BC(mem['POTT'])
        members.append(mem)
    for mem in members:                                   # one member at a time, 2 x 2 steps
        for n in STATE:
            F.host[n][...] = mem[n]
            F.to_device(GR, n)
        for _ in range(2):
            _diag(GR, F)
            step_matsuno(GR, F, 2)
        F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
        want.append({n: F.host[n].copy() for n in STATE})
    ms = MemberStream(GR, F, names=STATE, depth=2)
    # every host set passes twice: the second upload of a set must wait for its first download
    assert ms.advance(members + members, nsteps=2) == 6
    for mem, w in zip(members, want):
        for n in STATE:
            _eq(mem[n], w[n], n)
    assert not np.array_equal(want[0]['POTT'], want[1]['POTT'])


@pytest.mark.parametrize('build', ['strict', 'production'])
@pytest.mark.parametrize('fixture', ['ref_10deg_coupled.npz', 'ref_10deg_turb.npz'])
def test_coupled_increments_beside_the_fused_stage_kernel_on_the_gpu(fixture, build, request,
                                                                    monkeypatch):
    """DC_COUPLED_IMPL=2 on the B200 (round-1 verdict, item 7): fused dry stage kernel +
    TurbPrepBody / TurbApplyBody (the turbulence / surface / radiation increments applied after
    the Euler step instead of inside the tendency sum) against the REAL reference's outputs with
    non-zero coupling fields and with its own turbulence module in the loop; not the reference's
    summation order, hence the parity tolerances, not bit-exactness"""
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import step_matsuno
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.dyn_matsuno import Diagnostics
    from climate_model_b200.turb_main import Turbulence
    monkeypatch.setenv('DC_COUPLED_IMPL', '2')          # read by dc_create
    if build == 'strict':
        request.getfixturevalue('strict_library')
    g = load_golden(fixture)
    turb = 'T1_KMOM' in g
    GR = grid_from_golden(g)
    F = fields_from_golden(GR, g)
    TURB = Turbulence(GR, target=B200)
    _diag(GR, F)
    for ts in range(1, 11):
        Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
        if turb:
            TURB.compute_turbulence(GR, **F.get(TURB.fields_main, target=B200))
        step_matsuno(GR, F)
        if ts in (1, 2, 10):
            F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
            ref = {n: g['N%d_%s' % (ts, n)] for n in STATE}
            for n in STATE:
                e = state_err(n, F.host, ref)
                assert e <= TOL[n], 'N%d %s: %.3e > %.0e' % (ts, n, e, TOL[n])
    GR.close()
