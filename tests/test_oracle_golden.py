"""Pins the CPU oracle (oracle/dyncore_oracle.c) to the REAL reference.

The fixtures under tests/golden/ are outputs of the reference's own numba-CPU dynamical
core (tests/golden/make_golden.py -> oracle/run_reference.py).  The oracle follows the
reference's evaluation order without FMA contraction and uses the same libm, so the
comparison is BIT-EXACT (np.array_equal), far tighter than the 1e-9..1e-12 parity
tolerances used for the CUDA path.
"""
import numpy as np
import pytest

from helpers import STATE, golden_dims, interior, load_golden, oracle_from_golden

DIAG = ['PHI', 'PHIVB', 'PVTF', 'PVTFVB', 'POTTVB']


@pytest.fixture(scope='module')
def g10():
    return load_golden('ref_10deg_rand.npz')


def _eq(a, b, what):
    assert a.shape == b.shape, what
    assert np.array_equal(a, b), '%s: max|diff| = %g' % (what, np.nanmax(np.abs(a - b)))


def test_primary_and_secondary_diag_bit_exact(g10):
    O = oracle_from_golden(g10)
    O.primary_diag()
    for n in DIAG:                      # all columns incl. halos (dyn_diagnostics.py:139-195)
        _eq(O.F[n], g10['IN_' + n], n)
    O.secondary_diag()
    for n in ['RHO', 'RHOVB', 'TAIR', 'PAIR']:
        _eq(O.F[n], g10['IN_' + n], n)
    nx, ny, _, _ = golden_dims(g10)
    # WIND reads UWIND[nxs+1], which the python exchange_BC of the initial state leaves
    # NaN (main_grid.py:340-343): compare the interior
    sl = interior('WIND', nx, ny)
    _eq(O.F['WIND'][sl], g10['IN_WIND'][sl], 'WIND')


def test_stage1_intermediates_bit_exact(g10):
    """every array compute_tendencies writes, on the range the reference computes"""
    nx, ny, nz, _ = golden_dims(g10)
    O = oracle_from_golden(g10)
    O.primary_diag()
    O.F['COLP_OLD'][:] = O.F['COLP']
    O.compute_tendencies()
    full = (slice(None),) * 3
    box = lambda i1, j1: (slice(1, i1 + 1), slice(1, j1 + 1), slice(None))
    ranges = {
        'UFLX': full, 'VFLX': full, 'WWIND': full, 'COLP_NEW': full,      # BC-exchanged
        'FLXDIV': box(nx, ny), 'dCOLPdt': box(nx, ny),
        'WWIND_UWIND': box(nx + 1, ny), 'WWIND_VWIND': box(nx, ny + 1),
        'BFLX': box(nx, ny), 'RFLX': box(nx, ny),
        'CFLX': box(nx + 1, ny + 1), 'QFLX': box(nx + 1, ny + 1),
        'DFLX': box(nx, ny + 1), 'EFLX': box(nx, ny + 1),
        'SFLX': box(nx + 1, ny), 'TFLX': box(nx + 1, ny),
        'dUFLXdt': box(nx, ny),                                            # col nxs: garbage
        'dVFLXdt': (slice(1, nx + 1), slice(2, ny + 1), slice(None)),      # wall rows: NaN
        'dPOTTdt': box(nx, ny), 'dQVdt': box(nx, ny), 'dQCdt': box(nx, ny),
    }
    for n, sl in ranges.items():
        ref = g10['S1_' + n][sl]
        assert np.all(np.isfinite(ref)), n
        _eq(O.F[n][sl], ref, n)


@pytest.mark.parametrize('fixture,steps,with_diag', [
    ('ref_10deg_rand.npz', [1, 2, 10], True),
    ('ref_5deg.npz', [10, 50], False),
])
def test_matsuno_steps_bit_exact(fixture, steps, with_diag):
    g = load_golden(fixture)
    O = oracle_from_golden(g)
    O.primary_diag()
    done = 0
    for s in steps:
        O.step_matsuno(s - done)
        done = s
        for n in STATE + (DIAG + ['WWIND'] if with_diag else []):
            _eq(O.F[n], g['N%d_%s' % (s, n)], 'N%d %s' % (s, n))   # whole arrays, halos too


def test_dry_switch_leaves_dry_prognostics_unchanged(g10):
    """moisture is passive: i_moist=0 must give the same U, V, POTT, COLP"""
    a = oracle_from_golden(g10, i_moist=True)
    b = oracle_from_golden(g10, i_moist=False)
    for O in (a, b):
        O.primary_diag()
        O.step_matsuno(3)
    for n in ['UWIND', 'VWIND', 'POTT', 'COLP']:
        _eq(a.F[n], b.F[n], n)


def test_thread_count_does_not_change_results(g10):
    a = oracle_from_golden(g10)
    b = oracle_from_golden(g10)
    nthr = a.num_threads()
    a.primary_diag()
    a.step_matsuno(2)
    b.set_num_threads(1)
    b.primary_diag()
    b.step_matsuno(2)
    b.set_num_threads(nthr)
    for n in STATE:
        _eq(a.F[n], b.F[n], n)


# ---------------------------------------------------------------------------------------
# physics coupling terms: vertical turbulent transport of momentum, heat and
# moisture and the surface fluxes, with NON-ZERO KMOM / KHEAT / SMOM*FLX / SSHFLX / SLHFLX
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def gc():
    return load_golden('ref_10deg_coupled.npz')


def test_coupled_stage1_intermediates_bit_exact(gc):
    nx, ny, nz, _ = golden_dims(gc)
    O = oracle_from_golden(gc)
    O.primary_diag()
    O.secondary_diag()
    O.F['COLP_OLD'][:] = O.F['COLP']
    O.compute_tendencies()
    box = lambda i1, j1, j0=1: (slice(1, i1 + 1), slice(j0, j1 + 1), slice(None))
    ranges = {
        'KMOM_dUWINDdz': box(nx + 1, ny), 'KMOM_dVWINDdz': box(nx, ny + 1),
        'dUFLXdt_TURB': box(nx, ny), 'dVFLXdt_TURB': box(nx, ny, 2),
        'dPOTTdt_TURB': box(nx, ny), 'dQVdt_TURB': box(nx, ny),
        'dUFLXdt': box(nx, ny), 'dVFLXdt': box(nx, ny, 2), 'dPOTTdt': box(nx, ny),
        'dQVdt': box(nx, ny), 'dQCdt': box(nx, ny),
    }
    for n, sl in ranges.items():
        _eq(O.F[n][sl], gc['S1_' + n][sl], n)
    # the coupling terms are not a rounding-level effect in this fixture
    assert np.nanmax(np.abs(gc['S1_dUFLXdt_TURB'])) > 1e-2 * np.nanmax(np.abs(gc['S1_dUFLXdt']))


def test_coupled_step_matsuno_bit_exact(gc):
    """secondary_diag + step_matsuno as the reference's time loop (solver.py:99-101, :70-73)"""
    O = oracle_from_golden(gc)
    O.primary_diag()
    for ts in range(1, 11):
        O.secondary_diag()
        O.step_matsuno(1)
        if ts in (1, 2, 10):
            for n in STATE:
                assert np.array_equal(O.F[n], gc['N%d_%s' % (ts, n)], equal_nan=True), (ts, n)


# ---------------------------------------------------------------------------------------
# turbulence module (turb_compute.py): KMOM / KHEAT computed by the reference itself
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def gt():
    return load_golden('ref_10deg_turb.npz')


TURB_IN = ['PHIVB', 'HSURF', 'PHI', 'QV', 'WINDX', 'WINDY', 'POTTVB', 'POTT']


def test_turbulence_kernel_known_answers_bit_exact(gt):
    """the reference kernel run on seeded synthetic inputs where KMOM falls BETWEEN its clamps
    (on model states it never does: mixing length^2 x shear is ~1e4 times the upper clamp)"""
    O = oracle_from_golden(gt)
    O.set(**{n: gt['KAT_' + n] for n in TURB_IN})
    O.compute_turbulence()
    for n in ('KMOM', 'KHEAT'):
        assert np.array_equal(O.F[n][:, :, 1:-1], gt['KAT_' + n][:, :, 1:-1]), n
    k = gt['KAT_KMOM'][:, :, 1:-1]
    assert ((k > 1e-6) & (k < 0.01)).mean() > 0.9


def test_turbulence_coefficients_bit_exact(gt):
    O = oracle_from_golden(gt)
    O.primary_diag()
    O.secondary_diag()
    O.compute_turbulence()
    for n in ('KMOM', 'KHEAT'):
        assert np.array_equal(O.F[n][:, :, 1:-1], gt['T1_' + n][:, :, 1:-1], equal_nan=True), n
    k = gt['T1_KMOM'][1:-1, 1:-1, 1:-1]
    assert (k == 1e-6).any() and (k == 0.01).any()            # both regimes occur


def test_turbulence_in_the_time_loop_bit_exact(gt):
    """secondary_diag -> turbulence -> step_matsuno (solver.py:99-112, :70-73)"""
    O = oracle_from_golden(gt)
    O.primary_diag()
    for ts in range(1, 11):
        O.secondary_diag()
        O.compute_turbulence()
        O.step_matsuno(1)
        if ts in (1, 2, 10):
            for n in STATE:
                assert np.array_equal(O.F[n], gt['N%d_%s' % (ts, n)], equal_nan=True), (ts, n)
            for n in ('KMOM', 'KHEAT'):
                assert np.array_equal(O.F[n][:, :, 1:-1], gt['N%d_%s' % (ts, n)][:, :, 1:-1],
                                      equal_nan=True), (ts, n)
