#!/usr/bin/env python
"""Run the UNMODIFIED-ALGORITHM reference (Potopoles/Climate_Model, numba CPU path)
in THIS container and dump inputs/outputs of its Matsuno dyn-core step.

TEST INFRASTRUCTURE ONLY.  It needs /root/reference, which does not exist on the
GPU box, so nothing in the product, the `-m gpu` tests, smoke() or bench.py calls
it.  It is the tool that produced tests/golden/*.npz (see tests/golden/make_golden.py)
and that pins oracle/dyncore_oracle.c to the real reference.

No reference source is copied into the repository: the reference's *.py files are
copied to a scratch directory under /tmp at run time, two numba-0.65 typing fixes are
applied there, and three import shims are generated beside them:

  1. netCDF4.Dataset        -> read-only stand-in over scipy.io.netcdf_file
                               (io_initial_conditions.py:15,185-190)
  2. scipy.interpolate.interp2d (removed in SciPy 1.14) -> bilinear
                               RectBivariateSpline(kx=ky=1) with edge clamping
                               (io_initial_conditions.py:16,191-193)
  3. bin.rad_longwave_cython -> stub (rad_longwave.py:22-26 imports it unconditionally)
  4. misc_boundaries.py:37-38  `for j in [0,1,nys,nys+1]` -> four assignments
  5. dyn_functions.py:309-310  assigns KMOM_DWIND but returns KMOM_dDWINDdz
  6. namelist overrides: CPU mode, restart off, physics modules off, grid.

"Dry" configuration: all dyn switches stay 1, physics modules off,
and the coupling fields KMOM, KHEAT, SMOMXFLX, SMOMYFLX, SSHFLX, SLHFLX, dPOTTdt_RAD
are set to exactly 0.0 after ModelFields() built the initial state.
"""
import argparse
import os
import re
import shutil
import sys
import tempfile
import time

REF = '/root/reference'

NETCDF4_SHIM = '''
import numpy as np
from scipy.io import netcdf_file
class _Var:
    def __init__(self, v):
        self._v = v
    def __getitem__(self, idx):
        return np.array(self._v.data[idx], dtype=np.float64)
class Dataset:
    def __init__(self, filename, mode='r', format=None):
        self._f = netcdf_file(filename, 'r', mmap=False)
    def __getitem__(self, name):
        return _Var(self._f.variables[name])
'''

INTERP2D_SHIM = '''
import numpy as np
import scipy.interpolate as _si
class interp2d:
    """bilinear stand-in for the removed scipy.interpolate.interp2d(kind='linear')"""
    def __init__(self, x, y, z, kind='linear'):
        x = np.asarray(x, dtype=np.float64).ravel()
        y = np.asarray(y, dtype=np.float64).ravel()
        z = np.asarray(z, dtype=np.float64)
        ix = np.argsort(x); iy = np.argsort(y)
        self.x = x[ix]; self.y = y[iy]
        z = z[iy, :][:, ix]
        self.spl = _si.RectBivariateSpline(self.x, self.y, z.T, kx=1, ky=1, s=0)
    def __call__(self, xnew, ynew):
        xnew = np.atleast_1d(np.asarray(xnew, dtype=np.float64)).ravel()
        ynew = np.atleast_1d(np.asarray(ynew, dtype=np.float64)).ravel()
        xc = np.clip(xnew, self.x[0], self.x[-1])
        yc = np.clip(ynew, self.y[0], self.y[-1])
        ix = np.argsort(xc); iy = np.argsort(yc)
        out_sorted = self.spl(xc[ix], yc[iy])          # (len(x), len(y))
        out = np.empty_like(out_sorted)
        out[np.ix_(ix, iy)] = out_sorted
        return out.T                                    # (len(y), len(x))
_si.interp2d = interp2d
'''

BIN_STUB = '''
def calc_planck_intensity_c(*a, **k):
    raise NotImplementedError
def calc_surface_emission_c(*a, **k):
    raise NotImplementedError
def rad_calc_LW_RTE_matrix_c(*a, **k):
    raise NotImplementedError
'''


def sub1(text, pattern, repl, count=0, must=True):
    new, n = re.subn(pattern, repl, text, count=count, flags=re.M)
    if must and n == 0:
        raise RuntimeError('pattern not found: ' + pattern)
    return new


def _env(name, default):
    """namelist value that the environment may override at import time (oracle/_ref)"""
    return "type(%r)(__import__('os').environ.get('CMREF_%s', %r))" % (default, name, default)


def prepare_scratch(grid, coupling=False, dest=None, env_grid=False):
    """copy reference *.py + data/ to a scratch dir (or `dest`) and apply shims/patches;
    env_grid: the grid parameters of the patched namelist read CMREF_NZ, CMREF_LAT0_DEG,
    CMREF_LAT1_DEG, CMREF_DLAT_DEG, CMREF_DLON_DEG, CMREF_USE_TOPO from the environment
    (defaults = `grid`), so that one prepared tree serves every sample grid"""
    if dest is None:
        d = tempfile.mkdtemp(prefix='cm_ref_')
    else:
        d = dest
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
    for f in os.listdir(REF):
        if f.endswith('.py'):
            shutil.copy(os.path.join(REF, f), os.path.join(d, f))
    shutil.copytree(os.path.join(REF, 'data'), os.path.join(d, 'data'))
    os.makedirs(os.path.join(d, 'bin'))
    open(os.path.join(d, 'bin', '__init__.py'), 'w').close()
    with open(os.path.join(d, 'bin', 'rad_longwave_cython.py'), 'w') as f:
        f.write(BIN_STUB)
    with open(os.path.join(d, 'netCDF4.py'), 'w') as f:
        f.write(NETCDF4_SHIM)
    with open(os.path.join(d, '_interp2d_shim.py'), 'w') as f:
        f.write(INTERP2D_SHIM)

    # (4) misc_boundaries.py:37-38
    p = os.path.join(d, 'misc_boundaries.py')
    t = open(p).read()
    t = t.replace(
        "        for j in [0,1,nys,nys+1]:\n            FIELD[:,j,:] = wp(0.)\n",
        "        FIELD[:,0,:] = wp(0.)\n        FIELD[:,1,:] = wp(0.)\n"
        "        FIELD[:,nys,:] = wp(0.)\n        FIELD[:,nys+1,:] = wp(0.)\n", 1)
    assert 'FIELD[:,nys+1,:] = wp(0.)' in t
    open(p, 'w').write(t)

    # (5) dyn_functions.py:309-310
    p = os.path.join(d, 'dyn_functions.py')
    t = open(p).read()
    t = sub1(t, r'^        KMOM_DWIND = wp\(0\.\)$', '        KMOM_dDWINDdz = wp(0.)')
    open(p, 'w').write(t)

    # (6) namelist
    p = os.path.join(d, 'namelist.py')
    t = open(p).read()
    t = sub1(t, r'^i_comp_mode = 2$', 'i_comp_mode = 1')
    t = sub1(t, r'^i_load_from_restart = 1$', 'i_load_from_restart = 0')
    t = sub1(t, r'^i_save_to_restart   = 1$', 'i_save_to_restart   = 0')
    t = sub1(t, r'^i_simulation_mode = 2$', 'i_simulation_mode = 2')
    for name in ('i_surface_scheme', 'i_turbulence', 'i_radiation', 'i_microphysics'):
        t = sub1(t, r'^(\s*)%s = 1$' % name, r'\g<1>%s = 0' % name)
    # grid of the "longtime run" preset block
    val = (lambda k, cast: _env(k.upper(), cast(grid[k]))) if env_grid else \
          (lambda k, cast: repr(grid[k]))
    t = sub1(t, r'^    nz = 32$', '    nz = %s' % val('nz', int))
    t = sub1(t, r'^    lat0_deg = -84$', '    lat0_deg = %s' % val('lat0_deg', float))
    t = sub1(t, r'^    lat1_deg = 84$', '    lat1_deg = %s' % val('lat1_deg', float))
    t = sub1(t, r'^    dlat_deg = 1\.0$', '    dlat_deg = %s' % val('dlat_deg', float))
    t = sub1(t, r'^    dlon_deg = 1\.0$', '    dlon_deg = %s' % val('dlon_deg', float))
    t = sub1(t, r'^    i_out_nth_hour = 1/2\*24$',
             '    i_out_nth_hour = %r' % grid.get('i_out_nth_hour', 12.0))
    if env_grid:
        t = sub1(t, r'^i_use_topo = 1$', 'i_use_topo = %s' % _env('USE_TOPO', 1))
    elif not grid.get('use_topo', True):
        t = sub1(t, r'^i_use_topo = 1$', 'i_use_topo = 0')
    for key in ('UWIND_random_pert', 'VWIND_random_pert', 'POTT_random_pert',
                'QV_random_pert', 'COLP_random_pert'):
        if key in grid:
            t = sub1(t, r'^%s\s*=.*$' % key, '%s = %r' % (key, grid[key]))
    open(p, 'w').write(t)

    if coupling:
        # (7) io_read_namelist.py:66-67 switches the radiative-heating term of dyn_POTT.py
        # (:107-108) off when the radiation MODULE is off; the coupling fixture feeds the
        # dynamical core a seeded dPOTTdt_RAD directly, so the term is kept compiled in,
        # exactly as it is when i_radiation = 1
        p = os.path.join(d, 'io_read_namelist.py')
        t = open(p).read()
        t = sub1(t, r'^if i_POTT_radiation and not i_radiation:$', 'if False:')
        open(p, 'w').write(t)
    return d


GRIDS = {
    # config 1 of BASELINE.json: reference's own coarse testsuite grid
    '5deg':  dict(nz=8,  lat0_deg=-80, lat1_deg=80, dlat_deg=5,   dlon_deg=5,
                  i_out_nth_hour=8),
    # same grid, random perturbations on (exercises every cell; seed 3 of the reference)
    '5deg_rand': dict(nz=8, lat0_deg=-80, lat1_deg=80, dlat_deg=5, dlon_deg=5,
                  i_out_nth_hour=8, UWIND_random_pert=2.0, VWIND_random_pert=2.0,
                  POTT_random_pert=1.0, QV_random_pert=0.0005, COLP_random_pert=100.),
    # tiny grid with random perturbations: kernel-level (stage-1) golden vectors
    '10deg_rand': dict(nz=6, lat0_deg=-80, lat1_deg=80, dlat_deg=10, dlon_deg=10,
                  i_out_nth_hour=8, UWIND_random_pert=2.0, VWIND_random_pert=2.0,
                  POTT_random_pert=1.0, QV_random_pert=0.0005, COLP_random_pert=100.),
    '3deg':  dict(nz=12, lat0_deg=-84, lat1_deg=84, dlat_deg=3.0, dlon_deg=3.0),
    '2deg':  dict(nz=16, lat0_deg=-84, lat1_deg=84, dlat_deg=2.0, dlon_deg=2.0),
    # config 2/3 of BASELINE.json
    '1deg':  dict(nz=32, lat0_deg=-84, lat1_deg=84, dlat_deg=1.0, dlon_deg=1.0),
    '1deg_flat': dict(nz=32, lat0_deg=-84, lat1_deg=84, dlat_deg=1.0, dlon_deg=1.0,
                  use_topo=False),
}

COUPLING = ['KMOM', 'KHEAT', 'SMOMXFLX', 'SMOMYFLX', 'SSHFLX', 'SLHFLX', 'dPOTTdt_RAD']
STATE = ['UWIND', 'VWIND', 'POTT', 'COLP', 'QV', 'QC']
GRIDF = ['corf', 'corf_is', 'A', 'sigma_vb', 'dsigma', 'dxjs', 'dyis', 'lat_rad',
         'lat_is_rad', 'dlat_rad', 'dlon_rad', 'POTT_dif_coef', 'UVFLX_dif_coef',
         'moist_dif_coef']
STAGE1 = ['UFLX', 'VFLX', 'FLXDIV', 'dCOLPdt', 'COLP_NEW', 'WWIND', 'WWIND_UWIND',
          'WWIND_VWIND', 'BFLX', 'CFLX', 'DFLX', 'EFLX', 'RFLX', 'QFLX', 'SFLX', 'TFLX',
          'dUFLXdt', 'dVFLXdt', 'dPOTTdt', 'dQVdt', 'dQCdt']
STAGE1_COUPLING = ['KMOM_dUWINDdz', 'KMOM_dVWINDdz', 'dUFLXdt_TURB', 'dVFLXdt_TURB',
                   'dPOTTdt_TURB', 'dQVdt_TURB']
DIAG = ['PHI', 'PHIVB', 'PVTF', 'PVTFVB', 'POTTVB']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--grid', default='5deg')
    ap.add_argument('--steps', type=int, nargs='+', default=[1, 10],
                    help='dump the prognostic state after each of these step counts')
    ap.add_argument('--out', required=True)
    ap.add_argument('--time-steps', type=int, default=0,
                    help='additionally time this many step_matsuno calls')
    ap.add_argument('--perturb-ulp', type=int, default=0,
                    help='perturb U,V,POTT by random +-1ulp (tolerance-floor probe)')
    ap.add_argument('--keep', action='store_true')
    ap.add_argument('--stage1', action='store_true', help='dump stage-1 intermediates')
    ap.add_argument('--minimal', action='store_true',
                    help='dump only grid, inputs and prognostic states (small fixture)')
    ap.add_argument('--coupling', action='store_true',
                    help='fill the physics coupling fields (KMOM, KHEAT, surface fluxes, dPOTTdt_RAD) with '
                         'seeded random values instead of zeros: exercises the turbulence / '
                         'surface-flux terms of the dynamical core')
    ap.add_argument('--turbulence', action='store_true',
                    help='run the reference turbulence module (turb_compute.py: KMOM, KHEAT from '
                         'the bulk Richardson number) after secondary_diag in every time step, as '
                         'solver.py:106-112 does with i_turbulence = 1; dumps KMOM / KHEAT of '
                         'the first call')
    ap.add_argument('--dump-diag', action='store_true',
                    help='also dump the primary diagnostics and WWIND after each step count')
    args = ap.parse_args()

    import numpy as np
    grid = GRIDS[args.grid]
    d = prepare_scratch(grid, coupling=args.coupling)
    os.chdir(d)
    sys.path.insert(0, d)
    import _interp2d_shim  # noqa: F401  (must precede io_initial_conditions import)
    t0 = time.time()
    from io_read_namelist import CPU, gpu_enable
    from main_grid import Grid
    from main_fields import ModelFields
    from dyn_matsuno import step_matsuno
    from dyn_tendencies import compute_tendencies
    from dyn_org_discretizations import DiagnosticsFactory
    GR = Grid()
    F = ModelFields(GR, gpu_enable)
    for n in COUPLING:
        F.host[n][:] = 0.0
    if args.coupling:
        # physically plausible magnitudes; the reference's own turbulence / surface modules
        # stay off, the dynamical core only consumes these fields
        rng = np.random.default_rng(2024)
        for n, (lo, hi) in (('KMOM', (0.01, 0.2)), ('KHEAT', (0.05, 2.)),
                            ('SMOMXFLX', (-0.02, 0.02)), ('SMOMYFLX', (-0.02, 0.02)),
                            ('SSHFLX', (-5., 15.)), ('SLHFLX', (-5., 20.)),
                            ('dPOTTdt_RAD', (-2e-5, 2e-5))):
            F.host[n][:] = rng.uniform(lo, hi, size=F.host[n].shape)
    Diagnostics = DiagnosticsFactory(target=CPU)
    if args.turbulence:
        from turb_compute import compute_turbulence_cpu
        TURB_FIELDS = ['KMOM', 'KHEAT', 'PHIVB', 'HSURF', 'PHI', 'QV', 'WINDX', 'WINDY',
                       'POTTVB', 'POTT']          # turb_main.py:41-42

    if args.turbulence:
        # The module clamps KMOM to [1e-6, 0.01] and the default initial state is so stably
        # stratified that every value sits on the lower clamp.  Compress the vertical
        # potential-temperature gradient (x 0.02 around the lowest level) so that the bulk
        # Richardson number straddles its critical value and both clamps occur.
        P = F.host['POTT']
        P[:] = P[:, :, -1:] + 0.02 * (P - P[:, :, -1:])

    if args.perturb_ulp:
        rng = np.random.default_rng(12345)
        for n in ('UWIND', 'VWIND', 'POTT'):
            a = F.host[n]
            s = rng.integers(-args.perturb_ulp, args.perturb_ulp + 1, size=a.shape)
            F.host[n][:] = a * (1.0 + s * 2.0 ** -52)

    out = {}
    out['dims'] = np.array([GR.nx, GR.ny, GR.nz, GR.nb, GR.dt], dtype=np.int64)
    out['grid_params'] = np.array([grid['lat0_deg'], grid['lat1_deg'], grid['dlat_deg'],
                                   grid['dlon_deg']], dtype=np.float64)
    for n in GRIDF:
        out['GR_' + n] = np.array(GR.GRF[CPU][n])
    out['IN_HSURF'] = F.host['HSURF'].copy()
    if args.coupling:
        for n in COUPLING:
            out['IN_' + n] = F.host[n].copy()
    for n in STATE:
        out['IN_' + n] = F.host[n].copy()

    # solver.py:69-74 : one primary_diag before the loop
    Diagnostics.primary_diag(GR.GRF[CPU], **F.get(Diagnostics.fields_primary_diag, target=CPU))
    if not args.minimal:
        for n in DIAG:
            out['IN_' + n] = F.host[n].copy()
    # solver.py:99-101 : secondary_diag (makes RHO/RHOVB finite so the zeroed
    # turbulence terms evaluate to exactly 0)
    Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=CPU))
    if not args.minimal:
        for n in ('RHO', 'RHOVB', 'TAIR', 'PAIR', 'WIND'):
            out['IN_' + n] = F.host[n].copy()

    if args.turbulence:
        compute_turbulence_cpu(*[F.host[n] for n in TURB_FIELDS])
        out['T1_KMOM'] = F.host['KMOM'].copy()
        out['T1_KHEAT'] = F.host['KHEAT'].copy()
        # known-answer vectors of the kernel itself: seeded synthetic inputs a few
        # centimetres above the surface with metre-scale layers, where the mixing length is
        # small enough for KMOM to fall BETWEEN the clamps (it never does on model states)
        rng = np.random.default_rng(77)
        kat = {n: np.full_like(F.host[n], np.nan) for n in TURB_FIELDS}
        shp, shps = F.host['PHI'].shape, F.host['PHIVB'].shape
        kat['HSURF'] = rng.uniform(0., 2000., size=F.host['HSURF'].shape)
        kat['PHIVB'] = 9.81 * (kat['HSURF'] + rng.uniform(0.01, 0.08, size=shps))
        kat['PHI'] = 9.81 * (kat['HSURF'] + np.cumsum(rng.uniform(0.5, 2., size=shp)[:, :, ::-1],
                                                       axis=2)[:, :, ::-1])
        kat['QV'] = rng.uniform(0., 0.02, size=shp)
        kat['QV'][::3, ::2, 1] = kat['QV'][::3, ::2, 2]          # equal-argument branch of the log mean
        kat['WINDX'] = rng.uniform(-3., 3., size=shp)
        kat['WINDY'] = rng.uniform(-3., 3., size=shp)
        kat['WINDX'][::2, ::3, 3] = kat['WINDX'][::2, ::3, 2]    # min_wind_diff branch
        kat['WINDY'][1::4, :, 1] = kat['WINDY'][1::4, :, 0]
        kat['POTT'] = rng.uniform(280., 320., size=shp)
        kat['POTTVB'] = rng.uniform(280., 320., size=shps)
        compute_turbulence_cpu(*[kat[n] for n in TURB_FIELDS])
        for n in TURB_FIELDS:
            out['KAT_' + n] = kat[n]

    if args.stage1:
        # stage-1 intermediates: compute_tendencies only writes derived fields, so
        # calling it once before the first step leaves the trajectory unchanged.
        F.host['COLP_OLD'][:] = F.host['COLP'][:]
        compute_tendencies(GR, F)
        for n in STAGE1 + (STAGE1_COUPLING if args.coupling else []):
            out['S1_' + n] = F.host[n].copy()
    print('init+jit %.1f s' % (time.time() - t0), flush=True)

    nmax = max(args.steps)
    for ts in range(1, nmax + 1):
        Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=CPU))
        if args.turbulence:
            compute_turbulence_cpu(*[F.host[n] for n in TURB_FIELDS])
        step_matsuno(GR, F)
        if ts in args.steps:
            for n in (STATE + (DIAG + ['WWIND'] if args.dump_diag else []) +
                      (['KMOM', 'KHEAT'] if args.turbulence else [])):
                out['N%d_%s' % (ts, n)] = F.host[n].copy()
    if args.time_steps:
        ts_t = []
        for _ in range(args.time_steps):
            t1 = time.perf_counter()
            step_matsuno(GR, F)
            ts_t.append(time.perf_counter() - t1)
        import numba
        out['timing_s'] = np.array(ts_t)
        out['timing_threads'] = np.array([numba.get_num_threads()])
        print('step_matsuno median %.4f s on %d threads' %
              (float(np.median(ts_t)), numba.get_num_threads()), flush=True)
    np.savez_compressed(args.out, **out)
    print('wrote', args.out, flush=True)
    if not args.keep:
        os.chdir('/')
        shutil.rmtree(d, ignore_errors=True)


if __name__ == '__main__':
    main()
