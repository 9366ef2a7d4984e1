"""Grid (reference: main_grid.py:77-364): geometry of the regular lat-lon sigma grid, the
GRF dict of grid fields handed to the kernels, the time step from the CFL rule, and the
python exchange_BC used during set-up.

`Grid(**overrides)` takes the namelist's names (nz, lat0_deg, lat1_deg, dlat_deg, dlon_deg,
i_out_nth_hour, ...) so that several grids can live in one process; the reference freezes
one grid per process at import time.  A latitude band of a multi-GPU run is described by
`band=(rank, nranks)`.
"""
import ctypes

import numpy as np

from . import _lib
from . import namelist as nl
from .io_constants import con_omega, con_rE
from .io_initial_conditions import set_up_sigma_levels
from .io_read_namelist import B200, CPU, wp, wp_int
from .misc_utilities import Timer

# module-level extents of the namelist's default grid (main_grid.py:45-52)
nz = wp_int(nl.nz)
nzs = wp_int(nz + 1)
nx = wp_int((nl.lon1_deg - nl.lon0_deg) / nl.dlon_deg)
nxs = wp_int(nx + 1)
ny = wp_int((nl.lat1_deg - nl.lat0_deg) / nl.dlat_deg)
nys = wp_int(ny + 1)
nb = wp_int(nl.nb)

GRID_FIELD_NAMES = ['corf', 'corf_is', 'A', 'sigma_vb', 'dsigma', 'dxjs', 'dyis', 'lat_rad',
                    'lat_is_rad', 'dlat_rad', 'dlon_rad', 'POTT_dif_coef', 'UVFLX_dif_coef',
                    'moist_dif_coef']


def band_rows(ny, rank, nranks):
    """global mass rows [j0, j1] (1-based, inclusive) owned by `rank`: contiguous latitude
    bands, remainder rows to the southern ranks"""
    base, rem = divmod(int(ny), int(nranks))
    j0 = 1 + rank * base + min(rank, rem)
    return j0, j0 + base + (1 if rank < rem else 0) - 1


class Grid:

    def __init__(self, band=(0, 1), from_arrays=None, **overrides):
        P = {k: getattr(nl, k) for k in (
            'nz', 'nb', 'lon0_deg', 'lon1_deg', 'dlon_deg', 'lat0_deg', 'lat1_deg', 'dlat_deg',
            'CFL', 'i_out_nth_hour', 'i_sim_n_days', 'i_restart_nth_day', 'pair_top',
            'POTT_dif_coef', 'moist_dif_coef', 'i_moist_main_switch')}
        P['UVFLX_dif_coef'] = None
        # time step override [s]: a latitude window of a finer global grid keeps the global
        # grid's time step (the rule of main_grid.py:258-275 looks at the window's own rows)
        P['dt'] = None
        # physics coupling terms of the dynamical core (turbulent transport with KMOM / KHEAT,
        # surface fluxes): needed as soon as a physics module fills those fields; the
        # reference always evaluates them (on zero fields when its physics is off)
        P['i_coupling'] = int(bool(nl.i_turbulence or nl.i_surface_scheme))
        for k, v in overrides.items():
            if k not in P:
                raise KeyError('unknown grid parameter %r' % k)
            P[k] = v
        if P['UVFLX_dif_coef'] is None:
            P['UVFLX_dif_coef'] = nl.UVFLX_dif_coef_for(P['dlat_deg'])
        if P['nb'] != 1:
            raise NotImplementedError('nb > 1 not implemented.')
        if P['lon0_deg'] != 0 or P['lon1_deg'] != 360:
            raise NotImplementedError('In x direction only periodic boundaries implemented.')
        self.params = P
        self.pair_top = P['pair_top']
        self.i_moist_main_switch = int(P['i_moist_main_switch'])
        self.i_coupling = int(bool(P['i_coupling']))
        if self.i_coupling and int(band[1]) > 1:
            raise NotImplementedError('the physics coupling terms are not available with '
                                      'latitude bands')
        self.band = (int(band[0]), int(band[1]))
        self._dc = None
        self._dc_device = None
        if from_arrays is not None:
            self._from_arrays(from_arrays)
        else:
            self.create_new_grid()
        self.j0, self.j1 = band_rows(self.ny, *self.band)

    # ------------------------------------------------------------------ reference API
    def create_new_grid(self):
        """main_grid.py:112-296"""
        P = self.params
        self.lon0_deg, self.lon1_deg = P['lon0_deg'], P['lon1_deg']
        self.lat0_deg, self.lat1_deg = P['lat0_deg'], P['lat1_deg']
        self.dlon_deg, self.dlat_deg = P['dlon_deg'], P['dlat_deg']
        self.lon0_rad = self.lon0_deg / 180 * np.pi
        self.lon1_rad = self.lon1_deg / 180 * np.pi
        self.lat0_rad = self.lat0_deg / 180 * np.pi
        self.lat1_rad = self.lat1_deg / 180 * np.pi
        self.dlon_rad_1D = self.dlon_deg / 180 * np.pi
        self.dlat_rad_1D = self.dlat_deg / 180 * np.pi

        self.nz = wp_int(P['nz'])
        self.nzs = wp_int(self.nz + 1)
        self.nx = wp_int((self.lon1_deg - self.lon0_deg) / self.dlon_deg)
        self.nxs = wp_int(self.nx + 1)
        self.ny = wp_int((self.lat1_deg - self.lat0_deg) / self.dlat_deg)
        self.nys = wp_int(self.ny + 1)
        self.nb = wp_int(P['nb'])
        nx, nxs, ny, nys, nz, nb = self.nx, self.nxs, self.ny, self.nys, self.nz, self.nb
        self._index_arrays()
        ii, jj, iis, jjs = self.ii, self.jj, self.iis, self.jjs

        def full(fx, fy, v=np.nan):
            return np.full((fx, fy, 1), v, dtype=wp)

        self.lon_deg, self.lat_deg = full(nx + 2, ny + 2), full(nx + 2, ny + 2)
        self.lon_is_deg, self.lat_is_deg = full(nxs + 2, ny + 2), full(nxs + 2, ny + 2)
        self.lon_js_deg, self.lat_js_deg = full(nx + 2, nys + 2), full(nx + 2, nys + 2)
        self.dlon_rad = full(nx + 2, nys + 2, self.dlon_rad_1D)
        self.dlat_rad = full(nxs + 2, ny + 2, self.dlat_rad_1D)
        self.lon_deg[ii, jj, 0] = self.lon0_deg + (ii - nb + 0.5) * self.dlon_deg
        self.lon_is_deg[iis, jj, 0] = self.lon0_deg + (iis - nb) * self.dlon_deg
        self.lon_js_deg[ii, jjs, 0] = self.lon0_deg + (ii - nb + 0.5) * self.dlon_deg
        self.lat_deg[ii, jj, 0] = self.lat0_deg + (jj - nb + 0.5) * self.dlat_deg
        self.lat_js_deg[ii, jjs, 0] = self.lat0_deg + (jjs - nb) * self.dlat_deg
        self.lat_is_deg[iis, jj, 0] = self.lat0_deg + (jj - nb + 0.5) * self.dlat_deg
        for n in ('lon', 'lat', 'lon_is', 'lat_is', 'lon_js', 'lat_js'):
            setattr(self, n + '_rad', getattr(self, n + '_deg') / 180 * np.pi)

        self.dx, self.dxjs, self.dyis = full(nx + 2, ny + 2), full(nx + 2, nys + 2), \
            full(nxs + 2, ny + 2)
        self.dx[ii, jj, 0] = np.cos(self.lat_rad[ii, jj, 0]) * self.dlon_rad_1D * con_rE
        self.dxjs[ii, jjs, 0] = np.cos(self.lat_js_rad[ii, jjs, 0]) * self.dlon_rad_1D * con_rE
        self.dyis[iis, jj, 0] = self.dlat_rad_1D * con_rE
        self.dx = self.exchange_BC(self.dx)
        self.dxjs = self.exchange_BC(self.dxjs)
        self.dyis = self.exchange_BC(self.dyis)
        self.dy = self.dlat_rad * con_rE

        self.A = full(nx + 2, ny + 2)
        self.A[ii, jj, 0] = lat_lon_recangle_area(self.lat_rad[ii, jj, 0], self.dlon_rad_1D,
                                                  self.dlat_rad_1D)
        self.A = self.exchange_BC(self.A)

        self.corf, self.corf_is = full(nx + 2, ny + 2), full(nxs + 2, ny + 2)
        self.corf[ii, jj, 0] = 2 * con_omega * np.sin(self.lat_rad[ii, jj, 0])
        self.corf_is[iis, jj, 0] = 2 * con_omega * np.sin(self.lat_is_rad[iis, jj, 0])

        self.level = np.arange(0, nz)
        self.levels = np.arange(0, self.nzs)
        self.sigma_vb = np.full(self.nzs, np.nan, dtype=wp)
        self.dsigma = np.full(nz, np.nan, dtype=wp)
        set_up_sigma_levels(self, pair_top=self.pair_top)
        self.dsigma = self.dsigma[None, None, :]
        self.sigma_vb = self.sigma_vb[None, None, :]

        # TIME STEP (main_grid.py:258-275)
        mindx = np.nanmin(self.dx)
        self.CFL = P['CFL']
        self.i_out_nth_hour = P['i_out_nth_hour']
        self.nc_output_count = 0
        self.i_sim_n_days = P['i_sim_n_days']
        self.dt = int(self.CFL * mindx / 400)
        while self.i_out_nth_hour * 3600 % self.dt > 0:
            self.dt -= 1
        if P['dt'] is not None:
            self.dt = int(P['dt'])
        self._time_bookkeeping(P)

        # NUMERICAL DIFFUSION (main_grid.py:279-292)
        k = np.arange(nz)
        self.UVFLX_dif_coef = np.zeros((1, 1, nz), dtype=wp)
        self.POTT_dif_coef = np.zeros((1, 1, nz), dtype=wp)
        self.moist_dif_coef = np.zeros((1, 1, nz), dtype=wp)
        vert_reduce = .0
        self.UVFLX_dif_coef[0, 0, k] = wp(P['UVFLX_dif_coef']) * np.exp(
            -vert_reduce * (nz - k - 1) / nz)
        vert_reduce = 1.5
        self.POTT_dif_coef[0, 0, k] = wp(P['POTT_dif_coef']) * np.exp(
            -vert_reduce * (nz - k - 1) / nz)
        self.moist_dif_coef[0, 0, k] = wp(P['moist_dif_coef']) * np.exp(
            -vert_reduce * (nz - k - 1) / nz)
        self.copy_to_gpu()

    def _index_arrays(self):
        nb = self.nb
        self.i = np.arange(nb, self.nx + nb)
        self.i_s = np.arange(nb, self.nxs + nb)
        self.j = np.arange(nb, self.ny + nb)
        self.js = np.arange(nb, self.nys + nb)
        self.k = np.arange(self.nz)
        self.ii, self.jj = np.ix_(self.i, self.j)
        self.iis, self.jjs = np.ix_(self.i_s, self.js)

    def _time_bookkeeping(self, P):
        self.nts = P['i_sim_n_days'] * 3600 * 24 / self.dt
        self.ts = 0
        self.i_out_nth_ts = int(P['i_out_nth_hour'] * 3600 / self.dt)
        self.i_restart_nth_day = P['i_restart_nth_day']
        self.i_restart_nth_ts = int(self.i_restart_nth_day * 24 / P['i_out_nth_hour']
                                    * self.i_out_nth_ts) if P['i_out_nth_hour'] else 0
        self.sim_time_sec = 0
        self.timer = Timer()

    def _from_arrays(self, a):
        """grid taken from existing GRF arrays (reference layout) + dims: used to run on
        exactly the grid a reference run dumped (tests/golden)"""
        P = self.params
        self.nx, self.ny, self.nz = (wp_int(a['nx']), wp_int(a['ny']), wp_int(a['nz']))
        self.nxs, self.nys, self.nzs = (wp_int(self.nx + 1), wp_int(self.ny + 1),
                                        wp_int(self.nz + 1))
        self.nb = wp_int(1)
        self.dt = int(a['dt'])
        self._index_arrays()
        for n in GRID_FIELD_NAMES:
            setattr(self, n, np.ascontiguousarray(a[n], dtype=wp))
        self.i_out_nth_hour = P['i_out_nth_hour']
        self.i_sim_n_days = P['i_sim_n_days']
        self._time_bookkeeping(P)
        self.copy_to_gpu()

    def copy_to_gpu(self):
        """main_grid.py:298-315: GRF[target] dicts.  GRF[CPU] holds the host arrays in the
        reference layout; GRF[B200] the same arrays (the C library turns them into per-row /
        per-level device vectors in dc_create -- grid fields depend on latitude only)."""
        self.GRF = {CPU: {}, B200: {}}
        for n in GRID_FIELD_NAMES:
            self.GRF[CPU][n] = getattr(self, n)
            self.GRF[B200][n] = getattr(self, n)

    def exchange_BC(self, FIELD):
        """main_grid.py:319-358 (python, set-up only).  Note: like the reference it does
        not fill FIELD[nxs+1] of x-staggered fields."""
        dim2 = FIELD.ndim == 2
        fnx, fny = FIELD.shape[0], FIELD.shape[1]
        if fnx == self.nxs + 2 * self.nb:
            FIELD[0, ::] = FIELD[self.nxs - 1, ::]
            FIELD[self.nxs, ::] = FIELD[1, ::]
        else:
            FIELD[0, ::] = FIELD[self.nx, ::]
            FIELD[self.nx + 1, ::] = FIELD[1, ::]
        if fny == self.nys + 2 * self.nb:
            for j in [0, 1, self.nys, self.nys + 1]:
                if dim2:
                    FIELD[:, j] = wp(0.)
                else:
                    FIELD[:, j, :] = wp(0.)
        else:
            if dim2:
                FIELD[:, 0] = FIELD[:, 1]
                FIELD[:, self.ny + 1] = FIELD[:, self.ny]
            else:
                FIELD[:, 0, :] = FIELD[:, 1, :]
                FIELD[:, self.ny + 1, :] = FIELD[:, self.ny, :]
        return FIELD

    # ------------------------------------------------------------------ B200 handle
    def dyncore(self):
        """the libdyncore handle of this grid (created on first use), see include/dyncore.h"""
        if self._dc is None:
            L = _lib.lib()
            d = _lib.GridDesc(nx=int(self.nx), ny=int(self.ny), nz=int(self.nz),
                              j0=int(self.j0), j1=int(self.j1),
                              i_moist=self.i_moist_main_switch, dt=float(self.dt),
                              pair_top=float(self.pair_top), i_coupling=self.i_coupling)
            self._keep = []
            for n in _lib.GRID_FIELDS_2D + _lib.GRID_FIELDS_1D:
                a = np.ascontiguousarray(self.GRF[B200][n], dtype=np.float64)
                self._keep.append(a)
                setattr(d, n, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
            h = ctypes.c_void_p()
            _lib.check(L.dc_create(ctypes.byref(d), ctypes.byref(h)))
            self._dc = h
            NI, NJ, js = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            _lib.check(L.dc_get_layout(h, ctypes.byref(NI), ctypes.byref(NJ), ctypes.byref(js)))
            self.NI, self.NJ, self.jshift = NI.value, NJ.value, js.value
        return self._dc

    def close(self):
        if self._dc is not None:
            _lib.lib().dc_destroy(self._dc)
            self._dc = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lat_lon_recangle_area(lat, dlon, dlat):
    """main_grid.py:361-363"""
    return np.cos(lat) * dlon * dlat * con_rE ** 2
