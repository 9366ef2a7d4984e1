"""ctypes front-end of the CPU oracle (oracle/dyncore_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this.

Arrays are numpy float64 in the REFERENCE layout (i, j, k), k fastest
(main_fields.py:477-485).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GRID_FIELDS = ['A', 'dxjs', 'dyis', 'corf', 'corf_is', 'lat_rad', 'lat_is_rad', 'dlon_rad',
               'dlat_rad', 'sigma_vb', 'dsigma', 'UVFLX_dif_coef', 'POTT_dif_coef',
               'moist_dif_coef']

# name -> (stgx, stgy, 'nz' | 'nzs' | 1)   (main_fields.py:233-470)
FIELDS = {
    'COLP': (0, 0, 1), 'COLP_OLD': (0, 0, 1), 'COLP_NEW': (0, 0, 1), 'dCOLPdt': (0, 0, 1),
    'HSURF': (0, 0, 1),
    'UWIND': (1, 0, 'nz'), 'UWIND_OLD': (1, 0, 'nz'), 'VWIND': (0, 1, 'nz'),
    'VWIND_OLD': (0, 1, 'nz'), 'WWIND': (0, 0, 'nzs'),
    'POTT': (0, 0, 'nz'), 'POTT_OLD': (0, 0, 'nz'), 'QV': (0, 0, 'nz'), 'QV_OLD': (0, 0, 'nz'),
    'QC': (0, 0, 'nz'), 'QC_OLD': (0, 0, 'nz'),
    'UFLX': (1, 0, 'nz'), 'VFLX': (0, 1, 'nz'), 'FLXDIV': (0, 0, 'nz'),
    'BFLX': (0, 0, 'nz'), 'CFLX': (1, 1, 'nz'), 'DFLX': (0, 1, 'nz'), 'EFLX': (0, 1, 'nz'),
    'RFLX': (0, 0, 'nz'), 'QFLX': (1, 1, 'nz'), 'SFLX': (1, 0, 'nz'), 'TFLX': (1, 0, 'nz'),
    'WWIND_UWIND': (1, 0, 'nzs'), 'WWIND_VWIND': (0, 1, 'nzs'),
    'dUFLXdt': (1, 0, 'nz'), 'dVFLXdt': (0, 1, 'nz'), 'dPOTTdt': (0, 0, 'nz'),
    'dQVdt': (0, 0, 'nz'), 'dQCdt': (0, 0, 'nz'),
    'PHI': (0, 0, 'nz'), 'PHIVB': (0, 0, 'nzs'), 'PVTF': (0, 0, 'nz'), 'PVTFVB': (0, 0, 'nzs'),
    'POTTVB': (0, 0, 'nzs'),
    'TAIR': (0, 0, 'nz'), 'TAIRVB': (0, 0, 'nzs'), 'PAIR': (0, 0, 'nz'), 'PAIRVB': (0, 0, 'nzs'),
    'RHO': (0, 0, 'nz'), 'RHOVB': (0, 0, 'nzs'), 'WINDX': (0, 0, 'nz'), 'WINDY': (0, 0, 'nz'),
    'WIND': (0, 0, 'nz'),
    # physics coupling (order = orc_fields)
    'KMOM': (0, 0, 'nzs'), 'KHEAT': (0, 0, 'nzs'), 'SMOMXFLX': (0, 0, 1), 'SMOMYFLX': (0, 0, 1),
    'SSHFLX': (0, 0, 1), 'SLHFLX': (0, 0, 1),
    'KMOM_dUWINDdz': (1, 0, 'nzs'), 'KMOM_dVWINDdz': (0, 1, 'nzs'),
    'dUFLXdt_TURB': (1, 0, 'nz'), 'dVFLXdt_TURB': (0, 1, 'nz'), 'dPOTTdt_TURB': (0, 0, 'nz'),
    'dQVdt_TURB': (0, 0, 'nz'), 'dPOTTdt_RAD': (0, 0, 'nz'),
}

_dp = ctypes.POINTER(ctypes.c_double)


class _Grid(ctypes.Structure):
    _fields_ = ([('nx', ctypes.c_int), ('ny', ctypes.c_int), ('nz', ctypes.c_int),
                 ('i_moist', ctypes.c_int), ('dt', ctypes.c_double),
                 ('pair_top', ctypes.c_double)] + [(n, _dp) for n in GRID_FIELDS] +
                [('i_coupling', ctypes.c_int)])


class _Fields(ctypes.Structure):
    _fields_ = [(n, _dp) for n in FIELDS]


def build(force=False):
    """compile oracle/libdyncore_oracle.so with the committed Makefile"""
    so = os.path.join(_HERE, 'libdyncore_oracle.so')
    src = os.path.join(_HERE, 'dyncore_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        for fn in ('orc_continuity', 'orc_momentum', 'orc_temperature', 'orc_moisture',
                   'orc_compute_tendencies', 'orc_euler_forward', 'orc_primary_diag',
                   'orc_secondary_diag', 'orc_step_matsuno', 'orc_compute_turbulence'):
            getattr(_LIB, fn).argtypes = [ctypes.POINTER(_Grid), ctypes.POINTER(_Fields)]
            getattr(_LIB, fn).restype = None
        _LIB.orc_num_threads.restype = ctypes.c_int
        _LIB.orc_set_num_threads.argtypes = [ctypes.c_int]
    return _LIB


def field_shape(name, nx, ny, nz):
    stgx, stgy, dz = FIELDS[name]
    nk = {'nz': nz, 'nzs': nz + 1, 1: 1}[dz]
    return (nx + 2 + stgx, ny + 2 + stgy, nk)


class Oracle:
    """Holds a full set of model fields (reference layout) and runs the C oracle on them."""

    def __init__(self, nx, ny, nz, dt, grid, i_moist=True, pair_top=10000., i_coupling=False):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.dt = float(dt)
        self.grid = {n: np.ascontiguousarray(grid[n], dtype=np.float64) for n in GRID_FIELDS}
        self.F = {}
        for n in FIELDS:
            fill = 0.0 if n in ('WWIND', 'POTTVB') else np.nan   # io_initial_conditions.py:45-46
            self.F[n] = np.full(field_shape(n, nx, ny, nz), fill, dtype=np.float64)
        self._g = _Grid(nx=self.nx, ny=self.ny, nz=self.nz, i_moist=int(bool(i_moist)),
                        dt=self.dt, pair_top=float(pair_top), i_coupling=int(bool(i_coupling)))
        for n in GRID_FIELDS:
            setattr(self._g, n, self.grid[n].ctypes.data_as(_dp))
        self._f = _Fields()
        for n in FIELDS:
            setattr(self._f, n, self.F[n].ctypes.data_as(_dp))

    def set(self, **arrays):
        for n, a in arrays.items():
            a = np.asarray(a, dtype=np.float64)
            assert a.shape == self.F[n].shape, (n, a.shape, self.F[n].shape)
            self.F[n][...] = a

    def _call(self, fn):
        getattr(lib(), fn)(ctypes.byref(self._g), ctypes.byref(self._f))

    def continuity(self):
        self._call('orc_continuity')

    def momentum(self):
        self._call('orc_momentum')

    def temperature(self):
        self._call('orc_temperature')

    def moisture(self):
        self._call('orc_moisture')

    def compute_tendencies(self):
        self._call('orc_compute_tendencies')

    def euler_forward(self):
        self._call('orc_euler_forward')

    def primary_diag(self):
        self._call('orc_primary_diag')

    def secondary_diag(self):
        self._call('orc_secondary_diag')

    def compute_turbulence(self):
        self._call('orc_compute_turbulence')

    def step_matsuno(self, nsteps=1):
        for _ in range(nsteps):
            self._call('orc_step_matsuno')

    @staticmethod
    def num_threads():
        return lib().orc_num_threads()

    @staticmethod
    def set_num_threads(n):
        lib().orc_set_num_threads(int(n))
