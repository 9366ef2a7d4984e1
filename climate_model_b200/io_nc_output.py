"""NetCDF output (reference: io_nc_output.py:26-365 and io_functions.py:23-61).

`output_to_NC(GR, F)` writes `<output_path>/outNNNN.nc` with the reference's dimension and
variable names, dimension order (time, level[s], lat[s], lon[s]), 'f4' variables and the
zonal-mean profile variables; `constant_fields_to_NC` writes `constants.nc`.  The netCDF4
package is not part of this environment, so the files are classic NetCDF-3 written with
scipy.io.netcdf_file (same names and values; `testsuite.py` compares two such runs exactly
like the reference's testsuite compares its NETCDF4 files).

The fields are read from F.host: the caller copies the device state to the host first
(`fields_for_output` names what is needed, so that not all fields travel).
"""
import os

import numpy as np
from scipy.io import netcdf_file

from . import namelist as nl
from .io_constants import con_g
from .io_read_namelist import wp

# namelist.py:150-205 (output_fields): > 0 written, > 1 also as zonally averaged profile
output_fields = {
    'PSURF': 1, 'COLP': 1,
    'UWIND': 2, 'VWIND': 2, 'WIND': 2, 'WWIND': 2, 'VORT': 1,
    'POTT': 0, 'TAIR': 2, 'PHI': 1, 'PAIR': 0, 'RHO': 0,
    'SSHFLX': 1, 'SLHFLX': 1, 'SMOMXFLX': 0, 'SMOMYFLX': 0,
    'QV': 2, 'QC': 1, 'dQVdt': 1, 'dQVdt_TURB': 1, 'dPOTTdt_TURB': 0, 'dUFLXdt_TURB': 0,
    'dVFLXdt_TURB': 0, 'KMOM_dUWINDdz': 0, 'KMOM_dVWINDdz': 0, 'KHEAT': 0, 'KMOM': 1,
    'WVP': 0, 'CWP': 0,
}

# io_nc_output.py:82-89, minus the fields of the physics modules that are out of scope
# (QR, SOILMOIST, dPOTTdt_MIC, RAINRATE, ACCRAIN, RAIN)
DIRECT_FIELDS = ['UWIND', 'VWIND', 'WIND', 'POTT', 'TAIR', 'PHI', 'PAIR', 'RHO', 'COLP', 'QV',
                 'QC', 'dQVdt', 'dQVdt_TURB', 'dUFLXdt_TURB', 'dVFLXdt_TURB', 'dPOTTdt_TURB',
                 'KMOM', 'KHEAT', 'SMOMXFLX', 'SMOMYFLX', 'SSHFLX', 'SLHFLX', 'KMOM_dUWINDdz',
                 'KMOM_dVWINDdz']
_PROFILES = ['UWIND', 'VWIND', 'WWIND', 'VORT', 'POTT', 'TAIR', 'QV', 'QC']


def _wanted(F, name, fields):
    # the fused stepper forms the fluxes / tendencies of the reference's kernel decomposition in
    # registers and never stores them: such a field (dQVdt is in the default selection) is left
    # out of the file instead of being written as zeros; set_mode(GR, 'kernels') produces them
    GR = F._GR_ref() if getattr(F, '_GR_ref', None) else None
    if (name in F.KERNEL_MODE_ONLY and GR is not None and not GR.i_coupling and
            getattr(GR, '_mode', 'fused') == 'fused'):
        return 0
    return fields.get(name, 0) and (name in F.device or name in ('PSURF', 'VORT', 'WVP', 'CWP'))


def fields_for_output(F, fields=None):
    """device fields that output_to_NC reads with this selection (copy these to the host)"""
    fields = output_fields if fields is None else fields
    need = {'COLP'}
    for n in DIRECT_FIELDS + ['WWIND']:
        if _wanted(F, n, fields):
            need.add(n)
    if fields.get('VORT', 0):
        need.update(['UWIND', 'VWIND'])
    if fields.get('WVP', 0) or fields.get('CWP', 0):
        need.update(['PHIVB', 'RHO', 'QV', 'QC'])
    return sorted(n for n in need if n in F.device)


def NC_output_diagnostics(GR, UWIND, VWIND, WWIND, POTT, COLP, PVTF, PVTFVB, PHI, PHIVB, RHO,
                          QV, QC, want=('VORT', 'WWIND_ms', 'WVP', 'CWP')):
    """io_functions.py:23-61: relative vorticity, vertical wind in m/s, water vapour and cloud
    water path; a field that `want` does not name (or whose inputs are None) is None"""
    ii, jj = GR.ii, GR.jj
    VORT = WWIND_ms = WVP = CWP = None
    if 'VORT' in want and UWIND is not None and VWIND is not None:
        VORT = np.full((GR.nx + 2 * GR.nb, GR.ny + 2 * GR.nb, GR.nz), np.nan, dtype=wp)
        VORT[ii, jj, :] = (
            (+ (VWIND[ii + 1, jj, :] + VWIND[ii + 1, jj + 1, :]) / 2
             - (VWIND[ii - 1, jj, :] + VWIND[ii - 1, jj + 1, :]) / 2) / (2 * GR.dx[ii, jj, :])
            - (+ (UWIND[ii, jj + 1, :] + UWIND[ii + 1, jj + 1, :]) / 2
               - (UWIND[ii, jj - 1, :] + UWIND[ii + 1, jj - 1, :]) / 2) / (2 * GR.dy[ii, jj, :]))
    if 'WWIND_ms' in want and WWIND is not None and PHI is not None:
        WWIND_ms = WWIND.copy()
        ds = GR.dsigma[0, 0, :]
        WWIND_ms[ii, jj, 1:-1] = ((PHI[ii, jj, 1:] - PHI[ii, jj, :-1]) /
                                  (con_g * 0.5 * (ds[1:] + ds[:-1])) * WWIND[ii, jj, 1:-1])
    if ('WVP' in want or 'CWP' in want) and PHIVB is not None and RHO is not None:
        ALTVB = PHIVB / con_g
        dz = ALTVB[ii, jj, :-1] - ALTVB[ii, jj, 1:]
        WVP = np.sum(QV[ii, jj] * dz * RHO[ii, jj], 2)
        CWP = np.sum(QC[ii, jj] * dz * RHO[ii, jj], 2)
    return VORT, WWIND_ms, WVP, CWP


def _dimensions(ncf, GR, with_time):
    if with_time:
        ncf.createDimension('time', None)
    for n, v in (('lon', GR.nx), ('lons', GR.nxs), ('lat', GR.ny), ('lats', GR.nys),
                 ('level', GR.nz), ('levels', GR.nzs)):
        ncf.createDimension(n, int(v))
    if with_time:
        ncf.createVariable('time', 'f8', ('time',))[0] = GR.sim_time_sec / 3600 / 24
    nb = int(GR.nb)
    for n, a in (('lon', GR.lon_rad[GR.ii, nb + 1, 0]), ('lons', GR.lon_is_rad[GR.iis, nb + 1, 0]),
                 ('lat', GR.lat_rad[nb + 1, GR.jj, 0]), ('lats', GR.lat_js_rad[nb + 1, GR.jjs, 0]),
                 ('level', GR.level), ('levels', GR.levels)):
        ncf.createVariable(n, 'f4', (n,))[:] = np.asarray(a, dtype=np.float32).ravel()


def _need_coordinates(GR):
    if not hasattr(GR, 'lon_rad') or not hasattr(GR, 'level'):
        raise RuntimeError('this grid was built from dumped GRF arrays and has no coordinate '
                           'axes (lon_rad, lat_js_rad, level): NetCDF output needs a grid '
                           'made by create_new_grid')


def output_to_NC(GR, F, fields=None, output_path=None):
    """io_nc_output.py:26-318; returns the file name"""
    _need_coordinates(GR)
    fields = output_fields if fields is None else fields
    output_path = nl.output_path if output_path is None else output_path
    nx, nxs, ny, nys, nz, nzs, nb = (int(getattr(GR, n)) for n in
                                     ('nx', 'nxs', 'ny', 'nys', 'nz', 'nzs', 'nb'))
    H = F.host
    have = lambda n: n in F.device
    get = lambda n: H[n] if have(n) else None
    VORT, _, WVP, CWP = NC_output_diagnostics(
        GR, get('UWIND'), get('VWIND'), None, None, H['COLP'], None, None, None, get('PHIVB'),
        get('RHO'), get('QV'), get('QC'),
        want=[n for n in ('VORT',) if fields.get(n, 0)] +
             (['WVP', 'CWP'] if fields.get('WVP', 0) or fields.get('CWP', 0) else []))

    os.makedirs(output_path, exist_ok=True)
    filename = os.path.join(output_path, 'out' + str(GR.nc_output_count).zfill(4) + '.nc')
    ncf = netcdf_file(filename, 'w')
    _dimensions(ncf, GR, with_time=True)

    # DIRECT FIELDS (io_nc_output.py:82-128)
    for n in DIRECT_FIELDS:
        if not _wanted(F, n, fields):
            continue
        a = H[n]
        dimx, dimy, dimz = a.shape
        lon_str = 'lon' if dimx == nx + 2 * nb else 'lons'
        lat_str = 'lat' if dimy == ny + 2 * nb else 'lats'
        level_str = {nz: 'level', nzs: 'levels', 1: None}[dimz]
        inner = a[nb:dimx - 1, nb:dimy - 1, :].T            # (k, j, i)
        if level_str is None:
            ncf.createVariable(n, 'f4', ('time', lat_str, lon_str))[0] = inner[0]
        else:
            ncf.createVariable(n, 'f4', ('time', level_str, lat_str, lon_str))[0] = inner

    # PREPROCESSED FIELDS (io_nc_output.py:130-157)
    ii, jj = GR.ii, GR.jj
    if fields.get('PSURF', 0):
        ncf.createVariable('PSURF', 'f4', ('time', 'lat', 'lon'))[0] = \
            H['COLP'][ii, jj, 0].T + GR.pair_top
    if _wanted(F, 'WWIND', fields):
        ncf.createVariable('WWIND', 'f4', ('time', 'levels', 'lat', 'lon'))[0] = \
            (H['WWIND'][ii, jj, :] * H['COLP'][ii, jj, :]).T
    if fields.get('VORT', 0) and VORT is not None:
        ncf.createVariable('VORT', 'f4', ('time', 'level', 'lat', 'lon'))[0] = VORT[ii, jj, :].T
    if fields.get('WVP', 0) and WVP is not None:
        ncf.createVariable('WVP', 'f4', ('time', 'lat', 'lon'))[0] = WVP.T
    if fields.get('CWP', 0) and CWP is not None:
        ncf.createVariable('CWP', 'f4', ('time', 'lat', 'lon'))[0] = CWP.T

    # PROFILES (io_nc_output.py:160-213): zonal means, level axis reversed as in the reference
    for n in _PROFILES:
        if fields.get(n, 0) <= 1 or not (n == 'VORT' and VORT is not None or have(n)):
            continue
        a = VORT if n == 'VORT' else H[n]
        if n == 'UWIND':
            sel, lat_str = a[GR.iis, GR.jj, :], 'lat'
        elif n == 'VWIND':
            sel, lat_str = a[GR.ii, GR.jjs, :], 'lats'
        else:
            sel, lat_str = a[ii, jj, :], 'lat'
        level_str = 'levels' if a.shape[2] == nzs else 'level'
        ncf.createVariable(n + 'prof', 'f4', ('time', level_str, lat_str))[0] = \
            np.mean(sel, axis=0).T[::-1]
    ncf.close()
    return filename


def constant_fields_to_NC(GR, F, output_path=None):
    """io_nc_output.py:321-365"""
    _need_coordinates(GR)
    output_path = nl.output_path if output_path is None else output_path
    os.makedirs(output_path, exist_ok=True)
    filename = os.path.join(output_path, 'constants.nc')
    ncf = netcdf_file(filename, 'w')
    _dimensions(ncf, GR, with_time=False)
    ncf.createVariable('HSURF', 'f4', ('lat', 'lon'))[:] = F.host['HSURF'][GR.ii, GR.jj, 0].T
    ncf.close()
    return filename
