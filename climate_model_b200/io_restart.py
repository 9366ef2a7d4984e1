"""Restart files (reference: io_restart.py:23-111).

Same entry points and file naming as the reference (`<dir>/<dlat>_<dlon>_<nz>.pkl`, one
pickle).  The reference pickles its Grid and ModelFields objects wholesale; here the objects
own a C handle and torch device buffers, so the file holds plain data instead: the grid's
parameters and time bookkeeping, and the host copies of the fields that define the model
state (prognostic fields, HSURF and -- on a grid with i_coupling -- the physics coupling
inputs).  Everything else is a diagnostic of that state: `load_existing_fields` uploads the
state and runs one primary_diag (solver.py:69-74 does the same before its time loop), after
which the run continues bit for bit as if it had not been interrupted.
"""
import os
import pickle

import numpy as np

from .io_read_namelist import B200
from .main_grid import GRID_FIELD_NAMES, Grid

STATE_FIELDS = ['UWIND', 'VWIND', 'POTT', 'COLP', 'QV', 'QC', 'HSURF']
COUPLING_INPUTS = ['KMOM', 'KHEAT', 'SMOMXFLX', 'SMOMYFLX', 'SSHFLX', 'SLHFLX', 'dPOTTdt_RAD']


def restart_file_name(dlat_deg, dlon_deg, nz, directory='../restart'):
    return os.path.join(directory, str(dlat_deg).zfill(2) + '_' + str(dlon_deg).zfill(2) + '_' +
                        str(nz).zfill(3) + '.pkl')


def _grid_key(GR):
    P = GR.params
    return P['dlat_deg'], P['dlon_deg'], int(GR.nz)


def write_restart(GR, F, directory='../restart', verbose=True):
    """io_restart.py:23-60; copies the state from the device first.  Latitude bands: the bands
    are gathered on rank 0, which writes the same single file a one-device run writes (a
    restart file does not depend on the number of ranks)"""
    from .parallel_bands import gather_field
    names = STATE_FIELDS + (COUPLING_INPUTS if GR.i_coupling else [])
    for n in names:
        gather_field(GR, F, n)
    filename = restart_file_name(*_grid_key(GR), directory=directory)
    if GR.band[0] != 0:
        return filename
    if verbose:
        print('###########################################')
        print('WRITE RESTART')
        print('###########################################')
    grid = {'params': dict(GR.params), 'ts': GR.ts,
            'sim_time_sec': GR.sim_time_sec, 'nc_output_count': GR.nc_output_count,
            'from_arrays': None}
    if not hasattr(GR, 'lon_rad'):        # a grid made from dumped GRF arrays (tests/golden)
        a = {n: np.array(getattr(GR, n)) for n in GRID_FIELD_NAMES}
        a.update(nx=int(GR.nx), ny=int(GR.ny), nz=int(GR.nz), dt=int(GR.dt))
        grid['from_arrays'] = a
    out = {'GR': grid, 'F': {n: np.array(F.host[n]) for n in names}}
    os.makedirs(directory, exist_ok=True)
    with open(filename + '.tmp', 'wb') as f:
        pickle.dump(out, f, protocol=pickle.HIGHEST_PROTOCOL)
    os.replace(filename + '.tmp', filename)      # never leave a half-written restart file
    return filename


def _load(filename):
    if not os.path.exists(filename):
        raise ValueError('Restart File does not exist.')
    with open(filename, 'rb') as f:
        return pickle.load(f)


def load_restart_grid(dlat_deg, dlon_deg, nz, directory='../restart', band=(0, 1)):
    """io_restart.py:64-75: the grid of the interrupted run incl. its time-step counter;
    `band` = (rank, nranks) of the run that continues (need not be that of the writer)"""
    g = _load(restart_file_name(dlat_deg, dlon_deg, nz, directory))['GR']
    GR = Grid(band=band, from_arrays=g['from_arrays'], **g['params'])
    GR.ts, GR.sim_time_sec, GR.nc_output_count = g['ts'], g['sim_time_sec'], g['nc_output_count']
    return GR


def load_existing_fields(GR, directory='../restart'):
    """io_restart.py:77-111: ModelFields holding the saved state, uploaded, diagnostics
    recomputed"""
    from .dyn_matsuno import Diagnostics
    from .main_fields import ModelFields
    saved = _load(restart_file_name(*_grid_key(GR), directory=directory))['F']
    F = ModelFields(GR, gpu_enable=True, initialize=False)
    for n, a in saved.items():
        if n not in F.device:
            raise ValueError('restart file holds %s, which this grid has no buffer for '
                             '(i_coupling differs?)' % n)
        F.host[n][...] = a
    F.host['WWIND'][...] = 0.            # io_initial_conditions.py:45-46
    F.host['POTTVB'][...] = 0.
    F.copy_host_to_device(GR, F.ALL_FIELDS)
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    return F
