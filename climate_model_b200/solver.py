"""Time loop of the model (reference: solver.py:40-255), dynamical core only.

    python -m climate_model_b200.solver [--nsteps N] [--output DIR | --no-output]
                                        [--restart-dir DIR] [--save-restart] [--load-restart]
                                        [name=value ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        -m climate_model_b200.solver ...          # N latitude bands, one GPU each

Builds Grid and ModelFields from the namelist (plus `name=value` overrides of grid / initial
condition parameters), runs one primary_diag, then per time step: print-diagnostics every
`nth_ts_print_diag` steps (vmax, mean wind / temperature / COLP, NaN / over-speed crash check,
reference io_functions.py:70-114, reduced on the device: io_functions.py), secondary_diag,
turbulence (KMOM / KHEAT, with `i_turbulence=1`: turb_main.py; the grid is then made with
i_coupling=1 and the step runs the kernel decomposition with the turbulent-transport terms),
step_matsuno -- all on the device; NetCDF output every `i_out_nth_hour` (io_nc_output.py) and
restart files every `i_restart_nth_day` (io_restart.py) as solver.py:152-176.  The other
physics modules (surface, radiation, microphysics) of the reference are out of scope.
"""
import argparse
import time

import torch

from . import namelist as nl
from .dyn_matsuno import Diagnostics, step_matsuno
from .io_functions import print_ts_info
from .io_read_namelist import B200
from .main_fields import ModelFields
from .main_grid import Grid

GRID_KEYS = ('nz', 'lat0_deg', 'lat1_deg', 'dlat_deg', 'dlon_deg', 'i_out_nth_hour',
             'i_sim_n_days', 'i_restart_nth_day', 'CFL', 'pair_top', 'i_moist_main_switch',
             'i_coupling')


def run(nsteps=None, verbose=True, ic=None, i_turbulence=None, output_path=None,
        output_fields=None, restart_dir='../restart', i_save_to_restart=None,
        i_load_from_restart=None, **overrides):
    """returns (GR, F) after the run; `overrides`: grid parameters, `ic`: initial-condition
    parameters (initialize_fields); `i_turbulence`, `i_save_to_restart`, `i_load_from_restart`:
    namelist overrides; `output_path`: directory of the NetCDF output (None: no output);
    `nsteps`: number of steps of THIS call (default: up to GR.nts)"""
    from .io_nc_output import constant_fields_to_NC, fields_for_output, output_to_NC
    from .io_restart import load_existing_fields, load_restart_grid, write_restart
    from .parallel_bands import attach_communicator, gather_field, init_bands
    i_turbulence = int(nl.i_turbulence if i_turbulence is None else i_turbulence)
    i_save_to_restart = int(nl.i_save_to_restart if i_save_to_restart is None
                            else i_save_to_restart)
    i_load_from_restart = int(nl.i_load_from_restart if i_load_from_restart is None
                              else i_load_from_restart)
    if i_turbulence:
        overrides.setdefault('i_coupling', 1)
    band = init_bands()                           # (0, 1) unless launched by torchrun
    quiet = not verbose                           # the crash check runs whatever `verbose` is
    verbose = verbose and band[0] == 0
    if i_load_from_restart:                       # main_grid.py:88-90, main_fields.py:61-62
        GR = load_restart_grid(overrides.get('dlat_deg', nl.dlat_deg),
                               overrides.get('dlon_deg', nl.dlon_deg),
                               overrides.get('nz', nl.nz), directory=restart_dir, band=band)
        F = load_existing_fields(GR, directory=restart_dir)
    else:
        GR = Grid(band=band, **{k: v for k, v in overrides.items() if k in GRID_KEYS})
        F = ModelFields(GR, **(ic or {}))
    if band[1] > 1:
        attach_communicator(GR, F)
    if output_path is not None and band[0] == 0:
        constant_fields_to_NC(GR, F, output_path=output_path)     # solver.py:66
    if i_turbulence:
        from .turb_main import Turbulence
        F.TURB = Turbulence(GR, target=B200)      # main_fields.py:84-86
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    nts = int(GR.nts) if nsteps is None else GR.ts + int(nsteps)
    ts0 = GR.ts
    t0 = time.time()
    while GR.ts < nts:
        GR.timer.start('total')
        GR.ts += 1
        GR.sim_time_sec = GR.ts * GR.dt
        print_ts_info(GR, F, force=(GR.ts == 1), quiet=quiet)   # collective with bands
        GR.timer.start('diag')
        Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
        GR.timer.stop('diag')
        if i_turbulence:                          # solver.py:106-112
            GR.timer.start('turb')
            F.TURB.compute_turbulence(GR, **F.get(F.TURB.fields_main, target=B200))
            GR.timer.stop('turb')
        step_matsuno(GR, F)
        if output_path is not None and GR.i_out_nth_ts and GR.ts % GR.i_out_nth_ts == 0:
            GR.timer.start('IO')                  # solver.py:152-164
            for n in fields_for_output(F, output_fields):
                gather_field(GR, F, n)            # bands: collected on rank 0
            GR.nc_output_count += 1
            if band[0] == 0:
                output_to_NC(GR, F, fields=output_fields, output_path=output_path)
            GR.timer.stop('IO')
        if i_save_to_restart and GR.i_restart_nth_ts and GR.ts % GR.i_restart_nth_ts == 0:
            GR.timer.start('IO')                  # solver.py:170-176
            write_restart(GR, F, directory=restart_dir, verbose=verbose)
            GR.timer.stop('IO')
        GR.timer.stop('total')
    if F.torch_device.type == 'cuda':
        torch.cuda.synchronize()
    print_ts_info(GR, F, force=True, quiet=quiet)
    if verbose:
        cells = int(GR.nx) * int(GR.ny) * int(GR.nz)
        dt = time.time() - t0
        n = GR.ts - ts0
        print('%d steps in %.2f s: %.3g cell-updates/s, %.1f x faster than reality' %
              (n, dt, cells * n / dt, n * GR.dt / dt))
        GR.timer.print_report()
    return GR, F


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument('--nsteps', type=int, default=None)
    ap.add_argument('--output', default=nl.output_path, help='NetCDF output directory')
    ap.add_argument('--no-output', action='store_true')
    ap.add_argument('--restart-dir', default='../restart')
    ap.add_argument('--save-restart', action='store_true')
    ap.add_argument('--load-restart', action='store_true')
    ap.add_argument('overrides', nargs='*', help='name=value namelist overrides')
    a = ap.parse_args()
    ov, ic = {}, {}
    for s in a.overrides:
        k, v = s.split('=', 1)
        v = float(v) if '.' in v or 'e' in v.lower() else int(v)
        (ov if k in GRID_KEYS or k == 'i_turbulence' else ic)[k] = v
    run(nsteps=a.nsteps, ic=ic, output_path=None if a.no_output else a.output,
        restart_dir=a.restart_dir, i_save_to_restart=int(a.save_restart) or None,
        i_load_from_restart=int(a.load_restart) or None, **ov)


if __name__ == '__main__':
    main()
