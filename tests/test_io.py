"""NetCDF output, restart files and the testsuite comparison (reference
io_nc_output.py, io_restart.py, testsuite.py), on the host emulation of the library."""
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

from helpers import STATE, build_emu

GRID = dict(nz=6, lat0_deg=-80, lat1_deg=80, dlat_deg=10, dlon_deg=10, i_out_nth_hour=0.5)


@pytest.fixture(scope='module', autouse=True)
def emu_library():
    from climate_model_b200 import _lib
    prev = _lib.library_path()
    _lib.use_library(build_emu())
    yield
    if prev:
        _lib.use_library(prev)


def _run(tmp_path, nsteps, sub='out', **kw):
    from climate_model_b200 import solver
    ic = dict(UWIND_random_pert=2.0, VWIND_random_pert=2.0, POTT_random_pert=1.0,
              QV_random_pert=0.0005, COLP_random_pert=100.)
    return solver.run(nsteps=nsteps, verbose=False, ic=ic, output_path=str(tmp_path / sub),
                      restart_dir=str(tmp_path / 'restart'), **GRID, **kw)


def test_netcdf_output_names_dimensions_and_values(tmp_path):
    GR, F = _run(tmp_path, None, i_sim_n_days=1. / 24)
    assert GR.i_out_nth_ts * GR.dt == 1800 and GR.ts == 2 * GR.i_out_nth_ts
    assert GR.nc_output_count == 2
    out = tmp_path / 'out'
    assert sorted(os.listdir(out)) == ['constants.nc', 'out0001.nc', 'out0002.nc']
    nx, ny, nz = int(GR.nx), int(GR.ny), int(GR.nz)
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    with netcdf_file(str(out / 'out0002.nc'), 'r', mmap=False) as nc:
        assert {n: d for n, d in nc.dimensions.items()} == {
            'time': None, 'lon': nx, 'lons': nx + 1, 'lat': ny, 'lats': ny + 1, 'level': nz,
            'levels': nz + 1}
        v = nc.variables
        # namelist.py:150-205 selection, fields of the out-of-scope physics modules aside
        for n in ['UWIND', 'VWIND', 'WIND', 'WWIND', 'VORT', 'TAIR', 'PHI', 'COLP', 'PSURF', 'QV',
                  'QC', 'UWINDprof', 'VWINDprof', 'WWINDprof', 'TAIRprof', 'QVprof']:
            assert n in v, n
        assert 'POTT' not in v and 'RHO' not in v            # output_fields[...] == 0
        # the fused stepper never stores the tendencies: dQVdt is left out rather than
        # written as zeros (set_mode(GR, 'kernels') writes it)
        assert 'dQVdt' not in v
        assert v['UWIND'].dimensions == ('time', 'level', 'lat', 'lons')
        assert v['VWIND'].dimensions == ('time', 'level', 'lats', 'lon')
        assert v['WWIND'].dimensions == ('time', 'levels', 'lat', 'lon')
        assert v['COLP'].dimensions == ('time', 'lat', 'lon')
        assert v['UWINDprof'].dimensions == ('time', 'level', 'lat')
        assert v['UWIND'].data.dtype == np.dtype('>f4')
        assert v['time'][0] == GR.sim_time_sec / 3600 / 24
        f4 = lambda a: np.asarray(a, dtype=np.float32)
        assert np.array_equal(v['UWIND'][0], f4(F.host['UWIND'][1:-1, 1:-1, :].T))
        assert np.array_equal(v['VWIND'][0], f4(F.host['VWIND'][1:-1, 1:-1, :].T))
        assert np.array_equal(v['COLP'][0], f4(F.host['COLP'][1:-1, 1:-1, 0].T))
        assert np.array_equal(v['PSURF'][0], f4(F.host['COLP'][1:-1, 1:-1, 0].T + GR.pair_top))
        assert np.array_equal(v['WWIND'][0],
                              f4((F.host['WWIND'] * F.host['COLP'])[1:-1, 1:-1, :].T))
        # zonal-mean profile, level axis reversed (io_nc_output.py:189-213)
        prof = np.mean(F.host['TAIR'][1:-1, 1:-1, :], axis=0).T[::-1]
        assert np.array_equal(v['TAIRprof'][0], f4(prof))
        # vorticity against a direct evaluation of io_functions.py:27-41 at one point
        U, V = F.host['UWIND'], F.host['VWIND']
        i, j, k = 7, 5, 2
        ref = (((V[i + 1, j, k] + V[i + 1, j + 1, k]) / 2 - (V[i - 1, j, k] + V[i - 1, j + 1, k]) / 2)
               / (2 * GR.dx[i, j, 0])
               - ((U[i, j + 1, k] + U[i + 1, j + 1, k]) / 2 - (U[i, j - 1, k] + U[i + 1, j - 1, k]) / 2)
               / (2 * GR.dy[i, j, 0]))
        assert v['VORT'][0, k, j - 1, i - 1] == np.float32(ref)
    with netcdf_file(str(out / 'constants.nc'), 'r', mmap=False) as nc:
        assert np.array_equal(nc.variables['HSURF'][:],
                              np.asarray(F.host['HSURF'][1:-1, 1:-1, 0].T, dtype=np.float32))
        assert nc.variables['lat'][:].shape == (ny,)
    GR.close()


def test_restart_continues_bit_for_bit(tmp_path):
    from climate_model_b200.io_restart import restart_file_name, write_restart
    GR, F = _run(tmp_path, 3)
    fn = write_restart(GR, F, directory=str(tmp_path / 'restart'), verbose=False)
    assert fn == restart_file_name(10, 10, 6, str(tmp_path / 'restart')) and os.path.exists(fn)
    from climate_model_b200 import solver
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    for _ in range(2):                              # uninterrupted: 2 more steps
        GR.ts += 1
        Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
        step_matsuno(GR, F)
    F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
    want = {n: F.host[n].copy() for n in STATE}
    GR.close()
    GR2, F2 = _run(tmp_path, 2, sub='out2', i_load_from_restart=1)
    assert GR2.ts == 5 and GR2.sim_time_sec == 5 * GR2.dt
    F2.copy_device_to_host(GR2, F2.PROGNOSTIC_FIELDS)
    for n in STATE:
        assert np.array_equal(F2.host[n], want[n], equal_nan=True), n
    GR2.close()
    with pytest.raises(ValueError, match='does not exist'):
        solver.run(nsteps=1, verbose=False, restart_dir=str(tmp_path / 'nowhere'),
                   i_load_from_restart=1, **GRID)


def test_restart_written_by_the_time_loop(tmp_path):
    GR, F = _run(tmp_path, None, i_sim_n_days=1. / 24, i_restart_nth_day=0.5 / 24,
                 i_save_to_restart=1)
    assert GR.i_restart_nth_ts == GR.i_out_nth_ts
    assert os.listdir(tmp_path / 'restart') == ['10_10_006.pkl']
    GR.close()


def test_testsuite_comparison(tmp_path, capsys):
    from climate_model_b200.testsuite import compare_outputs
    GRa, Fa = _run(tmp_path, None, sub='a', i_sim_n_days=0.5 / 24)
    GRb, Fb = _run(tmp_path, None, sub='b', i_sim_n_days=0.5 / 24)
    a, b = str(tmp_path / 'a' / 'out0001.nc'), str(tmp_path / 'b' / 'out0001.nc')
    failed, dev_sum, devs = compare_outputs(a, b)
    assert not failed and dev_sum == 0 and 'Bitwise identical' in capsys.readouterr().out
    assert set(devs) == {'UWIND', 'VWIND', 'WWIND', 'COLP', 'PHI', 'QV', 'QC'}
    # a different run (other perturbation seed -> other state) must fail the comparison
    from climate_model_b200 import solver
    GRc, Fc = solver.run(nsteps=None, verbose=False, output_path=str(tmp_path / 'c'),
                         ic=dict(UWIND_random_pert=2.1), i_sim_n_days=0.5 / 24, **GRID)
    failed, _, devs = compare_outputs(a, str(tmp_path / 'c' / 'out0001.nc'), verbose=False)
    assert failed and devs['UWIND'] > 1e-4
    for g in (GRa, GRb, GRc):
        g.close()


def _torchrun_solver(world, port, **cfg):
    import json
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node=%d' % world, '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(here, 'solver_worker.py'), json.dumps(cfg)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def test_banded_solver_writes_the_same_output_and_restart(tmp_path):
    """the solver time loop on 3 latitude bands (torchrun, gloo): rank 0 gathers the bands and
    writes the same NetCDF file and the same restart file as the one-device run; a restart
    written by 3 ranks continues on 2 ranks bit for bit"""
    from climate_model_b200.testsuite import compare_outputs
    ic = dict(UWIND_random_pert=2.0, VWIND_random_pert=2.0, POTT_random_pert=1.0,
              QV_random_pert=0.0005, COLP_random_pert=100.)
    common = dict(verbose=True, ic=ic, i_sim_n_days=0.5 / 24, i_restart_nth_day=0.5 / 24,
                  i_save_to_restart=1, **GRID)
    GR, F = _run(tmp_path, None, sub='one', i_sim_n_days=0.5 / 24)
    nts = GR.ts
    GR.close()
    out = _torchrun_solver(3, 29731, output_path=str(tmp_path / 'three'),
                           restart_dir=str(tmp_path / 'restart3'), **common)
    assert out.count('vmax') >= 2 and 'WRITE RESTART' in out          # printed by rank 0 only
    one, three = str(tmp_path / 'one' / 'out0001.nc'), str(tmp_path / 'three' / 'out0001.nc')
    with netcdf_file(one, 'r', mmap=False) as a, netcdf_file(three, 'r', mmap=False) as b:
        assert sorted(a.variables) == sorted(b.variables)
        for n in a.variables:
            assert np.array_equal(a.variables[n][:], b.variables[n][:], equal_nan=True), n
    failed, dev_sum, _ = compare_outputs(one, three, verbose=False)
    assert not failed and dev_sum == 0
    # continue to the next output time: 2 bands from the 3-band restart == one device from the
    # same file == the uninterrupted one-device run
    from climate_model_b200 import solver
    cont = dict(nsteps=nts, verbose=False, i_load_from_restart=1, i_save_to_restart=0,
                restart_dir=str(tmp_path / 'restart3'), **GRID)
    _torchrun_solver(2, 29732, output_path=str(tmp_path / 'cont2'), **cont)
    GR1, F1 = solver.run(output_path=str(tmp_path / 'cont1'), **cont)
    assert GR1.ts == 2 * nts and GR1.nc_output_count == 2
    GR1.close()
    GR2, F2 = _run(tmp_path, None, sub='straight', i_sim_n_days=1. / 24)
    GR2.close()
    files = [str(tmp_path / d / 'out0002.nc') for d in ('cont2', 'cont1', 'straight')]
    for other in files[1:]:
        failed, dev_sum, devs = compare_outputs(files[0], other, verbose=False)
        assert not failed and dev_sum == 0 and len(devs) == 7, (other, devs)
