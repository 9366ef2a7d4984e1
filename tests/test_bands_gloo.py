"""Latitude-band decomposition on the CPU: world_size 2 and 3 over gloo, each rank running
the host emulation of the kernels on its band with the SAME Python orchestration, C entry
points (dc_stage_compute / dc_halo_pack / dc_halo_unpack / dc_stage_diag) and message layout
as the NCCL run on GPUs.  The assembled result must equal the single-band run BITWISE
(per-cell arithmetic is identical and nothing is reduced across ranks)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import (STATE, build_emu, fields_from_golden, grid_from_golden, load_golden)

HERE = os.path.dirname(os.path.abspath(__file__))


def _single(fixture, nsteps, moist):
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    prev = _lib.library_path()
    _lib.use_library(build_emu())
    g = load_golden(fixture)
    GR = grid_from_golden(g, i_moist_main_switch=moist)
    F = fields_from_golden(GR, g)
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    step_matsuno(GR, F, nsteps)
    from climate_model_b200.io_functions import diagnose_print_diag_fields
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    out = {n: F.host[n].copy() for n in STATE + ['PHI', 'WWIND']}
    out['run_diag'] = np.array(diagnose_print_diag_fields(GR, F))
    if prev:
        _lib.use_library(prev)
    return out


@pytest.mark.parametrize('world,fixture,moist', [(2, 'ref_10deg_rand.npz', 1),
                                                 (3, 'ref_10deg_rand.npz', 0),
                                                 (2, 'ref_5deg.npz', 1),
                                                 (2, 'ref_5deg.npz', 0)])
def test_banded_run_equals_single_band_bitwise(tmp_path, world, fixture, moist):
    nsteps = 3
    ref = _single(fixture, nsteps, moist)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node=%d' % world, '--master-addr', '127.0.0.1', '--master-port',
           str(29600 + world + 10 * moist), os.path.join(HERE, 'band_worker.py'), fixture,
           str(nsteps), str(tmp_path), str(moist)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    covered = 0
    for rank in range(world):
        b = np.load(os.path.join(str(tmp_path), 'band%d.npz' % rank))
        j0, j1 = int(b['j0']), int(b['j1'])
        covered += j1 - j0 + 1
        names = STATE[:4] + (STATE[4:] if moist else []) + ['PHI', 'WWIND']
        for n in names:
            a, e = b[n][:, j0:j1 + 1], ref[n][:, j0:j1 + 1]
            assert np.array_equal(a, e), 'rank %d %s: max|diff| %g' % (
                rank, n, np.nanmax(np.abs(a - e)))
        # device-side run diagnostics: maxima exact, the sums to rounding (other row grouping)
        d, e = b['run_diag'], ref['run_diag']
        assert d[0] == e[0] and d[4] == e[4] and d[5] == e[5] == 0.
        assert np.all(np.abs(d[1:4] - e[1:4]) <= 1e-13 * np.abs(e[1:4])), (d, e)
    assert covered == ref['POTT'].shape[1] - 2


def test_band_rows_partition():
    from climate_model_b200.main_grid import band_rows
    for ny in (16, 32, 168, 672, 1680, 7):
        for n in (1, 2, 3, 4, 8):
            rows = [band_rows(ny, r, n) for r in range(n)]
            assert rows[0][0] == 1 and rows[-1][1] == ny
            for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
                assert b0 == a1 + 1
            sizes = [b - a + 1 for a, b in rows]
            assert max(sizes) - min(sizes) <= 1
