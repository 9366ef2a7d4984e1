"""Multi-GPU latitude bands over NCCL (needs >= 2 GPUs; skipped on a single-GPU box):
the banded run on N GPUs must equal the single-GPU run bitwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import STATE, fields_from_golden, grid_from_golden, load_golden

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize('path', ['library', 'python'])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_nccl_banded_run_equals_single_gpu_bitwise(tmp_path, world, path):
    """path = 'library': the exchange, its overlap and the CUDA-graph replay run inside
    libdyncore (dc_set_comm + dc_step_matsuno on a band); 'python': dc_halo_pack -> torch
    NCCL send/recv -> dc_halo_unpack through the piecewise band entries"""
    if _ngpus() < world:
        pytest.skip('needs %d GPUs' % world)
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    _lib.use_library(_lib.DEFAULT_LIBRARY)
    fixture, nsteps, moist = 'ref_5deg.npz', 4, 1
    g = load_golden(fixture)
    GR = grid_from_golden(g, i_moist_main_switch=moist)
    F = fields_from_golden(GR, g)
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    step_matsuno(GR, F, nsteps)
    F.copy_device_to_host(GR, F.ALL_FIELDS)
    ref = {n: F.host[n].copy() for n in STATE + ['PHI', 'WWIND']}
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', BAND_BACKEND='nccl',
               DC_BAND_IN_LIBRARY='1' if path == 'library' else '0')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node=%d' % world, '--master-addr', '127.0.0.1', '--master-port',
           str(29700 + world + (10 if path == 'python' else 0)), os.path.join(HERE, 'band_worker.py'), fixture, str(nsteps),
           str(tmp_path), str(moist)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for rank in range(world):
        b = np.load(os.path.join(str(tmp_path), 'band%d.npz' % rank))
        j0, j1 = int(b['j0']), int(b['j1'])
        for n in STATE + ['PHI', 'WWIND']:
            a, e = b[n][:, j0:j1 + 1], ref[n][:, j0:j1 + 1]
            assert np.array_equal(a, e), 'rank %d %s: max|diff| %g' % (
                rank, n, np.nanmax(np.abs(a - e)))
