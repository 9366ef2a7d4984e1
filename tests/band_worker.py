"""worker of tests/test_bands_gloo.py / test_gpu_bands.py: one process per latitude band
(BAND_BACKEND=gloo: host emulation on the CPU; nccl: the CUDA library, one GPU per rank)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _held_equal(F, GR, n, before):
    """device rows of the held range only (rows outside it are never touched by an import)"""
    ja, jb = F.held_rows(GR, n)
    js = int(GR.jshift)
    fnx = F.field_shape(n)[0]
    a = F.device[n][:, ja + js:jb + js + 1, :fnx]
    b = before[:, ja + js:jb + js + 1, :fnx]
    import torch
    return bool(torch.equal(a, b) or torch.equal(torch.nan_to_num(a, nan=1e300),
                                                 torch.nan_to_num(b, nan=1e300)))


def main():
    fixture, nsteps, outdir, moist = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    import torch.distributed as dist
    from helpers import STATE, build_emu, fields_from_golden, grid_from_golden, load_golden
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.parallel_bands import attach_communicator
    backend = os.environ.get('BAND_BACKEND', 'gloo')
    if backend == 'nccl':
        import torch
        torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
        dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
        _lib.use_library(_lib.DEFAULT_LIBRARY)
        assert _lib.is_cuda()
    else:
        dist.init_process_group('gloo')
        _lib.use_library(build_emu())
    rank, world = dist.get_rank(), dist.get_world_size()
    g = load_golden(fixture)
    GR = grid_from_golden(g, band=(rank, world), i_moist_main_switch=moist)
    F = fields_from_golden(GR, g)
    attach_communicator(GR, F)
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    # two calls: with the in-library exchange the first step of a handle runs plain, later
    # steps are replayed from the captured CUDA graph -- both paths are in the comparison
    step_matsuno(GR, F, 1)
    step_matsuno(GR, F, nsteps - 1)
    if backend == 'nccl':
        want = os.environ.get('DC_BAND_IN_LIBRARY', '1') != '0'
        assert GR.comm.in_library == want
        assert bool(_lib.lib().dc_has_comm(GR.dyncore())) == want
    from climate_model_b200.io_functions import diagnose_print_diag_fields
    out = {'j0': GR.j0, 'j1': GR.j1,
           'run_diag': np.array(diagnose_print_diag_fields(GR, F))}   # all-reduced over the bands
    for n in STATE + ['PHI', 'WWIND']:
        F.to_host(GR, n)
        out[n] = F.host[n]
    # band-shaped host I/O (ModelFields.to_host_band / to_device_band): the same rows as the
    # whole-grid path, and a lossless round trip into the device fields
    # (host emulation only: on GPUs this path is exercised by bench.py behind its pre-flight)
    for n in (['UWIND', 'VWIND', 'POTT', 'COLP'] if backend != 'nccl' else []):
        ja, jb = F.held_rows(GR, n)
        hb = F.band_buffer(GR, n)
        F.to_host_band(GR, n, hb)
        assert np.array_equal(hb.numpy(), F.host[n][:, ja:jb + 1, :], equal_nan=True), n
        before = F.device[n].clone()
        F.device[n].fill_(-3.)
        F.to_device_band(GR, n, hb)
        assert _held_equal(F, GR, n, before), n
    np.savez(os.path.join(outdir, 'band%d.npz' % rank), **out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
