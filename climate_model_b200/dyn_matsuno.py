"""Matsuno predictor/corrector step (reference: dyn_matsuno.py:28-129).

`step_matsuno(GR, F)` advances the device state by one full step with a single C call
(dc_step_matsuno): OLD <- current, then twice (tendencies -> COLP <- COLP_NEW ->
pressure-weighted Euler forward -> primary diagnostics), all enqueued on the current torch
stream with no host round trip.  `step_matsuno_factories` is the same sequence spelled out
through the factories exactly as the reference's Python does it (kernel-level parity tests).
"""
import torch

from . import _lib
from .dyn_org_discretizations import DiagnosticsFactory, PrognosticsFactory
from .dyn_tendencies import compute_tendencies
from .io_read_namelist import B200

import os

# development switch (default off: measured in profiles/r2_stage_variants.md)
_PIPELINE = os.environ.get('DC_PIPELINE', '0') == '1'

Prognostics = PrognosticsFactory(target=B200)
Diagnostics = DiagnosticsFactory(target=B200)


def _bind_all(GR, F):
    """bind F.device to the handle of GR (main_fields.bind_all: the key lives on the handle)"""
    from .main_fields import bind_all
    bind_all(GR, F)


def set_mode(GR, mode):
    """'fused' (default): continuity + one fused stage kernel + diagnostics per stage;
    'kernels': the reference's kernel decomposition, every intermediate field written
    (dc_set_mode, include/dyncore.h)"""
    code = {'fused': _lib.DC_MODE_FUSED, 'kernels': _lib.DC_MODE_KERNELS}[mode]
    _lib.check(_lib.lib().dc_set_mode(GR.dyncore(), code))
    GR._mode = mode


def step_matsuno(GR, F, nsteps=1):
    GR.timer.start('step')
    _bind_all(GR, F)
    GR._state_version = getattr(GR, '_state_version', 0) + 1
    t = F.device['UWIND']
    stream = torch.cuda.current_stream(t.device).cuda_stream if t.is_cuda else 0
    if GR.band[1] > 1:
        from .parallel_bands import step_matsuno_banded
        if getattr(GR, 'comm', None) is None:
            raise RuntimeError('latitude-band run: call parallel_bands.attach_communicator(GR, F) '
                               'after torch.distributed.init_process_group')
        step_matsuno_banded(GR, F, nsteps, stream)
    else:
        if _PIPELINE and _lib.is_cuda() and not getattr(GR, '_pipelined', False):
            # one GPU, two concurrent chains: the next continuity beside the diagnostics sweep
            # (dc_set_comm with one rank: no communicator, only the side stream and the graphs)
            GR._pipelined = True
            if not GR.i_coupling:
                _lib.check(_lib.lib().dc_set_comm(GR.dyncore(), None, 0, 0, 1))
        _lib.check(_lib.lib().dc_step_matsuno(GR.dyncore(), int(nsteps), stream))
    GR.timer.stop('step')


def step_matsuno_factories(GR, F):
    """dyn_matsuno.py:28-129 call for call, on target B200"""
    d = F.device
    GR.timer.start('step')
    for n in ('COLP', 'UWIND', 'VWIND', 'POTT') + (('QV', 'QC') if GR.i_moist_main_switch else ()):
        d[n + '_OLD'].copy_(d[n])
    GR.timer.stop('step')
    for _stage in ('estimate', 'final'):
        compute_tendencies(GR, F)
        d['COLP'].copy_(d['COLP_NEW'])
        GR.timer.start('step')
        Prognostics.euler_forward(GR, GR.GRF[B200],
                                  **F.get(Prognostics.fields_prognostic, target=B200))
        GR.timer.stop('step')
        GR.timer.start('diag')
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        GR.timer.stop('diag')
