"""Run-time diagnostics (reference: io_functions.py:70-114).

`diagnose_print_diag_fields` and the crash check of `print_ts_info`, reduced ON THE DEVICE
(dc_run_diag: two deterministic passes, no atomics): the reference copies WIND, COLP and POTT to
the host every `nth_ts_print_diag` steps, here 7 numbers per latitude row travel.  With latitude
bands the row results of the ranks are combined with one small all-reduce.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import namelist as nl


def _row_results(GR, F):
    """(7, rows of this band) array of dc_run_diag's row vectors"""
    L, h = _lib.lib(), GR.dyncore()
    n = ctypes.c_size_t()
    _lib.check(L.dc_run_diag_bytes(h, ctypes.byref(n)))
    nelem = n.value // 8
    if getattr(F, '_run_diag_scratch', None) is None or F._run_diag_scratch.numel() < nelem:
        F._run_diag_scratch = torch.empty(nelem, dtype=torch.float64, device=F.torch_device)
    st = F._run_diag_scratch
    t = F.device['UWIND']
    stream = torch.cuda.current_stream(t.device).cuda_stream if t.is_cuda else 0
    from .dyn_matsuno import _bind_all
    _bind_all(GR, F)
    _lib.check(L.dc_run_diag(h, st.data_ptr(), nelem * 8, stream))
    rows = st[nelem - 7 * int(GR.NJ):nelem].reshape(7, int(GR.NJ)).cpu().numpy()
    j0, j1, js = int(GR.j0), int(GR.j1), int(GR.jshift)
    return rows[:, j0 + js:j1 + js + 1]


def diagnose_print_diag_fields(GR, F):
    """(max_wind, mean_wind, mean_temp, mean_colp) as io_functions.py:70-93, plus the inputs of
    the crash check (max UWIND, number of NaNs in UWIND)"""
    r = _row_results(GR, F)
    sums = np.array([r[q].sum() for q in range(4)])
    maxs = np.array([r[4].max(), r[5].max()])
    nans = float(r[6].sum())
    if GR.band[1] > 1:
        import torch.distributed as dist
        dev = F.torch_device
        group = getattr(getattr(GR, 'comm', None), 'group', None)
        s = torch.tensor(np.append(sums, nans), dtype=torch.float64, device=dev)
        m = torch.tensor(maxs, dtype=torch.float64, device=dev)
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
        s, maxs = s.cpu().numpy(), m.cpu().numpy()
        sums, nans = s[:4], float(s[4])
    s_w, s_p, s_ca, s_a = sums
    nz = int(GR.nz)
    return (float(maxs[0]), float(s_w / s_ca / nz), float(s_p / s_ca / nz), float(s_ca / s_a),
            float(maxs[1]), nans)


def print_ts_info(GR, F, force=False, quiet=False):
    """io_functions.py:95-135: the diagnostics line every nth_ts_print_diag steps and the crash
    check (NaN in UWIND or UWIND > 500 m/s -> ValueError('MODEL CRASH')); `quiet` only suppresses
    the line, never the check (the reference checks unconditionally)"""
    if not force and GR.ts % nl.nth_ts_print_diag != 0:
        return None
    GR.timer.start('diag')
    vmax, mean_wind, mean_temp, mean_colp, umax, nans = diagnose_print_diag_fields(GR, F)
    GR.timer.stop('diag')
    if GR.band[0] == 0 and not quiet:
        print(str(GR.ts) + '  ' + str(np.round(GR.sim_time_sec / 3600 / 24, 3)) + '\t days' +
              ' vmax: ' + str(np.round(vmax, 1)) + '  m/s vmean: ' + str(np.round(mean_wind, 3)) +
              ' m/s Tmean: ' + str(np.round(mean_temp, 7)) + '  K  COLP: ' +
              str(np.round(mean_colp, 2)) + ' Pa', flush=True)
    if nans > 0 or umax > 500:
        raise ValueError('MODEL CRASH')
    return vmax, mean_wind, mean_temp, mean_colp
