#!/usr/bin/env python
"""Kernel micro-benchmark: per-kernel device times of the Matsuno step on a synthetic state
filled directly on the device (no initial-condition builder), for quick A/B runs of library
build variants:  python tools/kbench.py [--lib path/to/libdyncore_x.so] [--steps 5] [--moist]
Not a bench.py number: it only prints the CUDA-event brackets of dc_profile_read."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

ap = argparse.ArgumentParser()
ap.add_argument('--lib', default=None)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--moist', type=int, default=0)
ap.add_argument('--mode', default='fused')
ap.add_argument('--nz', type=int, default=64)
ap.add_argument('--dlat', type=float, default=0.25)
ap.add_argument('--no-profile', action='store_true',
                help='plain step loop (no per-kernel event brackets): the step time only')
args = ap.parse_args()

import torch
from climate_model_b200 import _lib
_lib.use_library(args.lib or _lib.DEFAULT_LIBRARY)
from climate_model_b200.dyn_matsuno import Diagnostics, set_mode, step_matsuno
from climate_model_b200.io_read_namelist import B200
from climate_model_b200.main_fields import ModelFields
from climate_model_b200.main_grid import Grid

GR = Grid(nz=args.nz, lat0_deg=-84, lat1_deg=84, dlat_deg=args.dlat, dlon_deg=args.dlat,
          i_out_nth_hour=1.0, i_moist_main_switch=args.moist)
F = ModelFields(GR, initialize=False)
torch.manual_seed(0)
d = F.device
k = torch.arange(GR.nz, device=d['POTT'].device, dtype=torch.float64)[:, None, None]
d['POTT'].copy_(320. - 0.5 * k + torch.rand_like(d['POTT']))
d['COLP'].copy_(9.0e4 + 50. * torch.rand_like(d['COLP']))
d['UWIND'].copy_(5. * (torch.rand_like(d['UWIND']) - 0.5))
d['VWIND'].copy_(5. * (torch.rand_like(d['VWIND']) - 0.5))
d['QV'].copy_(0.005 * torch.rand_like(d['QV']))
L, h = _lib.lib(), GR.dyncore()
for n in ('UWIND', 'VWIND', 'POTT', 'COLP', 'QV', 'QC'):
    _lib.check(L.dc_exchange_bc(h, F.table[n][0], 0))
set_mode(GR, args.mode)
Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
step_matsuno(GR, F, 2)
torch.cuda.synchronize()
_lib.check(L.dc_profile_enable(h, 0 if args.no_profile else 1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step_matsuno(GR, F, args.steps)
e1.record()
torch.cuda.synchronize()
prof = {} if args.no_profile else _lib.profile_read(h)
ms = e0.elapsed_time(e1) / args.steps
cells = int(GR.nx) * int(GR.ny) * int(GR.nz)
print('%s: %.3f ms/step  %.2f Gcell/s  finite=%s  %s' % (
    os.path.basename(_lib.library_path()), ms, cells / ms / 1e6,
    bool(torch.isfinite(d['UWIND']).all().item()),
    {k: round(v[0] / args.steps, 3) for k, v in sorted(prof.items())}))
