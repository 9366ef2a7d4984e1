#!/usr/bin/env python
"""1-ulp-perturbation floor of the reference algorithm on the benchmarked shapes.

Runs the C oracle (bit-exact restatement of the reference's numba-CPU path) twice on the same
grid: once on the initial state, once with U, V, POTT multiplied by (1 + s * 2^-52), s random in
{-1, 0, 1} (the probe of oracle/run_reference.py --perturb-ulp 1), and
prints the reference testsuite's metric max|a-b|/max|b| per prognostic field after N steps.
The parity tolerances of tests/helpers.py must stay >= 20 x these numbers.

  python tools/ulp_floor.py --shape quarter --steps 10 50
TEST INFRASTRUCTURE (uses oracle/); CPU only."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

SHAPES = {
    # the per-rank band of BASELINE configs[3] at N = 8: 1440 x 84 x 64
    'quarter': dict(grid=dict(nz=64, lat0_deg=-10.5, lat1_deg=10.5, dlat_deg=0.25, dlon_deg=0.25,
                              i_out_nth_hour=1.0, dt=5), ic=dict(i_use_topo=0)),
    # a 32-row band of BASELINE configs[4]: 3600 x 32 x 96
    'tenth': dict(grid=dict(nz=96, lat0_deg=-1.6, lat1_deg=1.6, dlat_deg=0.1, dlon_deg=0.1,
                            i_out_nth_hour=1.0, dt=2), ic=dict(i_use_topo=0)),
    # BASELINE configs[1]/[2]
    'one': dict(grid=dict(nz=32, lat0_deg=-84, lat1_deg=84, dlat_deg=1.0, dlon_deg=1.0), ic=dict()),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='quarter', choices=sorted(SHAPES))
    ap.add_argument('--steps', type=int, nargs='+', default=[10, 50])
    args = ap.parse_args()
    from helpers import STATE, build_emu, state_err
    from climate_model_b200 import _lib
    _lib.use_library(build_emu())          # host emulation: only the field table is needed
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    from oracle.oracle import GRID_FIELDS, Oracle
    sh = SHAPES[args.shape]
    GR = Grid(i_moist_main_switch=1, **sh['grid'])
    F = ModelFields(GR, gpu_enable=False, device='cpu', **sh['ic'])
    runs = []
    for pert in (0, 1):
        O = Oracle(GR.nx, GR.ny, GR.nz, GR.dt, {n: GR.GRF['CPU'][n] for n in GRID_FIELDS})
        O.set(**{n: F.host[n] for n in ['HSURF'] + STATE})
        if pert:
            rng = np.random.default_rng(12345)
            for n in ('UWIND', 'VWIND', 'POTT'):
                s = rng.integers(-1, 2, size=O.F[n].shape)
                O.F[n][:] = O.F[n] * (1.0 + s * 2.0 ** -52)
        O.primary_diag()
        runs.append(O)
    done = 0
    print('%s: nx, ny, nz = %d, %d, %d  dt = %g' % (args.shape, GR.nx, GR.ny, GR.nz, GR.dt))
    for n_steps in sorted(args.steps):
        for O in runs:
            O.step_matsuno(n_steps - done)
        done = n_steps
        print('steps %3d: ' % n_steps + '  '.join(
            '%s %.1e' % (n, state_err(n, runs[1].F, runs[0].F)) for n in STATE), flush=True)


if __name__ == '__main__':
    main()
