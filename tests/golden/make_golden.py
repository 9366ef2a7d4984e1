#!/usr/bin/env python
"""Regenerate the golden fixtures in this directory from the REAL reference.

Needs /root/reference (build container only).  Each fixture is produced by
oracle/run_reference.py, which runs the reference's own numba-CPU dynamical core
(dyn_matsuno.step_matsuno and the factories of dyn_org_discretizations.py) with the
import shims / namelist overrides documented in that script, and dumps its inputs and
outputs.  Nothing here is computed by this repository's code.

  ref_10deg_rand.npz  36x16x6 grid, random perturbations on (np.random.seed(3) of the
                      reference): grid fields, initial state, primary+secondary
                      diagnostics, EVERY stage-1 intermediate of compute_tendencies,
                      and state + diagnostics after 1, 2 and 10 Matsuno steps.
  ref_5deg.npz        BASELINE.json configs[0]: 5 deg x 8 levels, elev.1-deg topography,
                      Gaussian wind perturbation: grid, initial state, prognostic state
                      after 10 and 50 steps.
  ref_10deg_coupled.npz  the 36x16x6 grid with NON-ZERO physics coupling fields (seeded random
                      KMOM, KHEAT, SMOMXFLX, SMOMYFLX, SSHFLX, SLHFLX): inputs, every stage-1
                      intermediate incl. KMOM_dUWINDdz / KMOM_dVWINDdz and the *_TURB
                      tendencies, state after 1, 2 and 10 steps.
  ref_10deg_turb.npz  the 36x16x6 grid with the reference's turbulence module
                      (turb_compute.py) called after secondary_diag in every time step
                      (solver.py:106-112): KMOM / KHEAT of the first call, state and
                      KMOM / KHEAT after 1, 2 and 10 steps.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
RUN = os.path.join(HERE, '..', '..', 'oracle', 'run_reference.py')

JOBS = [
    ('ref_10deg_rand.npz', ['--grid', '10deg_rand', '--steps', '1', '2', '10', '--stage1',
                            '--dump-diag']),
    ('ref_5deg.npz', ['--grid', '5deg', '--steps', '10', '50', '--minimal']),
    ('ref_10deg_coupled.npz', ['--grid', '10deg_rand', '--steps', '1', '2', '10', '--stage1',
                               '--coupling']),
    ('ref_10deg_turb.npz', ['--grid', '10deg_rand', '--steps', '1', '2', '10', '--minimal',
                            '--turbulence']),
]

if __name__ == '__main__':
    only = sys.argv[1:]
    for name, args in JOBS:
        if only and name not in only:
            continue
        subprocess.check_call([sys.executable, RUN, '--out', os.path.join(HERE, name)] + args)
