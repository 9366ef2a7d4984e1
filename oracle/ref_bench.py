#!/usr/bin/env python
"""Time the reference's own numba-CPU Matsuno step (oracle/_ref, prepared by
oracle/build_ref.py) on a sample grid and print one JSON line.

MEASUREMENT INFRASTRUCTURE for bench.py's `--impl reference` / `cpu_baseline` legs; never on the
product path.  Runs as its own process so that NUMBA_NUM_THREADS and the grid (CMREF_*
environment variables, read by the patched namelist at import) are fixed before the reference
is imported.  Follows SURVEY.md 8(d): dry configuration (physics modules off, coupling fields
zero), one primary_diag, warm-up steps to take the JIT out, then K timed step_matsuno calls.

  python oracle/ref_bench.py --steps 5 --warmup 1
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=1)
    args = ap.parse_args()
    if not os.path.exists(os.path.join(REF, 'dyn_matsuno.py')):
        print(json.dumps({'error': 'oracle/_ref is not prepared (oracle/build_ref.py)'}))
        return 2
    os.chdir(REF)                                   # the reference opens data/ relatively
    sys.path.insert(0, REF)
    t0 = time.time()
    import _interp2d_shim  # noqa: F401  (must precede io_initial_conditions)
    import numba
    import numpy as np
    from io_read_namelist import CPU, gpu_enable
    from main_grid import Grid
    from main_fields import ModelFields
    from dyn_matsuno import step_matsuno
    from dyn_org_discretizations import DiagnosticsFactory
    GR = Grid()
    F = ModelFields(GR, gpu_enable)
    for n in ('KMOM', 'KHEAT', 'SMOMXFLX', 'SMOMYFLX', 'SSHFLX', 'SLHFLX', 'dPOTTdt_RAD'):
        F.host[n][:] = 0.0
    D = DiagnosticsFactory(target=CPU)
    D.primary_diag(GR.GRF[CPU], **F.get(D.fields_primary_diag, target=CPU))
    D.secondary_diag(**F.get(D.fields_secondary_diag, target=CPU))
    for _ in range(max(1, args.warmup)):
        step_matsuno(GR, F)
    t_setup = time.time() - t0
    ts = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        step_matsuno(GR, F)
        ts.append(time.perf_counter() - t1)
    nx, ny, nz = int(GR.nx), int(GR.ny), int(GR.nz)
    ok = bool(np.isfinite(F.host['UWIND'][1:nx + 2, 1:ny + 1]).all())
    print(json.dumps({'nx': nx, 'ny': ny, 'nz': nz, 'dt': int(GR.dt), 'steps': args.steps,
                      'sec_per_step': ts, 'threads': int(numba.get_num_threads()),
                      'numba': numba.__version__, 'setup_and_jit_s': t_setup, 'finite': ok}))
    return 0


if __name__ == '__main__':
    sys.exit(main())
