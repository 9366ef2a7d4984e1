"""GPU parity at the BENCHMARKED shapes: both CUDA builds against the C oracle (bit-exact
restatement of the reference's numba-CPU path, tests/test_oracle_golden.py) on

  * a 1440 x 84 x 64 latitude band of BASELINE.json configs[3] (0.25 deg x 64 levels, the
    per-rank band of the 8-GPU run, the global grid's dt = 5 s), 10 and 50 steps, dry + moist;
  * a 3600 x 32 x 96 band of configs[4] (0.1 deg x 96 levels, dt = 2 s), 10 steps;
  * configs[1]/[2] in full (1 deg x 32 levels, elev.1-deg topography), 10 and 50 steps,
    dry and moist.

Metric and tolerances: tests/helpers.py (the reference testsuite's max|a-b|/max|b|,
testsuite.py:48-54).  The oracle's own 1-ulp-perturbation floor on these shapes
(tools/ulp_floor.py) is quoted next to TOL there.  The measured errors are also written to
gpurun_out/parity_bench_shapes.json.
"""
import json
import os

import numpy as np
import pytest

from helpers import STATE, TOL, state_err

pytestmark = pytest.mark.gpu

SHAPES = {
    'quarter_band': dict(dims=(1440, 84, 64, 5),
                         grid=dict(nz=64, lat0_deg=-10.5, lat1_deg=10.5, dlat_deg=0.25,
                                   dlon_deg=0.25, i_out_nth_hour=1.0, dt=5),
                         ic=dict(i_use_topo=0)),
    'tenth_band': dict(dims=(3600, 32, 96, 2),
                       grid=dict(nz=96, lat0_deg=-1.6, lat1_deg=1.6, dlat_deg=0.1, dlon_deg=0.1,
                                 i_out_nth_hour=1.0, dt=2),
                       ic=dict(i_use_topo=0)),
    'one_deg': dict(dims=(360, 168, 32, 20),
                    grid=dict(nz=32, lat0_deg=-84, lat1_deg=84, dlat_deg=1.0, dlon_deg=1.0),
                    ic=dict()),
}
CASES = [('quarter_band', 0, (10, 50)), ('quarter_band', 1, (10,)), ('tenth_band', 0, (10,)),
         ('one_deg', 0, (10, 50)), ('one_deg', 1, (10, 50))]

_oracle_cache = {}
_report = {}


def _initial_state(shape, moist):
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid
    sh = SHAPES[shape]
    GR = Grid(i_moist_main_switch=moist, **sh['grid'])
    assert (int(GR.nx), int(GR.ny), int(GR.nz), int(GR.dt)) == sh['dims']
    F = ModelFields(GR, **sh['ic'])
    return GR, F


def _oracle_states(shape, moist, steps, GR, F):
    """the oracle's prognostic state after each step count (computed once per case)"""
    key = (shape, moist)
    if key not in _oracle_cache:
        from oracle.oracle import GRID_FIELDS, Oracle
        O = Oracle(GR.nx, GR.ny, GR.nz, GR.dt, {n: GR.GRF['CPU'][n] for n in GRID_FIELDS},
                   i_moist=bool(moist))
        O.set(**{n: F.host[n] for n in ['HSURF'] + STATE})
        O.primary_diag()
        out, done = {}, 0
        for n in sorted(steps):
            O.step_matsuno(n - done)
            done = n
            out[n] = {m: O.F[m].copy() for m in STATE}
        _oracle_cache[key] = out
    return _oracle_cache[key]


@pytest.mark.parametrize('build', ['production', 'strict'])
@pytest.mark.parametrize('shape,moist,steps', CASES)
def test_benchmarked_shapes_against_oracle(shape, moist, steps, build):
    import torch
    from helpers import CUDA_LIB_STRICT
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    assert torch.cuda.is_available()
    _lib.use_library(CUDA_LIB_STRICT if build == 'strict' else _lib.DEFAULT_LIBRARY)
    try:
        assert _lib.is_cuda()
        GR, F = _initial_state(shape, moist)
        ref = _oracle_states(shape, moist, steps, GR, F)
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        done, bad = 0, []
        names = STATE if moist else STATE[:4]
        for n in sorted(steps):
            step_matsuno(GR, F, n - done)
            done = n
            F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
            errs = {m: state_err(m, F.host, ref[n]) for m in names}
            _report['%s moist=%d %s N=%d' % (shape, moist, build, n)] = errs
            bad += ['N=%d %s: %.3e > %.0e' % (n, m, e, TOL[m]) for m, e in errs.items()
                    if not (np.isfinite(e) and e <= TOL[m])]
        assert np.max(np.abs(F.host['UWIND'][1:-2, 1:-1])) > 5.      # a non-trivial flow
        assert not bad, '; '.join(bad)
    finally:
        _lib.use_library(_lib.DEFAULT_LIBRARY)
        try:
            os.makedirs('gpurun_out', exist_ok=True)
            with open('gpurun_out/parity_bench_shapes.json', 'w') as f:
                json.dump(_report, f, indent=1, sort_keys=True)
        except OSError:
            pass
