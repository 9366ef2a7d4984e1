#!/usr/bin/env python
"""bench.py -- dyn-core throughput on B200: cell-updates/s and fraction of the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2|cfg3|cfg1|cfg5]
                  [--moist] [--impl reference]

One "step" = one full Matsuno predictor/corrector step of the dynamical core
(dyn_matsuno.step_matsuno: OLD copies, 2 x (tendencies, Euler forward, diagnostics, BCs))
over the whole grid; one cell-update = one mass cell advanced by one such step.
Default workload = BASELINE.json configs[3]: 0.25 deg x 0.25 deg, 64 sigma levels, synthetic
initial state (the configuration the north-star target is quoted on; 61.9 M cells, 0.5 GB
per 3-D field, far larger than L2).  With --gpus N the SAME grid is split into N latitude
bands (strong scaling), one process per GPU (torchrun), halos exchanged over NCCL.

`--impl reference` times the reference algorithm's CPU implementation (the oracle port,
oracle/dyncore_oracle.c, OpenMP over all host cores) on a bounded sample of the same
workload and prints the same JSON line with "impl": "reference".
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[0..3]
    'cfg1': dict(name='5deg x 8 levels (reference testsuite grid), topography',
                 grid=dict(nz=8, lat0_deg=-80, lat1_deg=80, dlat_deg=5, dlon_deg=5,
                           i_out_nth_hour=8), ic=dict()),
    'cfg2': dict(name='1deg x 32 levels, elev.1-deg topography, dry dyn core',
                 grid=dict(nz=32, lat0_deg=-84, lat1_deg=84, dlat_deg=1.0, dlon_deg=1.0), ic=dict()),
    'cfg3': dict(name='1deg x 32 levels, elev.1-deg topography, moist tracers (QV, QC)',
                 grid=dict(nz=32, lat0_deg=-84, lat1_deg=84, dlat_deg=1.0, dlon_deg=1.0), ic=dict(),
                 moist=True),
    'cfg4': dict(name='0.25deg x 64 levels, synthetic initial state (no topography), dry dyn core',
                 grid=dict(nz=64, lat0_deg=-84, lat1_deg=84, dlat_deg=0.25, dlon_deg=0.25,
                           i_out_nth_hour=1.0), ic=dict(i_use_topo=0)),
    # BASELINE.json configs[4], WEAK scaling: every rank holds 210 latitude rows of the 0.1 deg
    # grid (72.6 M cells, the per-rank share of the full 3600 x 1680 x 96 grid on 8 GPUs), so
    # N ranks cover lat +-10.5 N deg and N = 8 is the whole configs[4] grid (lat +-84 deg).
    # Every rank builds and keeps only its own rows (ModelFields(band_local=True)); dt = the
    # full grid's 2 s at every N.
    'cfg5': dict(name='0.1deg x 96 levels, synthetic initial state, 210 rows per GPU (weak scaling; '
                      'N = 8 is the full 3600 x 1680 x 96 grid)',
                 grid=dict(nz=96, lat0_deg=-10.5, lat1_deg=10.5, dlat_deg=0.1, dlon_deg=0.1,
                           i_out_nth_hour=1.0, dt=2), ic=dict(i_use_topo=0), weak=True),
    # development proxy (not a BASELINE config): two ranks of this grid each hold the 84 rows a
    # rank of cfg4 holds at N = 8
    'cfg4_band8x2': dict(name='0.25deg x 64 levels, +-21 deg (2 x 84 rows: per-rank size of cfg4 at N = 8)',
                         grid=dict(nz=64, lat0_deg=-21, lat1_deg=21, dlat_deg=0.25, dlon_deg=0.25,
                                   i_out_nth_hour=1.0), ic=dict(i_use_topo=0)),
}

# algorithmic bytes per cell-update of the WHOLE step (SURVEY.md 8d counting rule)
STEP_BYTES = {False: 216, True: 296}
# algorithmic 3-D field accesses (reads + writes) per cell of each kernel LAUNCH as the step is
# decomposed today; x 8 B = bytes per cell per launch
KERNEL_ACCESSES = {
    'continuity': (3, 5), 'continuity_fused': (3, 5), 'stage_fused': (11.5, 11.5),
    'primary_diag_fused': (4, 4),
    'moist_euler': (0, 6), 'moist_stage': (0, 9), 'uvflx_prep': (15, 15), 'uflx_tendency': (13, 13),
    'vflx_tendency': (13, 13), 'pott_tendency': (6, 6), 'moist_tendency': (7, 7),
    'euler_forward': (9, 15), 'primary_diag': (6, 6), 'copy_old': (6, 10),
}


# with the physics coupling terms (i_coupling): + KMOM, RHOVB, PHI reads and 2 writes in the
# preparation; + PHIVB, RHO, KMOM_d*WINDdz reads and the *_TURB write in the tendencies
KERNEL_ACCESSES_COUPLED = {
    'uvflx_prep': 20, 'uflx_tendency': 17, 'vflx_tendency': 17, 'pott_tendency': 13,
    'moist_tendency': 14, 'secondary_diag': 15, 'turbulence': 9,
}

# DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` capture, profiles/r2_ncu_full_summary.csv): valid for that workload on one
# GPU only; anything else reports null
NCU_TRAFFIC = {
    ('cfg4', 1, False, 'stage_fused'): 0.5 * ((4.03218 + 1.48306) + (5.54234 + 1.48088)) * 1e9,
    ('cfg4', 1, True, 'stage_fused'): 0.5 * ((4.03218 + 1.48306) + (5.54234 + 1.48088)) * 1e9,
}


def peak_hbm_gbs():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: ONE `nvidia-smi -lms`
    process (the recipe of B200_PROFILING.md) started before the warm-up, so that its start-up
    (NVML initialisation, device attach) does not fall into the timed region; the rows whose
    time stamp lies between begin() and stop() are the ones reported.  (A fresh nvidia-smi
    process every 200 ms, as in the first version, perturbed the run it observed: when a
    start-up landed on the 90 ms timed region the step came out 0.2 ms longer.)"""
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0, period_ms=50):
        self.t0 = self.t1 = None
        self.out = tempfile.TemporaryFile(mode='w+')
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', str(period_ms)],
                stdout=self.out, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()
        time.sleep(0.06)                      # one more period: a row stamped inside the region
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.out.seek(0)
        rows, inside = [], []
        for line in self.out.read().splitlines():
            r = [x.strip() for x in line.split(',')]
            if len(r) < 8 or not r[1].replace('.', '').isdigit():
                continue
            rows.append(r)
            try:
                ts = datetime.datetime.strptime(r[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                if self.t0 is not None and self.t0 <= ts <= self.t1 + 0.05:
                    inside.append(r)
            except ValueError:
                pass
        self.out.close()
        use = inside or rows[-2:]             # no stamp inside a very short region: the last rows
        sm = sorted(float(r[1]) for r in use)
        reasons = set()
        for r in use:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                'sw_power_cap'), r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        pw = sorted(float(r[3]) for r in use if r[3].replace('.', '').isdigit())
        return {'sm_mhz': sm[len(sm) // 2] if sm else None,
                'sm_max_mhz': float(use[0][2]) if use else None,
                'power_w': pw[len(pw) // 2] if pw else None,
                'samples': len(use), 'in_timed_region': len(inside), 'reasons': sorted(reasons)}


def sample_grid(wl):
    """bounded sample of the workload for the CPU arm: the SAME grid spacing, levels and
    initial-state recipe restricted to an equatorial band of <= 84 rows (the per-rank band of
    the 8-GPU run)"""
    g = dict(wl['grid'])
    ny_full = int(round((g['lat1_deg'] - g['lat0_deg']) / g['dlat_deg']))
    rows = min(ny_full, 84)
    half = rows * g['dlat_deg'] / 2
    if rows < ny_full:
        g['lat0_deg'], g['lat1_deg'] = -half, half
    return g, rows < ny_full, half


def host_threads():
    """host cores this process may use (torchrun exports OMP_NUM_THREADS=1: not a limit of
    the box, so the CPU arm sets its thread count itself)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_numba(wl, moist, steps, warmup=1):
    """the reference's own numba-CPU step (oracle/_ref, prepared by oracle/build_ref.py from
    /root/reference in the build container) timed in a subprocess on all host threads.
    Returns None when the prepared tree or numba is not there (the port is timed instead)."""
    ref = os.path.join(ROOT, 'oracle', '_ref')
    if not os.path.exists(os.path.join(ref, 'dyn_matsuno.py')):
        return None
    g, banded, half = sample_grid(wl)
    env = dict(os.environ)
    nthr = host_threads()
    env.update(CMREF_NZ=str(g['nz']), CMREF_LAT0_DEG=repr(float(g['lat0_deg'])),
               CMREF_LAT1_DEG=repr(float(g['lat1_deg'])), CMREF_DLAT_DEG=repr(float(g['dlat_deg'])),
               CMREF_DLON_DEG=repr(float(g['dlon_deg'])),
               CMREF_USE_TOPO=str(int(wl['ic'].get('i_use_topo', 1))),
               NUMBA_NUM_THREADS=str(nthr), OMP_NUM_THREADS=str(nthr),
               NUMBA_CACHE_DIR=os.path.join(ROOT, 'oracle', '_ref', '.numba_cache'))
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT'):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'ref_bench.py'),
                              '--steps', str(steps), '--warmup', str(warmup)], env=env,
                             capture_output=True, text=True, timeout=1500)
        d = json.loads(out.stdout.strip().splitlines()[-1])
        if 'sec_per_step' not in d or not d.get('finite'):
            raise RuntimeError(str(d))
    except Exception as exc:                                     # noqa: BLE001
        print('numba reference arm failed, timing the C port instead: %r' % (exc,),
              file=sys.stderr)
        return None
    cells = d['nx'] * d['ny'] * d['nz']
    sample = ('%dx%dx%d band (lat +-%.4g deg) of the workload grid, %d Matsuno steps of the '
              'reference numba-CPU dyn core (moisture tracers always on in the reference), '
              'numba %s, JIT + set-up %.0f s not timed' % (
                  d['nx'], d['ny'], d['nz'], half if banded else g['lat1_deg'], steps,
                  d['numba'], d['setup_and_jit_s']))
    return cells, d['sec_per_step'], d['threads'], sample, (d['nx'], d['ny'], d['nz']), 'reference'


def cpu_sample(wl, moist, steps, warmup=1):
    """the oracle port (oracle/dyncore_oracle.c, OpenMP) on the same bounded sample.  The
    state is built with the host-side Python of the package (grid + initial conditions, pure
    numpy); the product library libdyncore.so is NOT loaded by this arm."""
    import numpy as np
    from climate_model_b200.io_initial_conditions import initialize_fields
    from climate_model_b200.main_grid import Grid
    from oracle.oracle import FIELDS, GRID_FIELDS, Oracle, field_shape
    g, banded, half = sample_grid(wl)
    GR = Grid(i_moist_main_switch=int(moist), **g)
    nx, ny, nz = int(GR.nx), int(GR.ny), int(GR.nz)
    names = ['POTTVB', 'WWIND', 'HSURF', 'COLP', 'PVTF', 'PVTFVB', 'POTT', 'UWIND', 'VWIND',
             'QV', 'QC']
    host = {n: np.full(field_shape(n, nx, ny, nz), np.nan) for n in names}
    initialize_fields(GR, host, **wl['ic'])
    Oracle.set_num_threads(host_threads())
    O = Oracle(nx, ny, nz, GR.dt, {n: GR.GRF['CPU'][n] for n in GRID_FIELDS}, i_moist=moist)
    O.set(**{n: host[n] for n in ['HSURF', 'UWIND', 'VWIND', 'POTT', 'COLP', 'QV', 'QC']})
    O.primary_diag()
    O.step_matsuno(warmup)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.step_matsuno(1)
        ts.append(time.perf_counter() - t0)
    cells = nx * ny * nz
    assert np.isfinite(O.F['UWIND'][1:-2, 1:-1]).all()
    sample = '%dx%dx%d band (lat +-%.4g deg) of the workload grid, %d Matsuno steps' % (
        nx, ny, nz, half if banded else g['lat1_deg'], steps)
    return cells, ts, O.num_threads(), sample, (nx, ny, nz), 'port'


def cpu_arm(wl, moist, steps, warmup=1, prefer_numba=True):
    r = cpu_sample_numba(wl, moist, steps, warmup) if prefer_numba else None
    return r if r is not None else cpu_sample(wl, moist, steps, warmup)


_JSON_FD = None


def claim_stdout():
    """stdout carries ONE line, the JSON result: everything else a library writes to file
    descriptor 1 (NCCL prints its version there) is sent to stderr; emit() writes the line"""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + '\n').encode())


def run_reference(args, wl, moist):
    """`--impl reference`: the reference's CPU implementation of the path on the box's host
    cores.  Under torchrun only rank 0 works; the other ranks exit 0."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cells, ts, threads, sample, dims, kind = cpu_arm(wl, moist, args.steps,
                                                     max(1, min(args.warmup, 2)),
                                                     prefer_numba=not args.port)
    sec = sum(ts) / len(ts)
    value = cells / sec
    line = {
        'impl': 'reference', 'metric': 'dyn-core cell-updates/s', 'value': value,
        'unit': 'cell-updates/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak' if wl.get('weak') else 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': wl['name'], 'moist': moist, 'sample_nx': dims[0],
                   'sample_ny': dims[1], 'sample_nz': dims[2],
                   'note': 'bounded sample: an equatorial latitude band of the workload grid; '
                           'cell-updates/s does not depend on the number of rows'},
        'cpu_baseline': {'value': value, 'unit': 'cell-updates/s', 'cores': threads,
                         'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'cell-updates/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


PARITY_FIELDS = ['UWIND', 'VWIND', 'POTT', 'COLP']


def band_timeline(GR, F, L, h, rank, world, path):
    """--timeline: per-stream event timeline of ONE banded step (in-library exchange) on every
    rank, written as JSON; shows which segment of the step is exposed"""
    import torch
    import torch.distributed as dist
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import step_matsuno
    step_matsuno(GR, F, 1)
    dist.barrier()
    torch.cuda.synchronize()
    _lib.check(L.dc_profile_enable(h, 2))
    step_matsuno(GR, F, 1)
    marks = _lib.timeline_read(h)
    _lib.check(L.dc_profile_enable(h, 0))
    allm = [None] * world
    dist.all_gather_object(allm, marks)
    if rank == 0:
        with open(path, 'w') as f:
            json.dump({'what': 'one banded Matsuno step, ms since the step began on each rank '
                               '(M = main stream, S = side stream)', 'ranks': allm}, f, indent=1)



def band_hash(F, GR, j0, j1):
    """exact hash of the owned rows j0..j1 (interior columns) of the prognostic fields"""
    import torch
    out = []
    js, nx = GR.jshift, int(GR.nx)
    for n in PARITY_FIELDS + (['QV', 'QC'] if GR.i_moist_main_switch else []):
        a = F.device[n][:, j0 + js:j1 + js + 1, 1:nx + 1].contiguous()
        out.append(a.view(torch.int64).sum())
    return torch.stack(out)


def banded_parity(GR, F, wl, moist, steps_done, rank, world, cells):
    import torch
    import torch.distributed as dist
    from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid, band_rows
    if cells > 120e6:
        return {'checked': False, 'why': 'the whole grid does not fit the replay on one GPU'}
    mine = band_hash(F, GR, int(GR.j0), int(GR.j1))
    allh = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    res = torch.zeros(1, dtype=torch.int64, device=F.torch_device)
    if rank == 0:
        G1 = Grid(band=(0, 1), i_moist_main_switch=int(moist), **wl['grid'])
        F1 = ModelFields(G1, **wl['ic'])
        Diagnostics.primary_diag(G1.GRF[B200],
                                 **F1.get(Diagnostics.fields_primary_diag, target=B200))
        step_matsuno(G1, F1, steps_done)
        bad = 0
        for r in range(world):
            a, b = band_rows(int(G1.ny), r, world)
            if not torch.equal(band_hash(F1, G1, a, b), allh[r]):
                bad += 1
        res[0] = bad
        del F1
        G1.close()
        torch.cuda.empty_cache()
    dist.broadcast(res, 0)
    return {'checked': True, 'bitwise_equal': int(res.item()) == 0, 'bands_differing': int(res.item()),
            'steps': steps_done, 'fields': PARITY_FIELDS + (['QV', 'QC'] if moist else []),
            'how': 'int64-sum hash of the owned rows of every band vs a single-GPU replay of '
                   'the whole grid on rank 0'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default='cfg4', choices=sorted(WORKLOADS))
    ap.add_argument('--moist', action='store_true')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--port', action='store_true',
                    help='CPU arm: time the C/OpenMP oracle port even when the numba reference '
                         '(oracle/_ref) is available')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--e2e-members', type=int, default=12,
                    help='1 GPU: host-resident states streamed through ensemble_stream.MemberStream '
                         'for the e2e number (0: only the one-state-at-a-time sequence)')
    ap.add_argument('--mode', default='fused', choices=['fused', 'kernels'],
                    help='fused stage kernel (default) or one kernel per reference kernel')
    ap.add_argument('--turbulence', action='store_true',
                    help='secondary line, not the headline: the reference time loop with its '
                         'turbulence module on (secondary_diag -> compute_turbulence -> '
                         'step_matsuno with the turbulent-transport terms, SURVEY 8f-2); runs '
                         'the kernel decomposition')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-events', action='store_true',
                    help='do not bracket the kernels with CUDA events in the timed region (no '
                         'roofline object); checks that the brackets do not perturb the step')
    ap.add_argument('--timeline', default=None,
                    help='N > 1: write the per-stream event timeline of one banded step to this file')
    ap.add_argument('--emu', action='store_true',
                    help='debug the bench LOGIC on a box without a GPU against the host emulation '
                         '(tests/emu); the line is tagged "emu": true and is never a measurement')
    args = ap.parse_args()
    claim_stdout()
    wl = dict(WORKLOADS[args.workload])
    if wl.get('weak'):            # per-GPU work fixed: the latitude range grows with N
        wl['grid'] = dict(wl['grid'], lat0_deg=wl['grid']['lat0_deg'] * args.gpus,
                          lat1_deg=wl['grid']['lat1_deg'] * args.gpus)
    moist = bool(args.moist or wl.get('moist', False))
    if args.impl == 'reference':
        return run_reference(args, wl, moist)

    import numpy as np
    import torch
    import torch.distributed as dist
    from climate_model_b200 import _lib
    from climate_model_b200.dyn_matsuno import Diagnostics, set_mode, step_matsuno
    from climate_model_b200.io_read_namelist import B200
    from climate_model_b200.main_fields import ModelFields
    from climate_model_b200.main_grid import Grid

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.emu:
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        from helpers import build_emu
        _lib.use_library(build_emu())

        class _Ev:                                       # wall-clock stand-in for CUDA events
            def __init__(self, enable_timing=True):
                self.t = 0.

            def record(self):
                self.t = time.perf_counter()

            def elapsed_time(self, other):
                return (other.t - self.t) * 1e3
        torch.cuda.Event = _Ev
        torch.cuda.synchronize = lambda *a, **k: None
        if world > 1:
            dist.init_process_group('gloo')
    else:
        assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        _lib.use_library(_lib.DEFAULT_LIBRARY)
        assert _lib.is_cuda()
    assert world == args.gpus, '--gpus %d but WORLD_SIZE=%d (launch with torchrun)' % (
        args.gpus, world)

    # started here, seconds before the warm-up: its start-up must not touch the timed regions
    sampler = ClockSampler(local_rank) if rank == 0 else None
    GR = Grid(band=(rank, world), i_moist_main_switch=int(moist),
              i_coupling=int(args.turbulence), **wl['grid'])
    # latitude bands: host arrays and the initial-state builder cover the rank's rows only
    F = ModelFields(GR, band_local=world > 1, **wl['ic'])
    if args.turbulence:
        from climate_model_b200.turb_main import Turbulence
        F.TURB = Turbulence(GR, target=B200)
        args.e2e_members = 0
        args.no_cpu_baseline = True          # the CPU sample below times the dry step

    def one_step():
        if args.turbulence:                  # solver.py:99-112
            Diagnostics.secondary_diag(**F.get(Diagnostics.fields_secondary_diag, target=B200))
            F.TURB.compute_turbulence(GR, **F.get(F.TURB.fields_main, target=B200))
        step_matsuno(GR, F, 1)
    if world > 1:
        from climate_model_b200.parallel_bands import attach_communicator
        attach_communicator(GR, F)
    set_mode(GR, args.mode)
    cells = int(GR.nx) * int(GR.ny) * int(GR.nz)
    Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
    L, h = _lib.lib(), GR.dyncore()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    barrier()
    if sampler:
        sampler.begin()
    def timed(kernel_events):
        """K steps between barriers, CUDA events on the launching stream, max over ranks"""
        _lib.check(L.dc_profile_enable(h, 1 if kernel_events else 0))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        if args.turbulence:
            for _ in range(args.steps):
                one_step()
        else:
            # K steps = ONE call of the public stepper (dc_step_matsuno enqueues all of them
            # without a host round trip; on bands consecutive steps overlap inside the library)
            step_matsuno(GR, F, args.steps)
        e1.record()
        barrier()
        t_ms = e0.elapsed_time(e1)
        p = _lib.profile_read(h) if kernel_events else {}
        _lib.check(L.dc_profile_enable(h, 0))
        if world > 1:
            t = torch.tensor([t_ms], device=F.torch_device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms, p

    # Two timed regions of K steps each, back to back.  The first, straight after the W warm-up
    # steps as the contract says, is the plain step loop a user runs = `value`; the second
    # carries a CUDA-event bracket around every kernel launch (the roofline's launch durations).
    # On a board that reaches its power limit the second region runs at lower clocks than the
    # first (round 2: 4.49 against 4.70 ms/step in the other order), `clocks` covers both.
    launches0 = L.dc_launch_count(h)
    ms, _ = timed(False)
    launches = L.dc_launch_count(h) - launches0
    ms_events, prof = (None, {}) if args.no_kernel_events else timed(True)
    clocks = sampler.stop() if sampler else None
    timeline_steps = 0
    if args.timeline and world > 1 and getattr(GR.comm, 'in_library', False):
        band_timeline(GR, F, L, h, rank, world, args.timeline)
        timeline_steps = 2
    ms_per_step = ms / args.steps
    value = cells / (ms_per_step * 1e-3)
    js = GR.jshift   # owned rows, interior columns (halo cells beyond are never read)
    ok = bool(torch.isfinite(F.device['UWIND'][:, GR.j0 + js:GR.j1 + js + 1, 1:int(GR.nx) + 1])
              .all().item())

    # ---- N > 1: the banded result must equal the single-device result BITWISE.
    # Rank 0 replays the same number of steps on the whole grid on its own GPU; every rank
    # hashes the rows it owns (sum of the fp64 bit patterns as int64, wrapping) and the
    # hashes are compared field by field, band by band.
    parity = None
    steps_done = args.warmup + args.steps * (1 if args.no_kernel_events else 2) + timeline_steps
    if world > 1 and not args.turbulence:
        parity = banded_parity(GR, F, wl, moist, steps_done, rank, world, cells)

    # ---- end to end through the public field API: host state -> device -> step -> host
    names = ['UWIND', 'VWIND', 'POTT', 'COLP'] + (['QV', 'QC'] if moist else [])
    F.copy_device_to_host(GR, F.PROGNOSTIC_FIELDS)
    h2d = sum(F.host[n].nbytes for n in names)
    barrier()
    t_e2e = []
    for _ in range(args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        for n in names:
            F.to_device(GR, n)
        Diagnostics.primary_diag(GR.GRF[B200],
                                 **F.get(Diagnostics.fields_primary_diag, target=B200))
        one_step()
        for n in names:
            F.to_host(GR, n)
        barrier()
        t_e2e.append(time.perf_counter() - t0)
    e2e_sec = sum(t_e2e) / len(t_e2e) if t_e2e else float('nan')
    e2e_min = min(t_e2e) if t_e2e else float('nan')
    if world > 1:
        t = torch.tensor([e2e_sec], device=F.torch_device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    e2e_what = ('pinned host state -> device (+layout transpose), primary_diag, 1 Matsuno step, '
                'device -> host state')
    e2e_extra = {}
    if world > 1:
        # latitude bands: every rank's host arrays hold only the rows it holds
        # (ModelFields(band_local=True)), so the sequence above moved band rows only; the bytes
        # are summed over the ranks
        b = torch.tensor([float(h2d)], device=F.torch_device, dtype=torch.float64)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        h2d = int(b.item())
        e2e_what = ('every rank: pinned band-shaped host state (its rows + 2 halo rows a side) -> '
                    'device (+layout transpose), primary_diag, 1 banded Matsuno step (NCCL halos '
                    'inside the library), device -> host; bytes summed over ranks, mean of the '
                    'slowest rank')
        ja = F._rows_of(GR)('POTT')[0]
        own = F.host['POTT'][1:int(GR.nx) + 1, int(GR.j0) - ja:int(GR.j1) - ja + 1]
        fin = torch.tensor([float(bool(np.isfinite(own).all()))], device=F.torch_device,
                           dtype=torch.float64)
        dist.all_reduce(fin, op=dist.ReduceOp.MIN)
        e2e_extra = {'finite': fin.item() == 1.0}
    if world == 1 and args.e2e_members > 0:
        # host-resident states (ensemble members) streamed through the device: every member's
        # upload, step and download are inside the timed region; the upload of member m+1 and
        # the download of member m-1 overlap the step of member m (ensemble_stream.py)
        from climate_model_b200.ensemble_stream import MemberStream, pinned_member
        sets = []
        for _ in range(3):
            mem = pinned_member(F, names)
            for n in names:
                mem[n][...] = F.host[n]
            sets.append(mem)
        stream = MemberStream(GR, F, names=names, depth=2)
        stream.advance(sets, nsteps=1)                       # warm-up: one pass over the sets
        barrier()
        t0 = time.perf_counter()
        n_mem = stream.advance([sets[m % 3] for m in range(args.e2e_members)], nsteps=1)
        barrier()
        stream_sec = (time.perf_counter() - t0) / n_mem
        # the headline e2e stays the ONE-state sequence above (what the reference does: it
        # advances one state); the streamed ensemble is reported beside it
        e2e_extra = {'streamed': {
            'value': cells / stream_sec, 'ms_per_state': stream_sec * 1e3, 'members': n_mem,
            'finite': bool(all(np.isfinite(m['POTT']).all() for m in sets)),
            'what': '%d host-resident states (pinned, reference layout) streamed through the '
                    'device, each: H2D, layout transpose, primary_diag, 1 Matsuno step, '
                    'transpose, D2H; upload / step / download of consecutive states overlap '
                    '(ensemble_stream.MemberStream); wall clock / states' % n_mem}}

    if rank == 0:
        peak, peak_src = peak_hbm_gbs()
        # dominant kernel of the step, from the CUDA-event brackets of the timed region
        kern = {k: v for k, v in prof.items() if v[1] > 0}
        top = max(kern, key=lambda k: kern[k][0]) if kern else None
        roof = None
        if top:
            key = top + '_fused' if (top in ('continuity', 'primary_diag') and
                                     args.mode == 'fused') else top
            acc = KERNEL_ACCESSES.get(key, (0, 0))[1 if moist else 0]
            if args.turbulence:
                acc = KERNEL_ACCESSES_COUPLED.get(top, acc)
            # one "launch" of the roofline = one pass of the kernel over the band: with latitude
            # bands the stage kernel runs as two launches per stage (boundary tile rows on the
            # side stream, interior on the main stream), whose times are SUMMED here (they
            # overlap in wall time, so the sum overstates the elapsed time)
            passes = kern[top][1]
            if world > 1 and top == 'stage_fused':
                passes = 2 * args.steps
            k_ms = kern[top][0] / passes
            cells_launch = cells // world
            bytes_launch = acc * 8 * cells_launch
            ach = bytes_launch / (k_ms * 1e-3) / 1e9
            roof = {'bound': 'hbm', 'kernel': top, 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                    'frac': ach / peak,
                    'traffic': NCU_TRAFFIC.get((args.workload, world, moist, top))
                    if args.mode == 'fused' else None,
                    'traffic_source': 'ncu --set full, profiles/r2_ncu_full_summary.csv '
                                      '(mean of the stage-1 and stage-2 launch)',
                    'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': bytes_launch, 'avg_launch_ms': k_ms,
                    'share_of_step': kern[top][0] / sum(v[0] for v in kern.values())}
        step_gbs = STEP_BYTES[moist] * value / 1e9
        line = {
            **({'emu': True} if args.emu else {}),
            'metric': 'dyn-core cell-updates/s', 'value': value, 'unit': 'cell-updates/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step,
            'ms_per_step_with_kernel_events': None if ms_events is None else ms_events / args.steps,
            'higher_is_better': True, 'scaling': 'weak' if wl.get('weak') else 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'nx': int(GR.nx), 'ny': int(GR.ny),
                       'nz': int(GR.nz), 'dt_s': int(GR.dt), 'moist': moist,
                       'parallelism': 'latitude bands x%d' % world,
                       'mode': 'kernels + turbulence module (coupled terms)' if args.turbulence
                       else args.mode,
                       'l2': 'inputs larger than L2 (%.1f GB state)' % (h2d / 1e9),
                       'finite': ok},
            'roofline': roof,
            'step_roofline': {'bound': 'hbm', 'algorithmic_bytes_per_cell_update': STEP_BYTES[moist],
                              'achieved': step_gbs / world, 'peak': peak, 'unit': 'GB/s per GPU',
                              'frac': step_gbs / world / peak},
            'kernels_ms_per_step': {k: v[0] / args.steps for k, v in sorted(kern.items())},
            'e2e': {'value': cells / e2e_sec, 'unit': 'cell-updates/s',
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': h2d,
                    'ms_per_step': e2e_sec * 1e3, 'ms_per_step_min': e2e_min * 1e3,
                    'repeats': args.e2e_steps, 'what': e2e_what, **e2e_extra},
            'gpu_launches': int(launches),
            'clocks': clocks,
        }
        if parity is not None:
            line['parity_vs_1gpu'] = parity
        if not args.no_cpu_baseline and world == 1:
            c_cells, ts, threads, sample, _dims, kind = cpu_arm(wl, moist, steps=3,
                                                                prefer_numba=not args.port)
            sec = sum(ts) / len(ts)
            line['cpu_baseline'] = {'value': c_cells / sec, 'unit': 'cell-updates/s',
                                    'cores': threads, 'kind': kind, 'sample': sample}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
