"""bench.py's output contract, checked without a GPU: `--impl reference` (the oracle port on
the host cores, the real reference arm) and the LOGIC of the B200 arm against the host
emulation (`--emu`; that line is tagged "emu": true and is never a measurement).  Exactly one
line on stdout, valid JSON, every key the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
             'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'e2e'}


def _bench(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args),
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, lines          # ONE JSON line, nothing else on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _bench('--impl', 'reference', '--port', '--workload', 'cfg1', '--steps', '2',
               '--warmup', '1')
    assert BASE_KEYS <= set(d) and d['impl'] == 'reference'
    assert d['metric'] == 'dyn-core cell-updates/s' and d['unit'] == 'cell-updates/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['dtype'] == 'f64' and d['data'] == 'synthetic' and 'workload' in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    c = d['config']
    assert (c['sample_nx'], c['sample_ny'], c['sample_nz']) == (72, 32, 8)


def test_reference_arm_runs_the_numba_reference_when_prepared():
    """with oracle/_ref prepared (build container: oracle/build_ref.py from /root/reference) the
    arm times the reference's own numba-CPU step, on every host thread even when the launcher
    exports OMP_NUM_THREADS=1 (torchrun does)"""
    import pytest
    if not os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'dyn_matsuno.py')):
        pytest.skip('oracle/_ref is not prepared')
    env = dict(os.environ, OMP_NUM_THREADS='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--workload', 'cfg1', '--steps', '2', '--warmup', '2'],
                       capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads(r.stdout.splitlines()[-1])
    cb = d['cpu_baseline']
    assert cb['kind'] == 'reference' and 'numba' in cb['sample'] and d['value'] > 0
    assert cb['cores'] == len(os.sched_getaffinity(0))
    # no product library in the reference arm's address space: it is a separate interpreter
    # that imports only oracle/ref_bench.py and the prepared reference tree
    src = open(os.path.join(ROOT, 'oracle', 'ref_bench.py')).read()
    assert 'climate_model_b200' not in src.replace('MEASUREMENT', '')


def test_b200_arm_logic_on_the_host_emulation():
    d = _bench('--emu', '--port', '--workload', 'cfg1', '--steps', '2', '--warmup', '1')
    assert d['emu'] is True                      # never mistaken for a measurement
    assert BASE_KEYS | {'roofline', 'step_roofline', 'gpu_launches', 'clocks', 'cpu_baseline',
                        'kernels_ms_per_step', 'ms_per_step_with_kernel_events'} <= set(d)
    assert d['n_gpus'] == 1 and d['steps'] == 2 and d['warmup'] == 1 and d['scaling'] == 'strong'
    assert d['config']['workload'].startswith('5deg') and d['config']['finite'] is True
    # ONE stepper call for the 2 timed steps: the x-halo fix once, then 2 steps x 2 stages x
    # (continuity, stage kernel, diagnostics)
    assert d['gpu_launches'] == 1 + 2 * 2 * 3
    e = d['e2e']
    # the headline e2e advances ONE host-resident state; the streamed ensemble is an extra
    assert e['h2d_bytes_per_step'] == e['d2h_bytes_per_step'] > 0 and e['value'] > 0
    assert e['what'].startswith('pinned host state') and e['repeats'] == 3
    st = e['streamed']
    assert st['members'] == 12 and st['value'] > 0 and st['finite'] is True
    assert d['step_roofline']['algorithmic_bytes_per_cell_update'] == 216
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['value'] > 0


def test_two_rank_line_under_torchrun_gloo():
    """the N > 1 arm exactly as the driver launches it (torchrun, one JSON line from rank 0),
    on the host emulation over gloo: latitude bands, max-over-ranks timing, band-shaped e2e"""
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29671', os.path.join(ROOT, 'bench.py'),
           '--gpus', '2', '--emu', '--workload', 'cfg1', '--steps', '2', '--warmup', '1',
           '--no-cpu-baseline']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d['n_gpus'] == 2 and d['emu'] is True
    assert d['config']['parallelism'] == 'latitude bands x2' and d['config']['finite'] is True
    e = d['e2e']
    assert e['what'].startswith('every rank: pinned band-shaped host state') and e['finite'] is True
    # the bands' rows plus two halo rows a side: a little more than the whole grid once
    whole = 8 * (75 * 34 * 8 + 74 * 35 * 8 + 74 * 34 * 8 + 74 * 34)
    assert whole < e['h2d_bytes_per_step'] < 1.3 * whole
    assert 'cpu_baseline' not in d


def test_clock_sampler_reports_the_rows_stamped_inside_the_timed_region():
    """the nvidia-smi log of the long-running sampler: only rows between begin() and stop()
    count; clocks, board power and throttle reasons come from those"""
    import datetime
    import importlib.util
    import tempfile
    import time
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.out, s.proc = tempfile.TemporaryFile(mode='w+'), None
    now = time.time()
    s.t0 = now - 0.2
    stamp = lambda t: datetime.datetime.fromtimestamp(t).strftime('%Y/%m/%d %H:%M:%S.%f')[:-3]
    idle = 'Not Active, Not Active, Not Active, Not Active'
    s.out.write('%s, 1965, 1965, 400.5, %s\n' % (stamp(now - 1.0), idle))        # warm-up
    s.out.write('%s, 1867, 1965, 990.1, Not Active, Not Active, Not Active, Active\n'
                % stamp(now - 0.1))
    s.out.write('%s, 1845, 1965, 995.0, Not Active, Not Active, Not Active, Active\n'
                % stamp(now - 0.05))
    s.out.flush()
    c = s.stop()
    assert c['in_timed_region'] == 2 and c['samples'] == 2
    assert c['sm_mhz'] in (1845.0, 1867.0) and c['sm_max_mhz'] == 1965.0
    assert c['reasons'] == ['sw_power_cap'] and c['power_w'] > 900
