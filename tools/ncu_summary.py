#!/usr/bin/env python
"""One line per profiled launch of an `ncu --set full` report:
    python tools/ncu_summary.py report.ncu-rep "header comment" > profiles/<name>.csv
duration, DRAM bytes, DRAM throughput, registers, active warps, FP64 pipe, issue slots,
instructions, shared-memory wavefronts, L2 hit rate, top stall reasons (warp samples)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
M = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
     'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
     'sm__warps_active.avg.pct_of_peak_sustained_active',
     'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
     'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
     'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
     'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
idx = {m: h.index(m) for m in M if m in h}
stall = [(i, c[len('smsp__pcsamp_warps_issue_stalled_'):]) for i, c in enumerate(h)
         if c.startswith('smsp__pcsamp_warps_issue_stalled_') and not c.endswith('_not_issued')]
ki = h.index('Kernel Name')


def num(s):
    try:
        return float(s.replace(',', ''))
    except ValueError:
        return 0.


def scaled(m, row):
    v, u = num(row[idx[m]]), units[idx[m]]
    if m == 'gpu__time_duration.sum':
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1., 's': 1e3}.get(u, 1.)
    if m.startswith('dram__bytes'):
        v *= {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.}.get(u, 1.)
    return v


w = csv.writer(sys.stdout)
print('"# %s; units: duration ms, DRAM bytes Gbyte, the rest as named"' % note)
w.writerow(['Kernel Name'] + [m for m in M if m in idx] + ['top stalls'])
for row in data:
    tot = sum(num(row[i]) for i, _ in stall) or 1.
    top = sorted(((num(row[i]) / tot, n) for i, n in stall), reverse=True)[:5]
    name = row[ki].replace('dc::', '')
    w.writerow([name] + ['%.6g' % scaled(m, row) for m in M if m in idx] +
               [' '.join('%s:%.0f%%' % (n, 100 * f) for f, n in top)])
