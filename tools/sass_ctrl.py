#!/usr/bin/env python
"""Print the scheduling control bits (stall count, write / read scoreboard, wait mask) beside
every SASS instruction of one kernel:  python tools/sass_ctrl.py lib.so <kernel-name-substring>
[lo hi]  (hex address range).  Which scoreboard a load signals and which instruction waits on
it is not in `cuobjdump -sass`'s text, only in the upper word of the encoding."""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
text = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
on = False
lines = text.split('\n')
i = 0
while i < len(lines):
    s = lines[i]
    if 'Function :' in s:
        on = pat in s
        if on:
            print(s.strip())
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/', s)
    if on and m and i + 1 < len(lines):
        m2 = re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/', lines[i + 1])
        if m2:
            c = (int(m2.group(1), 16) >> 41) & 0x7fffff
            a = int(m.group(1), 16)
            wb, rb = (c >> 5) & 7, (c >> 8) & 7
            if lo <= a <= hi:
                print('%04x %-64s st %2d wb %s rb %s wait %s' % (
                    a, m.group(2)[:64], c & 0xf, '-' if wb == 7 else wb, '-' if rb == 7 else rb,
                    format((c >> 11) & 0x3f, '06b')))
            i += 1
    i += 1
