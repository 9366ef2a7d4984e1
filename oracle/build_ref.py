#!/usr/bin/env python
"""Prepare oracle/_ref/: the reference (Potopoles/Climate_Model, numba CPU path) made runnable
for bench.py's `--impl reference` arm and `cpu_baseline` on the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is a directory of Python scripts (its
"compiled" form is numba's JIT output), so the recipe is: copy its *.py + data/ from where they
lie under /root/reference into oracle/_ref/ (git-ignored: never part of the history; it travels to
the GPU box like a built .so) and apply the behaviour-preserving import shims / numba-0.65 typing
fixes of oracle/run_reference.py.  The grid of the patched namelist is
read from CMREF_* environment variables at import time.  Run by __graft_entry__.build() when
/root/reference exists; the GPU box only uses the prepared tree (oracle/ref_bench.py).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')


def build(force=False):
    sys.path.insert(0, HERE)
    try:
        import run_reference as rr
    finally:
        sys.path.pop(0)
    if not os.path.isdir(rr.REF):
        return None                      # GPU box: the prepared tree (if any) is used as is
    stamp = os.path.join(DEST, '.prepared')
    src_m = max(os.path.getmtime(os.path.join(HERE, f)) for f in ('run_reference.py',
                                                                 'build_ref.py'))
    if not force and os.path.exists(stamp) and os.path.getmtime(stamp) >= src_m:
        return DEST
    rr.prepare_scratch(rr.GRIDS['1deg'], dest=DEST, env_grid=True)
    open(stamp, 'w').write('prepared from %s\n' % rr.REF)
    return DEST


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
