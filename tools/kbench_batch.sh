#!/bin/bash
# A/B runs of library build variants on one B200: tools/kbench.py for every .so under
# build/variants/ (git-ignored; built on the CPU container with different -D switches).
# Usage: gpurun --timeout 600 -- 'bash tools/kbench_batch.sh [extra kbench args]'
mkdir -p gpurun_out
: > gpurun_out/kbench_batch.log
for so in build/variants/*.so; do
    for rep in 1 2; do
        timeout 120 python tools/kbench.py --lib "$so" "$@" 2>&1 | tail -3 | tee -a gpurun_out/kbench_batch.log
    done
done
