#!/bin/bash
# single-GPU measurements of round 2 (moisture tile kernel, single-GPU pipelining)
mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee gpurun_out/r2_pytest_gpu.log
for rep in 1 2; do
python tools/kbench.py --moist 1 2>&1 | tail -1
DC_MOIST_IMPL=1 python tools/kbench.py --moist 1 2>&1 | tail -1
python tools/kbench.py --no-profile 2>&1 | tail -1
DC_PIPELINE=1 python tools/kbench.py --no-profile 2>&1 | tail -1
DC_PIPELINE=1 python tools/kbench.py --no-profile --steps 20 2>&1 | tail -1
python tools/kbench.py --no-profile --steps 20 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2_single_kbench.log
