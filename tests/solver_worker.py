"""worker of tests/test_io.py: the solver time loop under torchrun (latitude bands over gloo,
host emulation), NetCDF output and restart files written by rank 0"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    cfg = json.loads(sys.argv[1])
    from helpers import build_emu
    from climate_model_b200 import _lib, solver
    _lib.use_library(build_emu())
    GR, F = solver.run(**cfg)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
