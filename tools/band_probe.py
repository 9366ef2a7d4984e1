"""development probe: host-side enqueue time vs device time of the banded step
torchrun --nproc-per-node 2 tools/band_probe.py [lat_deg]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, torch.distributed as dist
from climate_model_b200 import _lib
from climate_model_b200.dyn_matsuno import Diagnostics, step_matsuno
from climate_model_b200.io_read_namelist import B200
from climate_model_b200.main_fields import ModelFields
from climate_model_b200.main_grid import Grid
from climate_model_b200.parallel_bands import attach_communicator

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
_lib.use_library(_lib.DEFAULT_LIBRARY)
lat = float(sys.argv[1]) if len(sys.argv) > 1 else 21.
GR = Grid(band=(rank, world), nz=64, lat0_deg=-lat, lat1_deg=lat, dlat_deg=0.25, dlon_deg=0.25,
          i_out_nth_hour=1.0)
F = ModelFields(GR, i_use_topo=0)
attach_communicator(GR, F)
Diagnostics.primary_diag(GR.GRF[B200], **F.get(Diagnostics.fields_primary_diag, target=B200))
for _ in range(5):
    step_matsuno(GR, F, 1)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
n = 100
t0 = time.perf_counter()
for _ in range(n):
    step_matsuno(GR, F, 1)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('rank %d: enqueue %.3f ms/step, total %.3f ms/step' % (rank, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3), flush=True)
dist.destroy_process_group()
