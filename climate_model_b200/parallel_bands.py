"""Latitude-band decomposition across the GPUs of one node.

The grid is cut into contiguous latitude bands, one per rank / GPU; every rank keeps the
full longitude circle (the periodic boundary stays local) and full sigma columns, plus two
halo rows on each side.  A Matsuno stage needs the state two rows beyond the band, so the
only communication is, once per stage, the exchange of the two boundary rows of the new
U, V, POTT (QV, QC) and COLP with the north and south neighbours: grouped NCCL send/recv
(torch.distributed batch_isend_irecv; NVLink on a B200 box).  Ranks 0 and R-1 own the
rigid walls.  No reductions cross ranks, so the banded run is bitwise identical to the
single-device run.

The reference has no multi-device path; its archived multiprocessing version
exchanged longitude slabs at four points per stage.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class BandCommunicator:
    def __init__(self, GR, F, group=None):
        self.GR, self.F, self.group = GR, F, group
        self.rank, self.nranks = GR.band
        assert dist.is_initialized() and dist.get_world_size(group) == self.nranks
        assert dist.get_rank(group) == self.rank
        n = ctypes.c_size_t()
        _lib.check(_lib.lib().dc_halo_bytes(GR.dyncore(), ctypes.byref(n)))
        self.nelem = n.value // 8
        dev = F.torch_device
        self.south = self.rank - 1 if self.rank > 0 else None
        self.north = self.rank + 1 if self.rank < self.nranks - 1 else None

        def buf(present):
            return torch.empty(self.nelem, dtype=torch.float64, device=dev) if present else None
        self.send_s, self.recv_s = buf(self.south is not None), buf(self.south is not None)
        self.send_n, self.recv_n = buf(self.north is not None), buf(self.north is not None)
        self.comm_stream = None

    @staticmethod
    def _ptr(t):
        return t.data_ptr() if t is not None else None

    def _post(self, stage, stream):
        """pack on `stream`, then post the grouped send/recv with both neighbours"""
        L, h = _lib.lib(), self.GR.dyncore()
        _lib.check(L.dc_halo_pack(h, stage, self._ptr(self.send_s), self._ptr(self.send_n),
                                  stream))
        ops = []
        if self.south is not None:
            ops.append(dist.P2POp(dist.isend, self.send_s, self.south, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv_s, self.south, self.group))
        if self.north is not None:
            ops.append(dist.P2POp(dist.isend, self.send_n, self.north, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv_n, self.north, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def _finish(self, stage, works, stream):
        for w in works:
            w.wait()          # NCCL: the current stream waits; gloo: the host waits
        _lib.check(_lib.lib().dc_halo_unpack(self.GR.dyncore(), stage, self._ptr(self.recv_s),
                                             self._ptr(self.recv_n), stream))

    def stage(self, stage, stream):
        """one Matsuno stage on this band.  On CUDA the boundary tile rows, the packing and
        the NCCL exchange run on a second, high-priority stream concurrently with the
        interior tile rows on the main stream."""
        L, h = _lib.lib(), self.GR.dyncore()
        dev = self.F.torch_device
        P = _lib
        if dev.type != 'cuda':
            # host emulation (tests): same call sequence, no streams to overlap
            for part in (P.DC_PART_CONT, P.DC_PART_BOUNDARY):
                _lib.check(L.dc_stage_compute(h, stage, part, stream))
            works = self._post(stage, stream)
            for part in (P.DC_PART_INTERIOR, P.DC_PART_COLP):
                _lib.check(L.dc_stage_compute(h, stage, part, stream))
            self._finish(stage, works, stream)
        else:
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream(device=dev, priority=-1)
            main, side = torch.cuda.current_stream(dev), self.comm_stream
            _lib.check(L.dc_stage_compute(h, stage, P.DC_PART_CONT, stream))
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ready)    # also orders after the previous stage's unpack
                _lib.check(L.dc_stage_compute(h, stage, P.DC_PART_BOUNDARY, side.cuda_stream))
                boundary_done = torch.cuda.Event()
                boundary_done.record(side)
                works = self._post(stage, side.cuda_stream)
            _lib.check(L.dc_stage_compute(h, stage, P.DC_PART_INTERIOR, stream))
            main.wait_event(boundary_done)       # both stage-kernel parts have read COLP
            _lib.check(L.dc_stage_compute(h, stage, P.DC_PART_COLP, stream))
            self._finish(stage, works, stream)
        _lib.check(L.dc_stage_diag(h, stage, stream))


def attach_communicator(GR, F, group=None, in_library=None):
    """make step_matsuno(GR, F) run the banded step (call once after dist.init_process_group).

    On CUDA the library gets its OWN NCCL communicator (dc_set_comm): rank 0 draws the unique
    id (dc_comm_unique_id), torch.distributed broadcasts its 128 bytes, every rank joins.  The
    banded step is then ONE C call per step_matsuno (dc_step_matsuno: exchange, overlap and
    CUDA-graph replay inside the library).  `in_library=False` (or the host emulation, which
    has no NCCL) keeps the exchange in Python: dc_halo_pack -> torch.distributed
    batch_isend_irecv -> dc_halo_unpack, the piecewise band entries of include/dyncore.h."""
    import os
    GR.comm = BandCommunicator(GR, F, group)
    if in_library is None:
        in_library = _lib.is_cuda() and os.environ.get('DC_BAND_IN_LIBRARY', '1') != '0'
    GR.comm.in_library = False
    if in_library:
        L, h = _lib.lib(), GR.dyncore()
        nbytes = _lib.DC_COMM_ID_BYTES
        ident = torch.zeros(nbytes, dtype=torch.uint8)
        if GR.band[0] == 0:
            buf = (ctypes.c_ubyte * nbytes)()
            _lib.check(L.dc_comm_unique_id(buf, nbytes))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.to(F.torch_device)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(ident, src=src, group=group)
        raw = bytes(ident.cpu().tolist())
        _lib.check(L.dc_set_comm(h, raw, nbytes, GR.band[0], GR.band[1]))
        GR.comm.in_library = True
        GR.comm.p2p = False
        if os.environ.get('DC_BAND_P2P', '1') != '0':
            # peer-memory exchange: every rank publishes the CUDA IPC handles of its receive
            # buffers and flags, the neighbours map them (dc_comm_p2p_connect); the boundary rows
            # then travel by copy engine + stream memory operations, NCCL stays as the fallback
            nb = _lib.DC_P2P_HANDLE_BYTES
            mine = (ctypes.c_ubyte * nb)()
            ok = L.dc_comm_p2p_handles(h, mine, nb) == 0
            t = torch.tensor(list(mine), dtype=torch.uint8, device=F.torch_device)
            allh = [torch.zeros_like(t) for _ in range(GR.band[1])]
            dist.all_gather(allh, t, group=group)
            rank, nranks = GR.band
            south = bytes(allh[rank - 1].cpu().tolist()) if rank > 0 else None
            north = bytes(allh[rank + 1].cpu().tolist()) if rank < nranks - 1 else None
            ok = ok and L.dc_comm_p2p_connect(h, south, north, nb) == 0
            # all or nothing: a rank that cannot map its neighbours keeps everybody on NCCL
            # (the all-reduce is also the barrier after which every rank's flags are mapped)
            flag = torch.tensor([1.0 if ok else 0.0], device=F.torch_device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if flag.item() == 1.0:
                _lib.check(L.dc_comm_p2p_enable(h, 1))
                GR.comm.p2p = True
            elif rank == 0:
                import sys
                print('peer-memory halo exchange unavailable (%s): using NCCL send/recv'
                      % L.dc_last_error().decode(), file=sys.stderr)
    return GR.comm


def step_matsuno_banded(GR, F, nsteps, stream):
    L, h, comm = _lib.lib(), GR.dyncore(), GR.comm
    if getattr(comm, 'in_library', False):
        _lib.check(L.dc_step_matsuno(h, int(nsteps), stream))
        return
    for _ in range(int(nsteps)):
        _lib.check(L.dc_step_begin(h, stream))
        for stage in (0, 1):
            comm.stage(stage, stream)


# ---------------------------------------------------------------------------------------
# process set-up and gathers for the drivers (solver.py, io_nc_output.py, io_restart.py)
# ---------------------------------------------------------------------------------------
def init_bands():
    """(rank, nranks) of this process.  Under torchrun (WORLD_SIZE > 1) the process group is
    created here: NCCL with one GPU per rank when the library runs on CUDA, gloo for the host
    emulation of the tests; a plain `python -m climate_model_b200.solver` is (0, 1)."""
    import os
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1:
        return 0, 1
    if not dist.is_initialized():
        if _lib.is_cuda():
            local = int(os.environ.get('LOCAL_RANK', '0'))
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        else:
            dist.init_process_group('gloo')
    return dist.get_rank(), dist.get_world_size()


def owned_rows(GR, fny):
    """rows [ja, jb] of a reference-layout field with fny rows that THIS rank's results are
    authoritative for: its band, plus the wall-side halo rows on the first and last rank"""
    rank, nranks = GR.band
    return (0 if rank == 0 else int(GR.j0)), (fny - 1 if rank == nranks - 1 else int(GR.j1))


def gather_field(GR, F, name, dst=0):
    """device -> host on every rank (F.to_host), then the owned rows of all bands are
    collected in F.host[name] of rank `dst`: that host array is then the whole field, as on
    one device.  Point-to-point sends in rank order (output / restart path, not the step)."""
    F.to_host(GR, name)
    rank, nranks = GR.band
    if nranks == 1:
        return
    group = getattr(getattr(GR, 'comm', None), 'group', None)
    host = F.host[name]
    fny = host.shape[1]
    on_gpu = F.torch_device.type == 'cuda'

    def wire(a):        # NCCL moves device memory, gloo host memory
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(F.torch_device) if on_gpu else t

    if rank == dst:
        for src in range(nranks):
            if src == dst:
                continue
            j0, j1 = band_rows_of(GR, src)
            ja, jb = (0 if src == 0 else j0), (fny - 1 if src == nranks - 1 else j1)
            buf = wire(np.empty((host.shape[0], jb - ja + 1, host.shape[2]), dtype=host.dtype))
            dist.recv(buf, src=src, group=group)
            host[:, ja:jb + 1, :] = buf.cpu().numpy()
    else:
        ja, jb = owned_rows(GR, fny)
        dist.send(wire(host[:, ja:jb + 1, :]), dst=dst, group=group)


def band_rows_of(GR, rank):
    from .main_grid import band_rows
    return band_rows(int(GR.ny), rank, GR.band[1])
