"""Streaming of HOST-resident model states through the device (no reference counterpart).

The reference keeps one model state on the device and steps it in place; `step_matsuno(GR, F)`
does the same here.  When the states live in host memory -- ensemble members, or a caller that
owns host arrays and wants them advanced (the e2e leg of bench.py) -- a step is bound by the
PCIe link, not by the kernels: 1.5 GB up and 1.5 GB down per member at 0.25 deg x 64 levels is
~55 ms against a 5 ms step.  `MemberStream.advance` hides all it can of that: the upload of
member m+1, the step of member m and the download of member m-1 run concurrently on three
CUDA streams (both directions of the link busy at once), through double-buffered staging
buffers in the reference layout and ONE device-resident field set:

    upload stream:    host(m) --H2D--> staging_in[d]
    compute stream:   dc_import_field (layout transpose) -> primary_diag -> nsteps x
                      step_matsuno -> dc_export_field -> staging_out[d]
    download stream:  staging_out[d] --D2H--> host(m)            (d = m mod depth)

Members are independent, every member's state is uploaded, advanced and written back in place;
the result is bitwise what F.to_device / primary_diag / step_matsuno / F.to_host give one
member at a time (tests/test_gpu_parity.py, tests/test_emu_parity.py).  Host arrays should be
pinned (`pinned_member`), otherwise the copies are synchronous and nothing overlaps.
"""
import numpy as np
import torch

from . import _lib
from .dyn_matsuno import Diagnostics, step_matsuno
from .io_read_namelist import B200


def pinned_member(F, names):
    """{name: page-locked host array in the reference layout} for one member"""
    out = {}
    for n in names:
        shape = F.host[n].shape
        if F.torch_device.type == 'cuda':
            out[n] = torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
        else:
            out[n] = np.empty(shape, dtype=np.float64)
    return out


class MemberStream:

    def __init__(self, GR, F, names=None, depth=2):
        if GR.band[1] > 1:
            raise NotImplementedError('MemberStream works on one device holding the whole '
                                      'latitude range')
        self.GR, self.F = GR, F
        if names is None:
            names = ['UWIND', 'VWIND', 'POTT', 'COLP'] + (
                ['QV', 'QC'] if GR.i_moist_main_switch else [])
        self.names = list(names)
        self.depth = int(depth)
        dev = F.torch_device
        self.cuda = dev.type == 'cuda'
        numel = {n: int(np.prod(F.host[n].shape)) for n in self.names}
        mk = lambda: [{n: torch.empty(numel[n], dtype=torch.float64, device=dev)
                       for n in self.names} for _ in range(self.depth)]
        self.staging_in, self.staging_out = mk(), mk()
        self.bytes_per_member = 8 * sum(numel.values())
        if self.cuda:
            self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            ev = lambda: [torch.cuda.Event() for _ in range(self.depth)]
            self.in_done, self.imported, self.comp_done, self.out_done = ev(), ev(), ev(), ev()
            self._host_busy = {}      # host buffer address -> event of its last download

    def advance(self, members, nsteps=1):
        """upload, advance by `nsteps` Matsuno steps and write back every member of
        `members` (iterable of {name: host array}); returns the number of members"""
        GR, F, L = self.GR, self.F, _lib.lib()
        h = GR.dyncore()
        fid = {n: F.table[n][0] for n in self.names}
        cur = torch.cuda.current_stream(F.torch_device) if self.cuda else None
        cs = cur.cuda_stream if self.cuda else 0
        count = 0
        for m, host in enumerate(members):
            d = m % self.depth
            again = m >= self.depth
            flat = {n: torch.from_numpy(host[n]).view(-1) for n in self.names}
            # ---- upload
            if self.cuda:
                if again:
                    self.s_in.wait_event(self.imported[d])     # staging_in[d] consumed
                for n in self.names:      # a host buffer reused while its download is in flight
                    busy = self._host_busy.get(flat[n].data_ptr())
                    if busy is not None:
                        self.s_in.wait_event(busy)
                with torch.cuda.stream(self.s_in):
                    for n in self.names:
                        self.staging_in[d][n].copy_(flat[n], non_blocking=True)
                    self.in_done[d].record(self.s_in)
                cur.wait_event(self.in_done[d])
            else:
                for n in self.names:
                    self.staging_in[d][n].copy_(flat[n])
            # ---- compute
            for n in self.names:
                st = self.staging_in[d][n]
                _lib.check(L.dc_import_field(h, fid[n], st.data_ptr(), st.numel() * 8, cs))
            if self.cuda:
                self.imported[d].record(cur)
            Diagnostics.primary_diag(GR.GRF[B200],
                                     **F.get(Diagnostics.fields_primary_diag, target=B200))
            step_matsuno(GR, F, nsteps)
            if self.cuda and again:
                cur.wait_event(self.out_done[d])                # staging_out[d] downloaded
            for n in self.names:
                st = self.staging_out[d][n]
                _lib.check(L.dc_export_field(h, fid[n], st.data_ptr(), st.numel() * 8, cs))
            # ---- download
            if self.cuda:
                self.comp_done[d].record(cur)
                self.s_out.wait_event(self.comp_done[d])
                with torch.cuda.stream(self.s_out):
                    for n in self.names:
                        flat[n].copy_(self.staging_out[d][n], non_blocking=True)
                    self.out_done[d] = torch.cuda.Event()
                    self.out_done[d].record(self.s_out)
                for n in self.names:
                    self._host_busy[flat[n].data_ptr()] = self.out_done[d]
            else:
                for n in self.names:
                    flat[n].copy_(self.staging_out[d][n])
            count += 1
        if self.cuda:
            self.s_out.synchronize()
            self._host_busy.clear()
        return count
