"""ModelFields (reference: main_fields.py:30-487): registry of model fields, host copies
in the reference layout (i, j, k) and device buffers in the B200 layout F[k][jd][i]
(longitude fastest, see csrc/dc_geom.h), with the reference's get/set/copy API.

Device buffers are torch tensors (PyTorch owns the memory); the C library only keeps their
pointers.  Only the dyn-core fields have device buffers.  The physics coupling fields (KMOM,
KHEAT, surface fluxes and the *_TURB tendencies) get device buffers when the grid was made
with i_coupling=1; otherwise they exist on the host for API compatibility and must be zero
(dry configuration).
"""
import weakref

import numpy as np
import torch

from . import _lib
from .io_initial_conditions import initialize_fields, initialize_fields_band
from .io_read_namelist import CPU, wp

# host-only fields the reference's factories name (stgx, stgy, dimz)
_HOST_ONLY = {'PSURF': (0, 0, 1)}
# device-buffer address -> Grid that owns the handle the buffer is bound to; lets the
# factories keep the reference's signatures (which carry no grid object)
OWNERS = weakref.WeakValueDictionary()
# device-buffer address -> the ModelFields the buffer belongs to (the factories bind the WHOLE
# field set a tensor comes from, so that library work fields like PGCOL follow it)
FOWNERS = weakref.WeakValueDictionary()


def bind_all(GR, F):
    """Point the handle of GR at the device buffers of F.  The bindings live on the HANDLE, so
    the "already bound" key is kept on the Grid: two ModelFields on one Grid, or a factory call
    with foreign tensors in between, always end up advancing the field set they were given."""
    key = (id(F),) + tuple(t.data_ptr() for t in F.device.values())
    if getattr(GR, '_bound_key', None) != key:
        L = _lib.lib()
        h = GR.dyncore()
        for n, t in F.device.items():
            _lib.check(L.dc_bind_field(h, F.table[n][0], t.data_ptr(), t.numel() * 8))
        GR._bound_key = key


def fields_owner_of(fields):
    for t in fields.values():
        if isinstance(t, torch.Tensor):
            F = FOWNERS.get(t.data_ptr())
            if F is not None:
                return F
    return None


def owner_of(fields):
    for t in fields.values():
        if isinstance(t, torch.Tensor):
            GR = OWNERS.get(t.data_ptr())
            if GR is not None:
                return GR
    raise RuntimeError('none of the given tensors was allocated by ModelFields.allocate_device: '
                       'cannot find the libdyncore handle (target B200 needs F.device tensors)')


COUPLING_FIELDS = ['KMOM', 'KHEAT', 'SMOMXFLX', 'SMOMYFLX', 'SSHFLX', 'SLHFLX', 'dPOTTdt_RAD']


class ModelFields:

    ALL_FIELDS = 'all_fields'
    PRINT_DIAG_FIELDS = 'print_diag_fields'
    NC_OUT_DIAG_FIELDS = 'nc_out_diag_fields'
    PROGNOSTIC_FIELDS = 'prognostic_fields'

    def __init__(self, GR, gpu_enable=True, device=None, initialize=True, band_local=False,
                 **ic_overrides):
        """band_local=True (latitude-band runs): F.host[n] holds only the rows this rank holds
        -- shape (fnx, held rows, nk), rows self.held_rows(GR, n) of the reference layout --
        the initial state is built band by band (initialize_fields_band) and host <-> device
        copies move those rows (dc_import_rows / dc_export_rows): nothing whole-grid sized in 3-D
        is allocated on any rank.  Default: the reference's whole-grid host arrays."""
        self.gpu_enable = gpu_enable
        self.band_local = bool(band_local)
        self._GR_ref = weakref.ref(GR)
        self.host, self.fdict = allocate_fields(GR, self._rows_of(GR) if band_local else None)
        self.device = {}
        self.set_field_groups()
        if device is None:
            device = 'cuda' if _lib.is_cuda() else 'cpu'
        self.torch_device = torch.device(device)
        if gpu_enable and _lib.is_cuda() and self.torch_device.type != 'cuda':
            raise RuntimeError('libdyncore runs on CUDA devices only; there is no CPU fallback')
        if initialize and self.band_local:
            # (device POTTVB / WWIND start as zeros, PVTF / PVTFVB are diagnostics: no host copies)
            initialize_fields_band(GR, self.host, self._rows_of(GR),
                                   diagnostics=not gpu_enable, **ic_overrides)
        elif initialize:
            initialize_fields(GR, self.host, **ic_overrides)
        self._bound = {}
        if gpu_enable:
            self.allocate_device(GR)
            self.copy_host_to_device(GR, field_group=self.ALL_FIELDS)

    def set_field_groups(self):
        self.field_groups = {
            self.ALL_FIELDS: list(self.host.keys()),
            self.PRINT_DIAG_FIELDS: ['COLP', 'WIND', 'POTT'],
            self.NC_OUT_DIAG_FIELDS: ['UWIND', 'VWIND', 'WWIND', 'POTT', 'COLP', 'PVTF',
                                      'PVTFVB', 'PHI', 'PHIVB', 'RHO', 'QV', 'QC'],
            self.PROGNOSTIC_FIELDS: ['UWIND', 'VWIND', 'POTT', 'COLP', 'QV', 'QC'],
        }

    # ------------------------------------------------------------ reference API
    def get(self, field_names, target=CPU):
        src = self.host if target == CPU else self.device
        return {n: src[n] for n in field_names if n in src or target == CPU}

    def set(self, field_dict, target=CPU):
        dst = self.host if target == CPU else self.device
        for n, a in field_dict.items():
            dst[n] = a

    def copy_host_to_device(self, GR, field_group):
        GR.timer.start('copy')
        for n in self.field_groups[field_group]:
            if n in self.device and dict.__contains__(self.host, n):
                self.to_device(GR, n)
        GR.timer.stop('copy')

    def copy_device_to_host(self, GR, field_group):
        GR.timer.start('copy')
        for n in self.field_groups[field_group]:
            if n in self.device:
                self.to_host(GR, n)
        GR.timer.stop('copy')

    # ------------------------------------------------------------ B200 layout
    def allocate_device(self, GR):
        """zero-filled device buffers [nk][NJ][NI] for every field of the library's registry,
        bound to the handle (dc_bind_field)"""
        h = GR.dyncore()
        L = _lib.lib()
        self.table = _lib.field_table()
        for n, (fid, sx, sy, nkk) in self.table.items():
            if n in _lib.COUPLING_ONLY_FIELDS and not GR.i_coupling:
                continue      # 8 more 3-D fields that the dry configuration never touches
            nk = {_lib.DC_NK_2D: 1, _lib.DC_NK_NZ: int(GR.nz), _lib.DC_NK_NZS: int(GR.nz) + 1}[nkk]
            t = torch.zeros((nk, GR.NJ, GR.NI), dtype=torch.float64, device=self.torch_device)
            self.device[n] = t
            OWNERS[t.data_ptr()] = GR
            FOWNERS[t.data_ptr()] = self
            _lib.check(L.dc_bind_field(h, fid, t.data_ptr(), t.numel() * 8))
        GR._bound_key = None
        bind_all(GR, self)

    def _stage(self, nelem):
        """device scratch holding one field in the reference layout (raw copy of the host
        array); dc_import_field / dc_export_field transpose between it and the bound field"""
        if getattr(self, '_staging', None) is None or self._staging.numel() < nelem:
            self._staging = torch.empty(nelem, dtype=torch.float64, device=self.torch_device)
        return self._staging[:nelem]

    def _stream(self):
        if self.torch_device.type == 'cuda':
            return torch.cuda.current_stream(self.torch_device).cuda_stream
        return 0

    def _rows_of(self, GR):
        """name -> (ja, jb): rows of the reference layout this rank holds (band + halo rows)"""
        j0, j1, ny = int(GR.j0), int(GR.j1), int(GR.ny)
        table = dict(_lib.field_table())
        table.update({k: (None, v[0], v[1], v[2]) for k, v in _HOST_ONLY.items()})

        def rows(n):
            fny = ny + 2 + table[n][2]
            return max(0, j0 - 2), min(j1 + 3, fny - 1)
        return rows

    def to_device(self, GR, n):
        """host (i, j, k) -> device F[k][jd][i]: one contiguous H2D copy + an on-device tiled
        transpose (dc_import_field).  The copy from a pinned host array is asynchronous on the
        current stream: do not modify F.host[n] before the stream has passed it (any later
        to_host / synchronize does)"""
        h = self.host[n]
        bind_all(GR, self)
        if self.band_local:
            ja, jb = self._rows_of(GR)(n)
            src = torch.from_numpy(h).view(-1)
            st = self._stage(src.numel())
            st.copy_(src, non_blocking=True)
            _lib.check(_lib.lib().dc_import_rows(GR.dyncore(), self.table[n][0], st.data_ptr(),
                                                 st.numel() * 8, ja, jb, self._stream()))
            return
        src = torch.from_numpy(h).view(-1)
        st = self._stage(src.numel())
        st.copy_(src, non_blocking=True)
        _lib.check(_lib.lib().dc_import_field(GR.dyncore(), self.table[n][0], st.data_ptr(),
                                              st.numel() * 8, self._stream()))

    def to_host(self, GR, n):
        """device -> host, the reverse of to_device; synchronises the stream at the end"""
        h = self.host[n]
        bind_all(GR, self)
        self._refresh_for_export(GR, n)
        if self.band_local:
            ja, jb = self._rows_of(GR)(n)
            dst = torch.from_numpy(h).view(-1)
            st = self._stage(dst.numel())
            st.copy_(dst, non_blocking=True)     # rows outside the held range keep their values
            _lib.check(_lib.lib().dc_export_rows(GR.dyncore(), self.table[n][0], st.data_ptr(),
                                                 st.numel() * 8, ja, jb, self._stream()))
            dst.copy_(st, non_blocking=True)
            if self.torch_device.type == 'cuda':
                torch.cuda.current_stream(self.torch_device).synchronize()
            return
        if GR.band[1] > 1:
            # only the rows this rank holds cross the bus (dc_export_rows); the rows of the other
            # ranks keep their host values
            ja, jb = self._rows_of(GR)(n)
            fnx, _, nk = h.shape
            st = self._stage(fnx * (jb - ja + 1) * nk)
            _lib.check(_lib.lib().dc_export_rows(GR.dyncore(), self.table[n][0], st.data_ptr(),
                                                 st.numel() * 8, ja, jb, self._stream()))
            h[:, ja:jb + 1, :] = st.view(fnx, jb - ja + 1, nk).cpu().numpy()
            return
        dst = torch.from_numpy(h).view(-1)
        st = self._stage(dst.numel())
        _lib.check(_lib.lib().dc_export_field(GR.dyncore(), self.table[n][0], st.data_ptr(),
                                              st.numel() * 8, self._stream()))
        dst.copy_(st, non_blocking=True)
        if self.torch_device.type == 'cuda':
            torch.cuda.current_stream(self.torch_device).synchronize()


    # fields the fused step does not keep up to date: the fluxes / tendencies of the reference's
    # kernel decomposition (the fused stage kernel forms them in registers)
    KERNEL_MODE_ONLY = ('UFLX', 'VFLX', 'FLXDIV', 'BFLX', 'CFLX', 'DFLX', 'EFLX', 'RFLX', 'QFLX',
                        'SFLX', 'TFLX', 'WWIND_UWIND', 'WWIND_VWIND', 'dUFLXdt', 'dVFLXdt',
                        'dPOTTdt', 'dQVdt', 'dQCdt', 'dCOLPdt')

    def _refresh_for_export(self, GR, n):
        """An export of a flux / tendency field after fused steps first evaluates the
        tendencies of the CURRENT state with the kernel decomposition (dc_compute_tendencies:
        it only writes derived fields, the trajectory is unchanged), once per state.  The
        library itself brings PVTF / PVTFVB / PHIVB up to date inside dc_export_field."""
        if (n not in self.KERNEL_MODE_ONLY or GR.band[1] > 1 or GR.i_coupling or
                getattr(GR, '_mode', 'fused') != 'fused'):
            return
        ver = getattr(GR, '_state_version', 0)
        if getattr(self, '_tend_version', None) == ver:
            return
        d = self.device
        # fields of the stepping path that the evaluation overwrites keep their values
        keep = {m: d[m].clone() for m in ('COLP_OLD', 'COLP_NEW', 'WWIND')}
        d['COLP_OLD'].copy_(d['COLP'])
        _lib.check(_lib.lib().dc_compute_tendencies(GR.dyncore(), self._stream()))
        for m, v in keep.items():
            d[m].copy_(v)
        self._tend_version = ver

    # ------------------------------------------------------------ latitude bands: band-shaped I/O
    # The reference-layout host arrays cover the whole grid on every rank.  A caller that only
    # owns its band (e2e leg of bench.py on N GPUs) moves just the rows this rank holds:
    # a contiguous (fnx, rows, nk) pinned buffer <-> device, and a strided device-side copy
    # between that and the reference-layout staging that dc_import_field / dc_export_field
    # work on (torch copies only, no extra kernels of this library).
    def field_shape(self, n):
        d = self.fdict[n]
        GR = OWNERS.get(self.device[n].data_ptr())
        return (int(GR.nx) + 2 + d['stgx'], int(GR.ny) + 2 + d['stgy'], int(d['dimz']))

    def held_rows(self, GR, n):
        """rows [ja, jb] of field n (reference layout) this rank holds: its band and the halo
        rows beside it (csrc/dc_api_impl.h: held_rows)"""
        fny = self.field_shape(n)[1]
        return max(0, int(GR.j0) - 2), min(int(GR.j1) + 3, fny - 1)

    def band_buffer(self, GR, n):
        """page-locked host tensor (fnx, held rows, nk) for to_device_band / to_host_band"""
        fnx, _, nk = self.field_shape(n)
        ja, jb = self.held_rows(GR, n)
        pin = self.torch_device.type == 'cuda'
        return torch.empty((fnx, jb - ja + 1, nk), dtype=torch.float64, pin_memory=pin)

    def _stage_band(self, nelem):
        if getattr(self, '_staging_b', None) is None or self._staging_b.numel() < nelem:
            self._staging_b = torch.empty(nelem, dtype=torch.float64, device=self.torch_device)
        return self._staging_b[:nelem]

    def to_device_band(self, GR, n, hb):
        fnx, fny, nk = self.field_shape(n)
        ja, jb = self.held_rows(GR, n)
        assert tuple(hb.shape) == (fnx, jb - ja + 1, nk), (n, tuple(hb.shape))
        sb = self._stage_band(hb.numel())
        sb.copy_(hb.view(-1), non_blocking=True)
        st = self._stage(fnx * fny * nk)
        st.view(fnx, fny, nk)[:, ja:jb + 1, :].copy_(sb.view(fnx, jb - ja + 1, nk))
        _lib.check(_lib.lib().dc_import_field(GR.dyncore(), self.table[n][0], st.data_ptr(),
                                              st.numel() * 8, self._stream()))

    def to_host_band(self, GR, n, hb):
        fnx, fny, nk = self.field_shape(n)
        ja, jb = self.held_rows(GR, n)
        assert tuple(hb.shape) == (fnx, jb - ja + 1, nk), (n, tuple(hb.shape))
        st = self._stage(fnx * fny * nk)
        _lib.check(_lib.lib().dc_export_field(GR.dyncore(), self.table[n][0], st.data_ptr(),
                                              st.numel() * 8, self._stream()))
        sb = self._stage_band(hb.numel())
        sb.view(fnx, jb - ja + 1, nk).copy_(st.view(fnx, fny, nk)[:, ja:jb + 1, :])
        hb.view(-1).copy_(sb, non_blocking=True)
        if self.torch_device.type == 'cuda':
            torch.cuda.current_stream(self.torch_device).synchronize()


class _LazyHost(dict):
    """host field dict that allocates a NaN-filled reference-layout array the first time a
    name is used (the reference allocates all 93 up front, main_fields.py:477-485; at
    0.25 deg x 64 levels that would be 0.5 GB per field of host memory never touched)"""

    def __init__(self, shapes, pin=False, zero=()):
        super().__init__()
        self._shapes = shapes
        self._pin = pin
        self._pinned = {}
        self._zero = set(zero)      # names that start as 0 instead of NaN

    def __missing__(self, n):
        if n not in self._shapes:
            raise KeyError(n)
        if self._pin:
            # page-locked host memory: H2D / D2H copies run asynchronously at full PCIe rate
            t = torch.empty(self._shapes[n], dtype=torch.float64, pin_memory=True)
            t.fill_(0. if n in self._zero else float('nan'))
            self._pinned[n] = t
            a = t.numpy()
        else:
            a = np.full(self._shapes[n], 0. if n in self._zero else np.nan, dtype=wp)
        self[n] = a
        return a

    def __setitem__(self, n, a):
        # keep the (possibly pinned) buffer: assignments of a same-shaped array copy into it
        if dict.__contains__(self, n) and dict.__getitem__(self, n) is not a:
            cur = dict.__getitem__(self, n)
            if getattr(a, 'shape', None) == cur.shape:
                cur[...] = a
                return
        super().__setitem__(n, a)

    def __contains__(self, n):
        return n in self._shapes

    def keys(self):
        return self._shapes.keys()


def allocate_fields(GR, rows_of=None):
    """host arrays in the reference layout, NaN-filled (main_fields.py:218-487); with
    `rows_of` (name -> (ja, jb)) only that row window of every field"""
    fdict, shapes = {}, {}
    nzmap = {'nz': int(GR.nz), 'nzs': int(GR.nzs), 1: 1}
    table = {}
    for n, (fid, sx, sy, nkk) in _lib.field_table().items():
        table[n] = (sx, sy, {_lib.DC_NK_2D: 1, _lib.DC_NK_NZ: 'nz', _lib.DC_NK_NZS: 'nzs'}[nkk])
    table.update(_HOST_ONLY)
    for n, (sx, sy, dz) in table.items():
        fdict[n] = {'stgx': sx, 'stgy': sy, 'dimz': nzmap[dz], 'dtype': wp}
        fny = int(GR.ny) + 2 + sy
        if rows_of is not None:
            ja, jb = rows_of(n)
            fny = jb - ja + 1
        shapes[n] = (int(GR.nx) + 2 + sx, fny, nzmap[dz])
    # the physics coupling inputs are zero until a physics module (or the caller) fills them
    # (the reference zeroes them when its modules are off)
    return _LazyHost(shapes, pin=_lib.is_cuda() and torch.cuda.is_available(),
                     zero=COUPLING_FIELDS), fdict
