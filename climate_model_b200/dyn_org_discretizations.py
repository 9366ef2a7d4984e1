"""Target-dispatch factories (reference: dyn_org_discretizations.py:75-393) -- the drop-in
boundary of the dynamical core.  Same classes, method names and argument names as the
reference (the argument names ARE the field lists, misc_utilities.py:63-73); target B200
enqueues the hand-written sm_100a kernels of libdyncore.so on the current torch stream.

The reference passes arrays to its kernels; libdyncore works on buffers bound by name
(dc_bind_field), so each call re-binds the tensors it is given (a pointer store per field)
and then runs the matching fine-grained entry of include/dyncore.h.
"""
import torch

from . import _lib
from .io_read_namelist import B200
from .misc_utilities import function_input_fields


def _stream(t):
    if t.is_cuda:
        return torch.cuda.current_stream(t.device).cuda_stream
    return 0


class _Factory:
    def __init__(self, target):
        if target != B200:
            raise NotImplementedError(
                "target %r: only target 'B200' is implemented here (the numba CPU/GPU targets "
                "live in the reference)" % (target,))
        self.target = target
        self._table = None

    def _run(self, entry, fields):
        """bind `fields` ({name: torch tensor}) and enqueue `entry` on the current stream.
        The grid (and with it the library handle) is found through the tensors: ModelFields
        registers every device buffer it allocates (main_fields.OWNERS)."""
        from .main_fields import bind_all, fields_owner_of, owner_of
        L = _lib.lib()
        GR = owner_of(fields)
        h = GR.dyncore()
        # the whole field set the tensors come from first (library work fields such as PGCOL
        # follow it), then the tensors actually passed; a foreign tensor invalidates the key
        F = fields_owner_of(fields)
        if F is not None:
            bind_all(GR, F)
        if self._table is None:
            self._table = _lib.field_table()
        stream = 0
        for n, t in fields.items():
            if t is None or n not in self._table:
                continue       # coupling field of a grid without i_coupling: identically zero
            if not isinstance(t, torch.Tensor):
                raise TypeError('field %s: target B200 needs the device tensor (F.device), got %s'
                                % (n, type(t).__name__))
            if _lib.is_cuda() and not t.is_cuda:
                raise RuntimeError('field %s is not on a CUDA device; there is no CPU fallback' % n)
            if F is None or F.device.get(n) is not t:
                _lib.check(L.dc_bind_field(h, self._table[n][0], t.data_ptr(), t.numel() * 8))
                GR._bound_key = None
            stream = _stream(t)
        GR._state_version = getattr(GR, '_state_version', 0) + 1
        _lib.check(getattr(L, entry)(h, stream))


class TendencyFactory(_Factory):
    """dyn_org_discretizations.py:75-294"""

    def __init__(self, target):
        super().__init__(target)
        self.fields_continuity = function_input_fields(self.continuity)
        self.fields_momentum = function_input_fields(self.momentum)
        self.fields_temperature = function_input_fields(self.temperature)
        self.fields_moisture = function_input_fields(self.moisture)

    def continuity(self, GR, GRF, UFLX, VFLX, FLXDIV, UWIND, VWIND, WWIND, COLP, dCOLPdt,
                   COLP_NEW, COLP_OLD):
        self._run('dc_continuity', dict(
            UFLX=UFLX, VFLX=VFLX, FLXDIV=FLXDIV, UWIND=UWIND, VWIND=VWIND, WWIND=WWIND,
            COLP=COLP, dCOLPdt=dCOLPdt, COLP_NEW=COLP_NEW, COLP_OLD=COLP_OLD))

    def momentum(self, GRF, dUFLXdt, dVFLXdt, UWIND, VWIND, WWIND, UFLX, VFLX, CFLX, QFLX,
                 DFLX, EFLX, SFLX, TFLX, BFLX, RFLX, PHI, PHIVB, COLP, COLP_NEW, POTT, PVTF,
                 PVTFVB, WWIND_UWIND, WWIND_VWIND, KMOM_dUWINDdz=None, KMOM_dVWINDdz=None,
                 KMOM=None, RHOVB=None, RHO=None, dUFLXdt_TURB=None, dVFLXdt_TURB=None,
                 SMOMXFLX=None, SMOMYFLX=None):
        self._run('dc_momentum', dict(
            dUFLXdt=dUFLXdt, dVFLXdt=dVFLXdt, UWIND=UWIND, VWIND=VWIND, WWIND=WWIND, UFLX=UFLX,
            VFLX=VFLX, CFLX=CFLX, QFLX=QFLX, DFLX=DFLX, EFLX=EFLX, SFLX=SFLX, TFLX=TFLX,
            BFLX=BFLX, RFLX=RFLX, PHI=PHI, PHIVB=PHIVB, COLP=COLP, COLP_NEW=COLP_NEW, POTT=POTT,
            PVTF=PVTF, PVTFVB=PVTFVB, WWIND_UWIND=WWIND_UWIND, WWIND_VWIND=WWIND_VWIND,
            KMOM_dUWINDdz=KMOM_dUWINDdz, KMOM_dVWINDdz=KMOM_dVWINDdz, KMOM=KMOM, RHOVB=RHOVB,
            RHO=RHO, dUFLXdt_TURB=dUFLXdt_TURB, dVFLXdt_TURB=dVFLXdt_TURB, SMOMXFLX=SMOMXFLX,
            SMOMYFLX=SMOMYFLX))

    def temperature(self, GRF, dPOTTdt, POTT, UFLX, VFLX, COLP, POTTVB, WWIND, COLP_NEW,
                    PHI=None, PHIVB=None, KHEAT=None, RHO=None, RHOVB=None, SSHFLX=None,
                    dPOTTdt_TURB=None, dPOTTdt_RAD=None):
        self._run('dc_temperature', dict(
            dPOTTdt=dPOTTdt, POTT=POTT, UFLX=UFLX, VFLX=VFLX, COLP=COLP, POTTVB=POTTVB,
            WWIND=WWIND, COLP_NEW=COLP_NEW, PHI=PHI, PHIVB=PHIVB, KHEAT=KHEAT, RHO=RHO,
            RHOVB=RHOVB, SSHFLX=SSHFLX, dPOTTdt_TURB=dPOTTdt_TURB, dPOTTdt_RAD=dPOTTdt_RAD))

    def moisture(self, GRF, dQVdt, QV, dQCdt, QC, UFLX, VFLX, COLP, WWIND, COLP_NEW,
                 dQVdt_TURB=None, PHI=None, PHIVB=None, KHEAT=None, RHO=None, RHOVB=None,
                 SLHFLX=None):
        self._run('dc_moisture', dict(
            dQVdt=dQVdt, QV=QV, dQCdt=dQCdt, QC=QC, UFLX=UFLX, VFLX=VFLX, COLP=COLP,
            WWIND=WWIND, COLP_NEW=COLP_NEW, dQVdt_TURB=dQVdt_TURB, PHI=PHI, PHIVB=PHIVB,
            KHEAT=KHEAT, RHO=RHO, RHOVB=RHOVB, SLHFLX=SLHFLX))


class DiagnosticsFactory(_Factory):
    """dyn_org_discretizations.py:299-346"""

    def __init__(self, target):
        super().__init__(target)
        self.fields_primary_diag = function_input_fields(self.primary_diag)
        self.fields_secondary_diag = function_input_fields(self.secondary_diag)

    def primary_diag(self, GRF, COLP, PVTF, PVTFVB, PHI, PHIVB, POTT, POTTVB, HSURF):
        self._run('dc_primary_diag', dict(
            COLP=COLP, PVTF=PVTF, PVTFVB=PVTFVB, PHI=PHI, PHIVB=PHIVB, POTT=POTT, POTTVB=POTTVB,
            HSURF=HSURF))

    def secondary_diag(self, POTTVB, TAIRVB, PVTFVB, COLP, PAIR, PAIRVB, PHI, POTT, TAIR,
                       RHO, RHOVB, PVTF, UWIND, VWIND, WINDX, WINDY, WIND):
        self._run('dc_secondary_diag', dict(
            POTTVB=POTTVB, TAIRVB=TAIRVB, PVTFVB=PVTFVB, PAIR=PAIR, PAIRVB=PAIRVB, POTT=POTT,
            TAIR=TAIR, RHO=RHO, RHOVB=RHOVB, PVTF=PVTF, UWIND=UWIND, VWIND=VWIND, WINDX=WINDX,
            WINDY=WINDY, WIND=WIND))


class PrognosticsFactory(_Factory):
    """dyn_org_discretizations.py:349-393"""

    def __init__(self, target):
        super().__init__(target)
        self.fields_prognostic = function_input_fields(self.euler_forward)

    def euler_forward(self, GR, GRF, UWIND_OLD, UWIND, VWIND_OLD, VWIND, COLP_OLD, COLP,
                      POTT_OLD, POTT, QV, QV_OLD, QC, QC_OLD, dUFLXdt, dVFLXdt, dPOTTdt, dQVdt,
                      dQCdt):
        self._run('dc_euler_forward', dict(
            UWIND_OLD=UWIND_OLD, UWIND=UWIND, VWIND_OLD=VWIND_OLD, VWIND=VWIND,
            COLP_OLD=COLP_OLD, COLP=COLP, POTT_OLD=POTT_OLD, POTT=POTT, QV=QV, QV_OLD=QV_OLD,
            QC=QC, QC_OLD=QC_OLD, dUFLXdt=dUFLXdt, dVFLXdt=dVFLXdt, dPOTTdt=dPOTTdt,
            dQVdt=dQVdt, dQCdt=dQCdt))
