"""Turbulence module (reference: turb_main.py:38-50, turb_compute.py:53-204): the vertical
mixing coefficients KMOM / KHEAT from the bulk Richardson number and a Blackadar mixing
length, one column kernel on the device (dc_compute_turbulence, csrc/dc_kernels.h:
TurbulenceBody).  They feed the vertical turbulent transport terms of the dynamical core
(grid made with i_coupling=1).  Same class / method / argument names as the reference.
"""
from .dyn_org_discretizations import _Factory
from .io_read_namelist import B200
from .misc_utilities import function_input_fields


class Turbulence(_Factory):

    def __init__(self, GR, target=B200):
        super().__init__(target)
        if not GR.i_coupling:
            raise RuntimeError('the turbulence module needs a grid made with i_coupling=1 '
                               '(KMOM / KHEAT have no device buffers otherwise)')
        self.fields_main = function_input_fields(self.compute_turbulence)

    def compute_turbulence(self, GR, KMOM, KHEAT, PHIVB, HSURF, PHI, QV, WINDX, WINDY,
                           POTTVB, POTT):
        self._run('dc_compute_turbulence', dict(
            KMOM=KMOM, KHEAT=KHEAT, PHIVB=PHIVB, HSURF=HSURF, PHI=PHI, QV=QV, WINDX=WINDX,
            WINDY=WINDY, POTTVB=POTTVB, POTT=POTT))
