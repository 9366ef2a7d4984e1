"""Derived namelist settings and guards (reference: io_read_namelist.py:27-92)."""
import numpy as np

from . import namelist as nl

if nl.nb > 1:
    raise NotImplementedError('nb > 1 not implemented.')
if nl.lon0_deg != 0 or nl.lon1_deg != 360:
    raise NotImplementedError('In x direction only periodic boundaries implemented.')
if nl.i_time_stepping != 'MATSUNO':
    raise NotImplementedError('only the Matsuno scheme is implemented')
if nl.working_precision != 'float64':
    raise NotImplementedError('the B200 dyn core computes in float64 '
                              '(float32 is a later step, SURVEY.md 8f-4)')
if nl.COLP_dif_coef > 0:
    raise NotImplementedError('no pressure diffusion implemented')

wp_int = np.int32
wp_str = 'float64'
wp = np.float64

# computation targets; B200 is the only one this package implements
CPU = 'CPU'
GPU = 'GPU'
B200 = 'B200'
gpu_enable = nl.i_comp_mode in (2, 3)

pair_top = wp(nl.pair_top)
POTT_dif_coef = wp(nl.POTT_dif_coef)
moist_dif_coef = wp(nl.moist_dif_coef)
