"""B200-native dynamical core of the Potopoles/Climate_Model lat-lon sigma-coordinate model.

Module names mirror the reference (namelist, main_grid, main_fields,
dyn_org_discretizations, dyn_tendencies, dyn_matsuno, solver); the numerical work runs in
hand-written sm_100a CUDA kernels (csrc/) behind the C ABI of include/dyncore.h, loaded
through ctypes (_lib.py).  PyTorch is used only to own device buffers and streams.
"""
__version__ = '0.1.0'
