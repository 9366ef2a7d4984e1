"""Physical constants (same names and values as the reference's io_constants.py:16-24)."""
import numpy as np

wp = np.float64

con_g = wp(9.81)
con_rE = wp(6371000)
con_omega = wp(7.292115E-5)
con_Rd = wp(287.058)
con_cp = wp(1005)
con_kappa = wp(con_Rd / con_cp)
con_Lh = wp(2264E3)
