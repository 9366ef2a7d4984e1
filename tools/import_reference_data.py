#!/usr/bin/env python
"""Convert the reference's two INPUT DATA files into climate_model_b200/data/ic_data.npz.

  data/elev.1-deg.nc       NetCDF-3 classic, 1-degree global surface elevation
                           (lon(360) = 0.5..359.5, lat(180) = 89.5..-89.5, data(1,180,360) int16 m)
  data/mean_vert_prof.dat  66-row standard-atmosphere table
                           (columns: z [m], g [m s-2], p [Pa], T [K], rho [kg m-3])

These are data assets a user of the reference already has (io_initial_conditions.py:158,186,257
read them), not source code; they are stored here as arrays so that the initial-condition
builder works on a machine without the reference checkout (the GPU box).
Usage: python tools/import_reference_data.py [/root/reference]
"""
import os
import sys

import numpy as np
from scipy.io import netcdf_file

ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
nc = netcdf_file(os.path.join(ref, 'data', 'elev.1-deg.nc'), 'r', mmap=False)
lon = np.array(nc.variables['lon'][:], dtype=np.float64)
lat = np.array(nc.variables['lat'][:], dtype=np.float64)
elev = np.array(nc.variables['data'][0, :, :])
profile = np.loadtxt(os.path.join(ref, 'data', 'mean_vert_prof.dat'))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'climate_model_b200',
                   'data', 'ic_data.npz')
np.savez_compressed(out, elev_lon=lon, elev_lat=lat, elev=elev, profile=profile)
print('wrote', os.path.normpath(out), elev.dtype, elev.shape, profile.shape)
