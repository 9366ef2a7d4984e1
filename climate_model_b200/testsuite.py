"""Compare the NetCDF output of two runs (reference: testsuite.py:15-80).

    python -m climate_model_b200.testsuite [--ref ../output_ref] [--test ../output_test]
                                           [--file out0002.nc] [--tolerance 1e-4]

Same metric, tolerance, field list and messages as the reference: per field
max|test - ref| / max|test|, failure above the tolerance, 'Bitwise identical' when the summed
deviation is 0.  Files are read with scipy.io.netcdf_file (xarray is not available here).
"""
import argparse
import os

import numpy as np
from scipy.io import netcdf_file

tolerance = 1E-4
test_fields = ['UWIND', 'VWIND', 'WWIND', 'COLP', 'PHI',
               'SURFTEMP', 'SLHFLX', 'SMOMXFLX', 'SMOMYFLX', 'SURFALBEDSW', 'SSHFLX',
               'SWFLXNET', 'LWFLXNET', 'QV', 'QC', 'dVFLXdt_TURB']


def compare_outputs(ref_file, test_file, fields=test_fields, tolerance=tolerance, verbose=True):
    """returns (failed, deviation_sum, {field: deviation})"""
    say = print if verbose else (lambda *a: None)
    failed, deviation_sum, devs = False, 0., {}
    with netcdf_file(ref_file, 'r', mmap=False) as ds_ref, \
            netcdf_file(test_file, 'r', mmap=False) as ds_test:
        for test_field in fields:
            say(test_field)
            if test_field not in ds_test.variables or test_field not in ds_ref.variables:
                say('Not in output file.')
                continue
            t = np.asarray(ds_test.variables[test_field][:], dtype=np.float64)
            r = np.asarray(ds_ref.variables[test_field][:], dtype=np.float64)
            maxv_test = np.abs(t).max()
            if maxv_test <= tolerance:
                say('all elements == 0')
                deviation = 0.
            else:
                deviation = float(np.abs(t - r).max() / maxv_test)
            devs[test_field] = deviation
            deviation_sum += deviation
            say(deviation)
            if not deviation <= tolerance:          # NaN fails too
                say('!!! Deviation !!!')
                say('max val is ' + str(maxv_test))
                failed = True
            say()
    if failed:
        say('Test failed!')
    else:
        say('Equal with tolerance ' + str(tolerance) + '    GREAT!')
        say('Deviation summed over variables is ' + str(deviation_sum))
        if deviation_sum == 0:
            say('Bitwise identical')
    return failed, deviation_sum, devs


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument('--ref', default='../output_ref')
    ap.add_argument('--test', default='../output_test')
    ap.add_argument('--file', default='out0002.nc')
    ap.add_argument('--tolerance', type=float, default=tolerance)
    a = ap.parse_args()
    failed, _, _ = compare_outputs(os.path.join(a.ref, a.file), os.path.join(a.test, a.file),
                                   tolerance=a.tolerance)
    raise SystemExit(1 if failed else 0)


if __name__ == '__main__':
    main()
