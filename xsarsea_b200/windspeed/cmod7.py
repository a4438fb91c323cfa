"""CMOD7 as a file-backed LUT (reference xsarsea/windspeed/cmod7.py:10-106).

The KNMI table `gmf_cmod7_vv.dat_little_endian` (float32, Fortran record markers, 250 x 73 x 51 =
wspd x phi x incidence, linear units, low resolution) is read on the host, transposed to (incidence, wspd, phi)
and uploaded; interpolation to the inversion grid and the dB conversion happen on the device like for any model.
"""
import os

import numpy as np

from .models import LutModel

_M, _N, _P = 250, 73, 51  # wspd 0.2..50 step 0.2; phi 0..180 step 2.5; inc 16..66 step 1 (cmod7.py:32-40)


class Cmod7Model(LutModel):
    _name_prefix = "gmf_"
    _priority = 1

    def __init__(self, name, path, **kwargs):
        super().__init__(name, **kwargs)
        self.path = path

    def _raw_lut_host(self, **kwargs):
        if not os.path.isdir(self.path):
            raise FileNotFoundError(self.path)
        sigma0_path = os.path.join(self.path, "gmf_cmod7_vv.dat_little_endian")
        if not os.path.isfile(sigma0_path):
            raise FileNotFoundError(sigma0_path)
        sigma0 = np.fromfile(sigma0_path, dtype="<f4")[1:-1]          # drop the record markers
        sigma0 = sigma0.reshape((_M, _N, _P), order="F")               # (wspd, phi, inc)
        self.wspd_step_lr, self.inc_step_lr, self.phi_step_lr = 0.2, 1, 2.5
        self.wspd_range, self.inc_range, self.phi_range = [0.2, 50.0], [16, 66], [0, 180]
        wspd = np.arange(0.2, 50.0 + 0.2, 0.2)
        inc = np.arange(16, 66 + 1, 1)
        phi = np.arange(0, 180 + 2.5, 2.5)
        vals = np.ascontiguousarray(np.transpose(sigma0, (2, 0, 1)), dtype=np.float64)
        return vals, inc, wspd, phi, "linear", "low"


def register_cmod7(topdir):
    """Register cmod7 from the directory holding the KNMI table (cmod7.py:78-106)."""
    Cmod7Model(Cmod7Model._name_prefix + "cmod7", topdir, pol="VV")
