"""xsarsea_b200: the wind-inversion hot path of umr-lops/xsarsea on B200 (sm_100a).

Same Python API as the reference for this path (xsarsea.windspeed.*, xsarsea.sigma0_detrend); the work is
done by hand-written CUDA kernels reached through a C-ABI shared library (include/xsarsea_b200.h).
"""
__version__ = "0.1.0"

from . import gradients, windspeed  # noqa: E402,F401
from .detrend import (  # noqa: E402,F401
    dir_meteo_to_oceano,
    dir_meteo_to_sample,
    dir_oceano_to_meteo,
    dir_sample_to_meteo,
    dir_to_180,
    dir_to_360,
    sigma0_detrend,
)

# names of xsarsea/__init__.py:1-11 that belong to the hot path or are one-line helpers around it; get_test_file (HTTP
# download of test data) and read_sarwing_owi (NetCDF-4 reader) are out of scope
__all__ = ["windspeed", "gradients", "sigma0_detrend", "dir_meteo_to_sample", "dir_sample_to_meteo", "dir_meteo_to_oceano",
           "dir_oceano_to_meteo", "dir_to_180", "dir_to_360", "__version__"]
