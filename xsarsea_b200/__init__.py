"""xsarsea_b200: the wind-inversion hot path of umr-lops/xsarsea on B200 (sm_100a).

Same Python API as the reference for this path (xsarsea.windspeed.*, xsarsea.sigma0_detrend); the work is
done by hand-written CUDA kernels reached through a C-ABI shared library (include/xsarsea_b200.h).
"""
__version__ = "0.1.0"

from . import windspeed  # noqa: E402,F401
from .detrend import sigma0_detrend  # noqa: E402,F401

__all__ = ["windspeed", "sigma0_detrend", "__version__"]
