"""windspeed module, for retrieving wind speed from sigma0 and models (B200 implementation).

Mirrors the names exported by xsarsea/windspeed/__init__.py:5-34 (register_pickle_luts excepted: the legacy
sarwing pickle format is out of scope).
"""
__all__ = [
    "invert_from_model",
    "available_models",
    "get_model",
    "register_cmod7",
    "register_nc_luts",
    "register_luts",
    "nesz_flattening",
    "GmfModel",
    "Model",
    "gmfs",
    "gmfs_impl",
    "get_dsig",
    "get_dsig_wspd",
]

from . import gmfs, gmfs_impl  # noqa: F401
from .cmod7 import register_cmod7
from .gmfs import GmfModel
from .models import Model, available_models, get_model, register_luts, register_nc_luts
from .utils import get_dsig, get_dsig_wspd, nesz_flattening
from .windspeed import invert_from_model
