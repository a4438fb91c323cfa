"""Wind retrieval from sigma0 and geophysical models -- B200 implementation of the `xsarsea.windspeed` namespace.

The public names are those of the reference package for this path (`register_pickle_luts` excepted: the legacy
sarwing pickle format is out of scope):

    inversion        invert_from_model (+ invert_to_speed_dir: the same inversion with the callers' abs / angle /
                     dir_sample_to_meteo post-processing fused into the kernels)
    model registry   Model, GmfModel, available_models, get_model, register_luts, register_nc_luts, register_cmod7
    dsig helpers     get_dsig, get_dsig_wspd, nesz_flattening
    submodules       gmfs (GmfModel and its decorator), gmfs_impl (the 13 built-in GMFs, evaluated on the device)
"""
from . import gmfs, gmfs_impl, models, utils
from . import cmod7 as _cmod7
from . import windspeed as _inversion

Model = models.Model
GmfModel = gmfs.GmfModel
available_models, get_model = models.available_models, models.get_model
register_luts, register_nc_luts = models.register_luts, models.register_nc_luts
register_cmod7 = _cmod7.register_cmod7
get_dsig, get_dsig_wspd, nesz_flattening = utils.get_dsig, utils.get_dsig_wspd, utils.nesz_flattening
invert_from_model = _inversion.invert_from_model
invert_to_speed_dir = _inversion.invert_to_speed_dir

__all__ = sorted([
    "invert_from_model", "invert_to_speed_dir", "available_models", "get_model", "register_cmod7", "register_nc_luts", "register_luts",
    "nesz_flattening", "GmfModel", "Model", "gmfs", "gmfs_impl", "get_dsig", "get_dsig_wspd",
])
