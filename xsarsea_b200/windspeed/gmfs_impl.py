"""The 13 analytic GMFs of the reference (xsarsea/windspeed/gmfs_impl.py), registered as device-native models.

The formulas themselves live in CUDA (xsarsea_b200/csrc/xs_gmf.cu): CMOD5 / CMOD5.N with the Zhang-A and
Mouche HH polarisation ratios (gmfs_impl.py:8-210), CMOD-IFR2 (:213-303) and the eight VH power-law/sigmoid
models (:325-707).  Registration attributes (pol, wspd_range, phi range) follow the reference's decorators
and the phi auto-detection result of gmfs.py:134-158 (all VV/HH models are even in phi -> [0, 180]).
"""
from .._native import GMF_IDS
from .gmfs import GmfModel

_COPOL = {
    "gmf_cmod5": "VV",
    "gmf_cmod5n": "VV",
    "gmf_cmod5n_pr_zhangA": "HH",
    "gmf_cmod5n_pr_mouche1": "HH",
    "gmf_cmodifr2": "VV",
}


def _device_only(name):
    def f(inc, wspd, phi=None):
        raise NotImplementedError(f"{name} is evaluated on the device; call the model object instead")

    f.__name__ = name
    return f


def _register_builtins():
    for name, ident in GMF_IDS.items():
        if name in GmfModel._registry:
            continue
        if name in _COPOL:
            GmfModel._register_function(_device_only(name), name, [0.2, 50.0], _COPOL[name], "linear",
                                        _device_id=ident, _phi_range=[0.0, 180.0])
        else:
            GmfModel._register_function(_device_only(name), name, [3.0, 80.0], "VH", "linear", _device_id=ident,
                                        _phi_range=None)


_register_builtins()
