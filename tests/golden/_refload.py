"""Dev-time loader for the *real* reference arithmetic (needs /root/reference; never used at test time).

xarray and dask are not installed in this image, so `import xsarsea` fails.  The numeric kernels do not
need them: we put two inert stub modules on sys.modules, pre-seed an empty `xsarsea` package whose __path__
points at the reference source, and import `xsarsea.windspeed`.  The inversion kernel is a closure nested in
`invert_from_model` (src/xsarsea/windspeed/windspeed.py:183-282); it is pulled out with `ast` and compiled
with the identical guvectorize call of :306-323.  Nothing from the reference is copied into this repository.
"""
import ast
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"


def _stub_modules():
    xr = types.ModuleType("xarray")

    class DataArray:  # only isinstance checks touch it
        pass

    xr.DataArray = DataArray
    for n in ("zeros_like", "open_dataset", "where", "merge"):
        setattr(xr, n, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError(n)))
    dask = types.ModuleType("dask")
    da = types.ModuleType("dask.array")

    class Array:
        pass

    da.Array = Array
    da.broadcast_arrays = np.broadcast_arrays
    dask.array = da
    sys.modules.setdefault("xarray", xr)
    sys.modules.setdefault("dask", dask)
    sys.modules.setdefault("dask.array", da)


def load_reference():
    """Return the reference's `xsarsea.windspeed` module (GMFs registered, numba kernels real)."""
    _stub_modules()
    if "xsarsea" not in sys.modules:
        pkg = types.ModuleType("xsarsea")
        pkg.__path__ = [REF_SRC + "/xsarsea"]
        sys.modules["xsarsea"] = pkg
    import xsarsea.windspeed as ws  # noqa

    return ws


def reference_gmf(name, ftype="numba_vectorize"):
    ws = load_reference()
    from xsarsea.windspeed.models import Model

    return Model._available_models[name]._gmf_function(ftype)


def reference_inversion_kernel(closure, parallel=True):
    """Compile the reference's __invert_from_model_1d with `closure` supplying its free variables.

    closure keys: np_sigma0_co_lut_db [wspd,phi,inc], np_wspd_dim, np_phi_dim, np_inc_dim, phi_180,
    np_sigma0_cr_lut_db [wspd,inc], np_wspd_lut_cr, np_inc_cr_dim, dsig_co  (the rest is derived here exactly
    as windspeed.py:139-168 derives it).
    """
    from numba import complex128, float64, guvectorize, void

    src = open(REF_SRC + "/xsarsea/windspeed/windspeed.py").read()
    tree = ast.parse(src)
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "__invert_from_model_1d":
            fn = node
    assert fn is not None
    mod = ast.Module(body=[fn], type_ignores=[])
    g = dict(closure)
    g["np"] = np
    g.setdefault("d_antenna", 2)
    g.setdefault("d_azi", 2)
    g.setdefault("dwspd_fg", 2)
    np_phi_lut, np_wspd_lut = np.meshgrid(g["np_phi_dim"], g["np_wspd_dim"])
    g["np_phi_lut"] = np_phi_lut
    g["np_wspd_lut"] = np_wspd_lut
    g["np_wspd_lut_co_antenna"] = np_wspd_lut * np.cos(np.radians(np_phi_lut))
    g["np_wspd_lut_co_azi"] = np_wspd_lut * np.sin(np.radians(np_phi_lut))
    exec(compile(mod, "<reference windspeed.py:183-282>", "exec"), g)
    pyfunc = g["__invert_from_model_1d"]
    vect = guvectorize(
        [void(float64[:], float64[:], float64[:], float64[:], complex128[:], complex128[:], complex128[:])],
        "(n),(n),(n),(n),(n)->(n),(n)",
        fastmath={"nnan": False},
        target="parallel" if parallel else "cpu",
    )(pyfunc)
    return vect, pyfunc
