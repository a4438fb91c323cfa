"""Model registry and LUT handling: the B200 counterpart of xsarsea/windspeed/models.py.

Same public surface as the reference (`Model`, `LutModel`, `NcLutModel`, `available_models`, `get_model`,
`register_nc_luts`, `register_luts`; reference lines cited per function); the LUT arithmetic -- GMF outer
product, low->high linear interpolation, dB/linear conversion -- runs on the device through the C ABI
(`xs_lut_build`, `xs_lut_interp_axis`, `xs_lut_to_db`, `xs_lut_to_linear`).  Device LUTs are cached per
(model, kwargs, units), which the reference does not do (it rebuilds the LUT on every call and dask block).
"""
from __future__ import annotations

import glob
import logging
import os

import numpy as np

from .. import _device as dev
from .. import _xr

logger = logging.getLogger("xsarsea.windspeed.models")


def _grid(r, step):
    """np.linspace grid of the reference (gmfs.py:385-390, models.py:154-160): bit-identical node values."""
    if r is None:
        return None
    return np.linspace(r[0], r[1], num=int(np.round((r[1] - r[0]) / step) + 1))


class DeviceLut:
    """A LUT resident in HBM: `data` is a torch CUDA float64 tensor [inc, wspd(, phi)], grids are host arrays."""

    def __init__(self, data, inc, wspd, phi, units, resolution):
        self.data, self.inc, self.wspd, self.phi = data, inc, wspd, phi
        self.units, self.resolution = units, resolution

    @property
    def dims(self):
        return ("incidence", "wspd") + (("phi",) if self.phi is not None else ())

    def to_host(self, name=None, attrs=None):
        coords = {"incidence": self.inc, "wspd": self.wspd}
        if self.phi is not None:
            coords["phi"] = self.phi
        a = dict(units=self.units, resolution=self.resolution)
        a.update(attrs or {})
        return _xr.make_dataarray(self.data.cpu().numpy(), self.dims, coords, a, name)


class Model:
    """Abstract model (GMF or LUT), reference models.py:15-303.  Not instantiated by users; registered models
    are listed by :func:`available_models`."""

    _available_models = {}
    _name_prefix = ""
    _priority = None

    def __init__(self, name, **kwargs):
        # attribute semantics of models.py:28-50
        self.name = name
        self.pol = kwargs.pop("pol", None)
        self.units = kwargs.pop("units", None)
        self.phi_range = kwargs.pop("phi_range", None)
        self.wspd_range = kwargs.pop("wspd_range", None)
        self.__dict__.update(kwargs)
        self.resolution = kwargs.pop("resolution", None)
        if not hasattr(self, "inc_range"):
            self.inc_range = [16.0, 66.0]
        self.inc_step_lr = kwargs.pop("inc_step_lr", 1.0)
        self.wspd_step_lr = kwargs.pop("wspd_step_lr", 0.2)
        self.phi_step_lr = kwargs.pop("phi_step_lr", 2.5)
        self.inc_step = kwargs.pop("inc_step", 0.1)
        self.wspd_step = kwargs.pop("wspd_step", 0.1)
        self.phi_step = kwargs.pop("phi_step", 1.0)
        self._lut_cache = {}
        self.__class__._available_models[name] = self
        logger.debug("register model %s pol=%s units=%s", name, self.pol, self.units)

    @property
    def short_name(self):
        pre = self.__class__._name_prefix
        if pre and self.name.startswith(pre):
            return self.name.replace(pre, "", 1)
        return None

    @property
    def iscopol(self):
        """True if model is copol (models.py:176-179)"""
        return len(set(self.pol)) == 1

    @property
    def iscrosspol(self):
        """True if model is crosspol (models.py:181-184)"""
        return len(set(self.pol)) == 2

    # -- LUT pipeline on the device -------------------------------------------------------------------------
    def _raw_lut_device(self, **kwargs) -> DeviceLut:
        raise NotImplementedError(self.__class__)

    def _normalize_lut_device(self, lut: DeviceLut, **kwargs) -> DeviceLut:
        """models.py:82-174: decide whether to interpolate, build the target grids, interpolate per dimension
        (incidence, wspd, phi -- the order xarray.interp applies scipy interp1d in)."""
        if lut.units not in ("linear", "dB"):
            raise ValueError(f"Unknown lut units '{lut.units}'. Allowed are '['linear', 'dB']'")
        resolution = kwargs.pop("resolution", "high") or "high"
        if resolution not in ("high", "low"):
            raise ValueError(f"unknown resolution {resolution!r}")
        sfx = "" if resolution == "high" else "_lr"          # the step attributes that apply: *_step or *_step_lr
        axes = ("inc", "wspd") + (("phi",) if self.iscopol else ())
        names = [f"{a}_step{sfx}" for a in axes]
        # a LUT that already has the requested resolution is re-interpolated only if a step is overridden (:119-135)
        same_res = resolution == lut.resolution
        overridden = any(getattr(self, n) != kwargs.get(n, getattr(self, n)) for n in names)
        if not same_res or overridden:
            inc_step, wspd_step, phi_step = (kwargs.pop(f"{a}_step{sfx}", getattr(self, f"{a}_step{sfx}"))
                                             for a in ("inc", "wspd", "phi"))
            inc, wspd = _grid(self.inc_range, inc_step), _grid(self.wspd_range, wspd_step)
            phi = _grid(self.phi_range, phi_step) if lut.phi is not None else None
            data = lut.data
            data = dev.lut_interp_axis(data, 0, lut.inc, inc)
            data = dev.lut_interp_axis(data, 1, lut.wspd, wspd)
            if phi is not None:
                data = dev.lut_interp_axis(data, 2, lut.phi, phi)
            lut = DeviceLut(data, inc, wspd, phi, lut.units, resolution)
        return lut

    def to_lut_device(self, units="linear", **kwargs) -> DeviceLut:
        """`to_lut` that leaves the result in HBM (what `invert_from_model` consumes); cached."""
        key = (units, tuple(sorted((k, repr(v)) for k, v in kwargs.items())), self._state_key())
        hit = self._lut_cache.get(key)
        if hit is not None:
            self._apply_step_side_effects(**dict(kwargs))
            return hit
        lut = self._raw_lut_device(**dict(kwargs))       # kwargs are consumed twice, models.py:201-203
        lut = self._normalize_lut_device(lut, **dict(kwargs))
        if units == "dB":
            if lut.units == "linear":
                lut = DeviceLut(dev.lut_to_db(lut.data), lut.inc, lut.wspd, lut.phi, "dB", lut.resolution)  # :215
        elif units == "linear":
            if lut.units == "dB":
                lut = DeviceLut(dev.lut_to_linear(lut.data), lut.inc, lut.wspd, lut.phi, "linear", lut.resolution)  # :221
        elif units is not None:
            raise ValueError(f"Unit not known: {units}. Known are 'dB' or 'linear' ")
        self._lut_cache[key] = lut
        return lut

    def _state_key(self):
        """Everything on the model that the LUT depends on besides the call's kwargs."""
        return tuple(repr(getattr(self, a, None)) for a in (
            "inc_range", "wspd_range", "phi_range", "inc_step", "wspd_step", "phi_step", "inc_step_lr", "wspd_step_lr",
            "phi_step_lr"))

    def _apply_step_side_effects(self, **kwargs):
        """`_raw_lut` of GmfModel stores the steps it was called with on the model (gmfs.py:367-379)."""

    def to_lut(self, units="linear", **kwargs):
        """Get the model lut (models.py:186-230).  units: 'linear' | 'dB' | None.  Returns a DataArray
        (`sigma0_model`, dims incidence, wspd[, phi])."""
        lut = self.to_lut_device(units=units, **kwargs)
        return lut.to_host(name="sigma0_model", attrs=dict(model=self.name, pol=self.pol))

    def to_netcdf(self, file):
        """Save the model as a LUT file in the reference's schema (models.py:232-262).  NetCDF-3 through
        scipy.io.netcdf_file (netCDF4/h5 libraries are optional and absent from the build image)."""
        from scipy.io import netcdf_file

        resolution = "low" if self.iscopol else "high"
        lut = self.to_lut_device(units="dB", resolution=resolution)
        with netcdf_file(file, "w") as nc:
            nc.units, nc.pol, nc.resolution, nc.model = "dB", self.pol, resolution, str(self.short_name)
            nc.inc_range = np.asarray(self.inc_range, dtype=np.float64)
            nc.wspd_range = np.asarray(self.wspd_range, dtype=np.float64)
            nc.wspd_step = float(np.round(np.unique(np.diff(lut.wspd)), decimals=2)[0])
            nc.inc_step = float(np.round(np.unique(np.diff(lut.inc)), decimals=2)[0])
            dims = [("incidence", lut.inc), ("wspd", lut.wspd)]
            if lut.phi is not None:
                nc.phi_range = np.asarray(self.phi_range, dtype=np.float64)
                nc.phi_step = float(np.round(np.unique(np.diff(lut.phi)), decimals=2)[0])
                dims.append(("phi", lut.phi))
            for n, g in dims:
                nc.createDimension(n, g.size)
                v = nc.createVariable(n, "d", (n,))
                v[:] = g
            v = nc.createVariable("sigma0_model", "d", tuple(n for n, _ in dims))
            v[:] = lut.data.cpu().numpy()

    def __call__(self, inc, wspd, phi=None, broadcast=False):
        raise NotImplementedError(self.__class__)

    def __repr__(self):
        return f"<{self.__class__.__name__}('{self.name}') pol={self.pol}>"


class LutModel(Model):
    """Abstract class for LUT-backed models (models.py:306-347)."""

    _name_prefix = "nc_lut_"
    _priority = None

    def _raw_lut_host(self, **kwargs):
        """-> (values [inc, wspd(, phi)] float64, inc, wspd, phi|None, units, resolution)"""
        raise NotImplementedError

    def _raw_lut_device(self, **kwargs) -> DeviceLut:
        vals, inc, wspd, phi, units, resolution = self._raw_lut_host(**kwargs)
        return DeviceLut(dev.to_device(np.asarray(vals, dtype=np.float64)), np.asarray(inc, dtype=np.float64),
                         np.asarray(wspd, dtype=np.float64), None if phi is None else np.asarray(phi, dtype=np.float64),
                         units, resolution)

    def __call__(self, inc, wspd, phi=None, units=None, **kwargs):
        """Interpolate the LUT at scalar or 1-D coordinates (models.py:318-347); anything else raises
        NotImplementedError like the reference."""
        args = [v for v in (inc, wspd, phi) if v is not None]
        all_scalar = all(np.isscalar(v) for v in args)
        try:
            all_1d = all(v.ndim == 1 for v in args)
        except AttributeError:
            all_1d = False
        if not (all_scalar or all_1d):
            raise NotImplementedError("Only scalar or 1D array are implemented for LutModel")
        lut = self.to_lut_device(units=units, **kwargs)
        if (lut.phi is not None) != (phi is not None):
            raise ValueError("phi must be given for a co-pol LUT and omitted for a cross-pol one")
        tgt = [np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in args]
        data = lut.data
        for ax, (src, dst) in enumerate(zip((lut.inc, lut.wspd, lut.phi), tgt)):
            data = _interp_nan_outside(data, ax, src, dst)
        out = data.cpu().numpy()
        if all_scalar:
            return out.item()
        names = ["incidence", "wspd", "phi"][:len(tgt)]
        return _xr.make_dataarray(out, names, dict(zip(names, tgt)), dict(model=self.name, units=self.units),
                                  "sigma0_gmf")


def _interp_nan_outside(data, axis, src, dst):
    """xarray's default interp (no bounds_error): NaN outside the source grid."""
    inside = (dst >= src[0]) & (dst <= src[-1])
    res = dev.lut_interp_axis(data, axis, src, np.clip(dst, src[0], src[-1]))
    if not inside.all():
        import torch

        idx = torch.as_tensor(np.flatnonzero(~inside), device=res.device)
        res.index_fill_(axis, idx, float("nan"))
    return res


def _read_nc(path):
    """Read a LUT file written in the reference's schema -> (global attrs, {name: array})."""
    if _xr.HAVE_XARRAY:  # pragma: no cover - NetCDF-4 files need xarray + a backend
        try:
            import xarray as xr

            with xr.open_dataset(path) as ds:
                attrs = dict(ds.attrs)
                arrays = {k: np.asarray(ds[k]) for k in list(ds.variables)}
            return attrs, arrays
        except Exception:
            pass
    from scipy.io import netcdf_file

    with netcdf_file(path, "r", mmap=False) as nc:
        attrs = {}
        for k, v in nc._attributes.items():
            attrs[k] = v.decode() if isinstance(v, bytes) else (v.copy() if isinstance(v, np.ndarray) else v)
        arrays = {k: np.array(v[:]) for k, v in nc.variables.items()}
    return attrs, arrays


class NcLutModel(LutModel):
    """LUT stored in a netcdf file in xsarsea format (models.py:350-410)."""

    _priority = 10

    @property
    def short_name(self):
        return self._short_name

    def __init__(self, path, **kwargs):
        name = os.path.splitext(os.path.basename(path))[0]
        attrs, _ = _read_nc(path)
        for attr in ["units", "pol", "model", "resolution", "inc_range", "wspd_range", "phi_range", "inc_step",
                     "wspd_step", "phi_step"]:
            if attr in attrs:
                v = attrs[attr]
                if isinstance(v, np.ndarray):
                    v = [x.item() for x in v.reshape(-1)]
                    if not attr.endswith("_range"):
                        v = v[0]
                elif isinstance(v, np.generic):
                    v = v.item()
                kwargs[attr] = v
        self._short_name = kwargs.pop("model")
        if kwargs["resolution"] == "low":  # models.py:392-395
            kwargs["inc_step_lr"] = kwargs.pop("inc_step")
            kwargs["wspd_step_lr"] = kwargs.pop("wspd_step")
            kwargs["phi_step_lr"] = kwargs.pop("phi_step", None)
        super().__init__(name, **kwargs)
        self.path = path

    def _raw_lut_host(self, **kwargs):
        if not os.path.isfile(self.path):
            raise FileNotFoundError(self.path)
        attrs, arrays = _read_nc(self.path)
        vals = np.asarray(arrays["sigma0_model"], dtype=np.float64)
        phi = np.asarray(arrays["phi"], dtype=np.float64) if vals.ndim == 3 else None
        return vals, arrays["incidence"], arrays["wspd"], phi, attrs["units"], attrs["resolution"]


def register_nc_luts(topdir, gmf_names=None):
    """Register all netcdf luts `nc_lut_*.nc` found under `topdir` (models.py:413-450)."""
    for path in glob.glob(os.path.join(topdir, f"{NcLutModel._name_prefix}*.nc")):
        path = os.path.abspath(os.path.join(topdir, path))
        name = os.path.basename(path).replace(".nc", "")
        if gmf_names is None or name in gmf_names:
            NcLutModel(path)


def available_models(pol=None):
    """pandas.DataFrame of registered models, indexed by name, columns alias/pol/model (models.py:453-498):
    among models sharing a short name the one with the lowest `_priority` owns the alias."""
    import pandas as pd

    rows = [(n, m.short_name, m._priority, m.pol, m) for n, m in Model._available_models.items()]
    df = pd.DataFrame(rows, columns=["name", "short_name", "priority", "pol", "model"]).set_index("name")
    df.index.name = None
    if len(df):
        order = df.sort_values("priority", ascending=True, kind="stable")
        aliased = order.drop_duplicates("short_name")
    else:
        aliased = df
    non_aliased = df.drop(aliased.index).copy()
    non_aliased["short_name"] = None
    out = pd.concat([aliased, non_aliased]).rename(columns=dict(short_name="alias")).drop(columns="priority")
    if pol is not None:
        out = out[out.pol == pol]
    return out


def get_model(name):
    """Model by name or alias; a Model instance is returned as is (models.py:510-538)."""
    if isinstance(name, Model):
        return name
    if name in Model._available_models:
        return Model._available_models[name]
    avail = available_models()
    hits = avail[avail.alias == name]
    if len(hits) != 1:
        raise KeyError(f"model {name} not found")
    return hits.model.iloc[0]


def register_luts(topdir=None, topdir_cmod7=None):
    """Register gmf models, and optionally nc luts and cmod7 (models.py:541-568)."""
    from . import gmfs
    from .cmod7 import register_cmod7

    gmfs.GmfModel.activate_gmfs_impl()
    if topdir is not None:
        register_nc_luts(topdir)
    if topdir_cmod7 is not None:
        register_cmod7(topdir_cmod7)
