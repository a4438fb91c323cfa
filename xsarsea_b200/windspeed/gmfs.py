"""GmfModel: models defined by an analytic function -- counterpart of xsarsea/windspeed/gmfs.py.

The 13 built-in GMFs (gmfs_impl.py) are evaluated by FP64 CUDA kernels (`xs_gmf_eval`, `xs_lut_build`); the
Python scalar functions the reference JIT-compiles with numba have no equivalent here.  A user-registered
Python GMF (`GmfModel.register`) cannot be compiled to CUDA, so it is evaluated on the host with numpy.vectorize
to produce its LUT *array*, which then enters the device path exactly like a file-backed LUT (SURVEY.md hard
part 6); inversion with such a model runs on the GPU like any other.
"""
from __future__ import annotations

import logging

import numpy as np

from .. import _device as dev
from .. import _native as nat
from .. import _xr
from .models import DeviceLut, Model, _grid

logger = logging.getLogger("xsarsea.windspeed")


class GmfModel(Model):
    """Model from an analytical function (gmfs.py:13-395)."""

    _name_prefix = "gmf_"
    _priority = 3
    _registry = {}
    _deferred_registrations = []

    @classmethod
    def register(cls, name=None, pol=None, units="linear", defer=True, **kwargs):
        """Decorator registering a scalar gmf function `f(inc, wspd, phi)` (gmfs.py:23-105).

        name must start with 'gmf_' (defaults to the function name); wspd_range defaults to [0.2, 50] for
        co-pol and [3, 80] for cross-pol; with defer=True the registration waits for `activate_gmfs_impl`.
        """

        def inner(func):
            gmf_name = name or func.__name__
            if not gmf_name.startswith(cls._name_prefix):
                raise ValueError(f"gmf function must start with '{cls._name_prefix}'. Got {gmf_name}")
            wspd_range = kwargs.pop("wspd_range", None)
            if wspd_range is None:
                wspd_range = [0.2, 50.0] if len(set(pol)) == 1 else [3.0, 80.0]
            if defer:
                cls._deferred_registrations.append((func, gmf_name, wspd_range, pol, units, kwargs))
            else:
                cls._register_function(func, gmf_name, wspd_range, pol, units, **kwargs)
            return func

        return inner

    @classmethod
    def _register_function(cls, func, name, wspd_range, pol, units, **kwargs):
        gmf = cls(name, func, wspd_range, pol, units, **kwargs)
        cls._registry[name] = gmf

    @classmethod
    def activate_gmfs_impl(cls, gmfs_names=None, **kwargs):
        """Process deferred registrations, optionally filtered by name (gmfs.py:112-125)."""
        for func, name, wspd_range, pol, units, reg_kwargs in cls._deferred_registrations:
            if gmfs_names is None or name in gmfs_names:
                cls._register_function(func, name, wspd_range, pol, units, **{**reg_kwargs, **kwargs})

    def __init__(self, name, gmf_pyfunc_scalar, wspd_range=[0.2, 50.0], pol=None, units=None, **kwargs):
        device_id = kwargs.pop("_device_id", None)
        if device_id is not None:
            # built-in: phi usage is a property of the formula, no probing call needed
            phi_range = kwargs.pop("_phi_range")
        else:
            # gmfs.py:134-158: probe the scalar function for phi usage and symmetry
            sigma0_gmf = [gmf_pyfunc_scalar(35.0, 0.2, 90.0)]
            try:
                gmf_pyfunc_scalar(35.0, 0.2, None)
                phi_range = None
            except TypeError:
                sigma0_gmf = [np.abs(gmf_pyfunc_scalar(35.0, 0.2, phi) - gmf_pyfunc_scalar(35.0, 0.2, -phi))
                              for phi in [0, 90, 180, 270]]
                phi_range = [0.0, 180.0] if min(sigma0_gmf) < 1e-15 else [0.0, 360.0]
            if (units == "dB" and min(sigma0_gmf) > 0) or (units == "linear" and min(sigma0_gmf) < 0):
                logger.info(f"Possible bad units '{units}'  for gmf {name}")
        super().__init__(name, units=units, pol=pol, wspd_range=wspd_range, phi_range=phi_range, **kwargs)
        self._gmf_pyfunc_scalar = gmf_pyfunc_scalar
        self._device_id = device_id

    # -- evaluation ---------------------------------------------------------------------------------------------
    def _eval_broadcast(self, inc, wspd, phi):
        """Element-wise evaluation of already-broadcast numpy arrays (reference K3, gmfs.py:210-214)."""
        if self._device_id is None:
            f = np.vectorize(self._gmf_pyfunc_scalar, otypes=[np.float64])
            return f(inc, wspd, phi) if self.phi_range is not None else f(inc, wspd)
        torch = nat.torch_cuda()
        f32 = inc.dtype == np.float32 and wspd.dtype == np.float32   # signature ffd->f, gmfs.py:211
        rdt = np.float32 if f32 else np.float64
        ti = dev.to_device(np.ascontiguousarray(inc, dtype=rdt))
        tw = dev.to_device(np.ascontiguousarray(wspd, dtype=rdt))
        tp = None
        if self.phi_range is not None:
            tp = dev.to_device(np.ascontiguousarray(phi, dtype=np.float64))
        return dev.gmf_eval(self._device_id, ti, tw, tp).cpu().numpy()

    def _eval_outer(self, inc, wspd, phi):
        """Outer product [inc, wspd(, phi)] on the device (reference K2, gmfs.py:218-230); returns a tensor."""
        if self._device_id is None:
            f = np.vectorize(self._gmf_pyfunc_scalar, otypes=[np.float64])
            if phi is None:
                vals = f(inc[:, None], wspd[None, :])
            else:
                vals = f(inc[:, None, None], wspd[None, :, None], phi[None, None, :])
            return dev.to_device(np.ascontiguousarray(vals, dtype=np.float64))
        return dev.lut_build(self._device_id, inc, wspd, phi)

    def __call__(self, inc, wspd, phi=None, broadcast=False, numba=True):
        """sigma0 = gmf(inc, wspd[, phi]) with the shape rules of gmfs.py:266-348: all scalars -> float; all 1-D ->
        outer-product DataArray (dims incidence, wspd[, phi] or the inputs' own dim names); any N-D input or
        broadcast=True -> element-wise on the broadcast arrays.  `numba` is accepted and ignored."""
        args = [v for v in (inc, wspd, phi) if v is not None]
        all_scalar = all(np.isscalar(v) for v in args)
        all_1d = all(hasattr(v, "ndim") and v.ndim == 1 for v in args)
        if any(hasattr(v, "ndim") and v.ndim > 1 for v in args):
            broadcast = True
        has_phi = phi is not None
        if has_phi != (self.phi_range is not None):
            if has_phi:   # cross-pol gmfs accept and ignore phi (gmfs_impl.py:326)
                pass
            else:
                raise TypeError(f"{self.name} needs phi")
        if _xr.is_dask(inc) or _xr.is_dask(wspd) or (has_phi and _xr.is_dask(phi)):
            return self._call_dask(inc, wspd, phi)
        if broadcast:
            b = np.broadcast_arrays(*[np.asarray(v) for v in args])
            inc_b, wspd_b = b[0], b[1]
            phi_b = b[2] if has_phi else None
            vals = self._eval_broadcast(inc_b, wspd_b, phi_b if self.phi_range is not None else None)
            template = next((v for v in (inc, wspd, phi) if _xr.is_labelled(v)), None)
            if template is not None and tuple(template.shape) == vals.shape:
                out = _xr.like(template, vals.astype(np.float64) if vals.dtype != np.float32 else vals)
            else:
                out = vals
        elif all_scalar:
            return float(self._eval_broadcast(np.float64(inc).reshape(1), np.float64(wspd).reshape(1),
                                              np.float64(phi).reshape(1) if self.phi_range is not None else None)[0])
        elif all_1d:
            gi, gw = np.asarray(inc, dtype=np.float64), np.asarray(wspd, dtype=np.float64)
            gp = np.asarray(phi, dtype=np.float64) if has_phi else None
            vals = self._eval_outer(gi, gw, gp if self.phi_range is not None else None).cpu().numpy()
            if has_phi and self.phi_range is None:
                vals = np.repeat(vals[..., None], gp.size, axis=-1)
            defaults = [("incidence", inc), ("wspd", wspd)] + ([("phi", phi)] if has_phi else [])
            dims = [v.dims[0] if hasattr(v, "dims") else d for d, v in defaults]
            coords = {dim: np.asarray(v) for dim, (_, v) in zip(dims, defaults)}
            out = _xr.make_dataarray(vals, dims, coords)
        else:
            raise ValueError("Non 1d shape must all have the same shape")
        try:
            out.attrs["units"] = self.units
        except AttributeError:
            pass
        return out

    def _call_dask(self, inc, wspd, phi):  # pragma: no cover - dask is absent from the build image
        import dask.array as da

        arrs = [v.data if _xr.is_labelled(v) else v for v in (inc, wspd) + ((phi,) if phi is not None else ())]
        arrs = da.broadcast_arrays(*arrs)
        has_phi = phi is not None

        def block(*blk):
            return self._eval_broadcast(blk[0], blk[1], blk[2] if has_phi and self.phi_range is not None else None)

        res = da.map_blocks(block, *arrs, dtype=np.float64)
        template = next((v for v in (inc, wspd, phi) if _xr.is_labelled(v)), None)
        return _xr.like(template, res, attrs=dict(units=self.units)) if template is not None else res

    # -- LUT ------------------------------------------------------------------------------------------------------
    def _resolve_raw(self, kwargs, mutate):
        """Resolution / step selection of `_raw_lut` (gmfs.py:353-379); mutates self.*_step like the reference."""
        resolution = kwargs.pop("resolution", "low")
        if resolution not in ["low", "high", None]:
            raise ValueError('kwargs resolution must be "low" or "high" or None, or not provided')
        if resolution is None:
            resolution = "low" if self.iscopol else "high"
        sfx = "_lr" if resolution == "low" else ""
        steps = []
        for k in ("inc_step", "wspd_step", "phi_step"):
            v = kwargs.pop(k + sfx, getattr(self, k + sfx))
            if mutate:
                setattr(self, k + sfx, v)
            steps.append(v)
        return resolution, steps

    def _apply_step_side_effects(self, **kwargs):
        self._resolve_raw(kwargs, mutate=True)

    def _raw_lut_device(self, **kwargs) -> DeviceLut:
        resolution, (inc_step, wspd_step, phi_step) = self._resolve_raw(kwargs, mutate=True)
        inc, wspd = _grid(self.inc_range, inc_step), _grid(self.wspd_range, wspd_step)
        phi = _grid(self.phi_range, phi_step) if self.phi_range is not None else None
        return DeviceLut(self._eval_outer(inc, wspd, phi), inc, wspd, phi, self.units, resolution)
