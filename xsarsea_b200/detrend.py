"""sigma0_detrend -- counterpart of xsarsea/detrend.py:9-68.

The GMF profile of the first image line is evaluated on the device (`xs_gmf_eval` for analytic models, LUT
interpolation for file-backed ones), normalised by its nanmean, and the raster is divided by it in one streaming
pass (`xs_detrend`, HBM-bound).
"""
from __future__ import annotations

import logging

import numpy as np

from . import _device as dev
from . import _native as nat
from . import _xr
from .windspeed.models import get_model

logger = logging.getLogger("xsarsea")


def sigma0_detrend(sigma0, inc_angle, wind_speed_gmf=np.array([10.0]), wind_dir_gmf=np.array([45.0]),
                   model="gmf_cmod5n"):
    """compute `sigma0_detrend` from `sigma0` and `inc_angle` (detrend.py:9-68).

    sigma0 : linear sigma0, dims (..., line, sample) -- a leading `pol` dim is allowed, every pol gets the same ratio
    inc_angle : labelled incidence angle (deg) with a `line` dim; only its first line is used (detrend.py:55)
    wind_speed_gmf, wind_dir_gmf : 0-D or size-1 numpy arrays (m/s, deg relative to antenna)
    """
    model = get_model(model)
    if wind_speed_gmf.ndim > 1 or wind_dir_gmf.ndim > 1:
        raise ValueError("wind_speed_gmf and wind_dir_gmf must be 0D or 1D")
    for var in [wind_speed_gmf, wind_dir_gmf]:
        if var.ndim == 1 and var.size > 1:
            raise ValueError("wind_speed_gmf and wind_dir_gmf size must be 1 or 0")
    torch = nat.torch_cuda()
    resident = type(sigma0).__module__.split(".")[0] == "torch"
    if resident:
        # device-resident extension: CUDA tensors in ([..., line, sample] sigma0, [line, sample] incidence), CUDA tensor out
        inc_t = inc_angle if type(inc_angle).__module__.split(".")[0] == "torch" else torch.as_tensor(np.asarray(inc_angle))
        t_inc = inc_t.reshape(-1, inc_t.shape[-1])[0].detach().cuda().to(torch.float64).contiguous()   # stays on the device: no sync
        inc_line = None
    else:
        inc_line = np.asarray(inc_angle.isel(line=0).data, dtype=np.float64).reshape(-1)  # needs a labelled array, like the reference
        t_inc = dev.to_device(inc_line)
    wspd = float(np.asarray(wind_speed_gmf).reshape(-1)[0])
    phi = float(np.asarray(wind_dir_gmf).reshape(-1)[0])

    if getattr(model, "_device_id", None) is not None:
        t_w = torch.full_like(t_inc, wspd)
        t_p = torch.full_like(t_inc, phi) if model.phi_range is not None else None
        profile = dev.gmf_eval(model._device_id, t_inc, t_w, t_p)
    else:
        # LUT-backed or host-defined model: evaluate the profile through the model's own call
        if inc_line is None:
            inc_line = t_inc.cpu().numpy()
        if model.phi_range is not None:
            vals = model(inc_line, np.array([wspd]), np.array([phi]))
        else:
            vals = model(inc_line, np.array([wspd]))
        profile = dev.to_device(np.asarray(vals, dtype=np.float64).reshape(-1))

    w = int(t_inc.numel())
    if resident:
        if sigma0.shape[-1] != w:
            raise ValueError(f"sigma0 sample axis ({sigma0.shape[-1]}) does not match inc_angle ({w})")
        t_s0 = sigma0.cuda()
        if t_s0.dtype not in (torch.float32, torch.float64):
            t_s0 = t_s0.to(torch.float64)
        return dev.detrend(t_s0.reshape(-1, w), profile).reshape(sigma0.shape)
    s0 = np.asarray(sigma0.data if _xr.is_labelled(sigma0) else sigma0)
    if s0.shape[-1] != w:
        raise ValueError(f"sigma0 sample axis ({s0.shape[-1]}) does not match inc_angle ({w})")
    f32 = s0.dtype == np.float32
    t_s0 = dev.to_device(np.ascontiguousarray(s0, dtype=np.float32 if f32 else np.float64).reshape(-1, w))
    out = dev.detrend(t_s0, profile).cpu().numpy().reshape(s0.shape)
    if _xr.is_labelled(sigma0):
        res = _xr.like(sigma0, out, name=getattr(sigma0, "name", None), attrs=dict(getattr(sigma0, "attrs", {})))
        res.attrs["comment"] = f"detrended with model {model.name}"
        return res
    return out


# ---- direction-convention helpers (reference detrend.py:96-201) --------------------------
# One-line conversions applied by callers to the inversion output (antenna convention <-> meteorological /
# oceanographic conventions).  Host-side numpy: they work on scalars, numpy arrays and labelled arrays alike.

def dir_meteo_to_sample(meteo_dir, ground_heading):
    """Meteorological N/S direction (deg, clockwise from north, "from") -> image (sample-axis) convention, radians."""
    return np.pi / 2 - np.deg2rad(meteo_dir - ground_heading)


def dir_sample_to_meteo(sample_dir, ground_heading):
    """Image convention (deg, relative to the sample axis) -> meteorological direction (deg)."""
    return 90 - sample_dir + ground_heading


def dir_meteo_to_oceano(meteo_dir):
    """Meteorological ("from") -> oceanographic ("to") convention, degrees in [0, 360)."""
    return (meteo_dir + 180) % 360


def dir_oceano_to_meteo(oceano_dir):
    """Oceanographic ("to") -> meteorological ("from") convention, degrees in [0, 360)."""
    return (oceano_dir - 180) % 360


def dir_to_180(angle):
    """Wrap an angle in degrees to [-180, 180)."""
    return (angle + 180) % 360 - 180


def dir_to_360(angle):
    """Wrap an angle in degrees to [0, 360)."""
    return (angle + 360) % 360
