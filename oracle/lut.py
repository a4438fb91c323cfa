"""ORACLE (test infrastructure): host-side LUT recipes of the reference, restated functionally.

Follows windspeed/gmfs.py:350-395 (`GmfModel._raw_lut`: grids + resolution default "low"),
windspeed/models.py:82-174 (`_normalize_lut`: resolution default "high", interpolation decision, target
grids) and models.py:186-230 (`to_lut`: unit conversion).  SURVEY.md appendix A.1 is the decision table.
"""
from __future__ import annotations

import numpy as np

from . import COPOL_MODELS, MODEL_IDS, interp_axis, lut_build, to_db, to_linear

# models.py:38-48 defaults
DEFAULT_STEPS = dict(inc_step_lr=1.0, wspd_step_lr=0.2, phi_step_lr=2.5, inc_step=0.1, wspd_step=0.1, phi_step=1.0)


def model_meta(name: str) -> dict:
    """Registration attributes of the 13 built-in GMFs (gmfs_impl.py:116,213,325,...; gmfs.py:87-95,134-158)."""
    assert name in MODEL_IDS
    if name in COPOL_MODELS:
        pol = "HH" if "_pr_" in name else "VV"
        return dict(pol=pol, units="linear", inc_range=[16.0, 66.0], wspd_range=[0.2, 50.0], phi_range=[0.0, 180.0])
    return dict(pol="VH", units="linear", inc_range=[16.0, 66.0], wspd_range=[3.0, 80.0], phi_range=None)


def grid(r, step):
    """gmfs.py:385-390 / models.py:154-160"""
    if r is None:
        return None
    return np.linspace(r[0], r[1], num=int(np.round((r[1] - r[0]) / step) + 1))


def raw_lut(name: str, **kwargs):
    """GmfModel._raw_lut: returns (lut_linear, (inc, wspd, phi|None), resolution)."""
    meta = model_meta(name)
    steps = dict(DEFAULT_STEPS)
    resolution = kwargs.pop("resolution", "low")
    if resolution not in ("low", "high", None):
        raise ValueError('kwargs resolution must be "low" or "high" or None, or not provided')
    copol = meta["phi_range"] is not None
    if resolution is None:
        resolution = "low" if copol else "high"
    sfx = "_lr" if resolution == "low" else ""
    inc_step = kwargs.pop("inc_step" + sfx, steps["inc_step" + sfx])
    wspd_step = kwargs.pop("wspd_step" + sfx, steps["wspd_step" + sfx])
    phi_step = kwargs.pop("phi_step" + sfx, steps["phi_step" + sfx])
    inc = grid(meta["inc_range"], inc_step)
    wspd = grid(meta["wspd_range"], wspd_step)
    phi = grid(meta["phi_range"], phi_step)
    lut = lut_build(name, inc, wspd, phi)
    model_steps = dict(steps)
    model_steps.update({"inc_step" + sfx: inc_step, "wspd_step" + sfx: wspd_step, "phi_step" + sfx: phi_step})
    return lut, (inc, wspd, phi), resolution, model_steps


def normalize_lut(name, lut, grids, lut_resolution, model_steps, **kwargs):
    """Model._normalize_lut (models.py:108-172) for a linear-unit analytic LUT."""
    meta = model_meta(name)
    copol = meta["phi_range"] is not None
    resolution = kwargs.pop("resolution", "high")
    if resolution is None:
        resolution = "high"
    sfx = "_lr" if resolution == "low" else ""
    if resolution == lut_resolution:
        keys = ["inc_step" + sfx, "wspd_step" + sfx] + (["phi_step" + sfx] if copol else [])
        do_interp = any(model_steps[k] != kwargs.get(k, model_steps[k]) for k in keys)
    else:
        do_interp = False
    if resolution != lut_resolution or do_interp:
        inc_step = kwargs.pop("inc_step" + sfx, model_steps["inc_step" + sfx])
        wspd_step = kwargs.pop("wspd_step" + sfx, model_steps["wspd_step" + sfx])
        phi_step = kwargs.pop("phi_step" + sfx, model_steps["phi_step" + sfx])
        new = (grid(meta["inc_range"], inc_step), grid(meta["wspd_range"], wspd_step),
               grid(meta["phi_range"], phi_step))
        for axis, (xs, xd) in enumerate(zip(grids, new)):
            if xd is not None:
                lut = interp_axis(lut, axis, xs, xd)
        grids = new
    return lut, grids, resolution


def to_lut(name: str, units="linear", **kwargs):
    """Model.to_lut for an analytic GMF: returns (lut, (inc, wspd, phi|None)).

    kwargs are consumed twice exactly as models.py:201-203 does (raw_lut and normalize_lut each get a copy).
    """
    lut, grids, res, model_steps = raw_lut(name, **dict(kwargs))
    lut, grids, _ = normalize_lut(name, lut, grids, res, model_steps, **dict(kwargs))
    if units == "dB":
        lut = to_db(lut)
    elif units not in ("linear", None):
        raise ValueError(f"Unit not known: {units}. Known are 'dB' or 'linear' ")
    return lut, grids


__all__ = ["model_meta", "grid", "raw_lut", "normalize_lut", "to_lut", "to_db", "to_linear"]
