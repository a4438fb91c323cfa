"""invert_from_model: wind inversion from sigma0 and a model -- counterpart of xsarsea/windspeed/windspeed.py.

Same signature, argument handling, warnings/errors, return containers and attrs as the reference
(windspeed.py:18-130 routing, :333-388 container dispatch, :395-439 attrs and returns).  The numeric part
(:132-331, the numba gufunc) is `xs_invert` on the GPU: the dB prologue (:126-128), the co-pol argmin, the
cross-pol/dual argmin and the dual-pol merge (:426-428) are all done on the device; rasters are streamed to the
GPU in row blocks on side streams so host<->device copies overlap the scan.
"""
from __future__ import annotations

import logging
import warnings
from collections import OrderedDict

import numpy as np

from .. import _device as dev
from .. import _native as nat
from .. import _xr
from .models import get_model

logger = logging.getLogger("xsarsea.windspeed")

_PLAN_CACHE: "OrderedDict[tuple, dev.InversionPlan]" = OrderedDict()
_PLAN_CACHE_MAX = 4
BLOCK_PIXELS = 1 << 25  # pixels per streamed block (32 Mi px: 1.3 GB of f64 inputs, 1 GB of outputs)


def _get_plan(model_co, model_cr, dsig_co, kwargs):
    """Device LUTs + scan image for a model pair (what windspeed.py:139-181 sets up on every call), cached."""
    luts = []
    for m in (model_co, model_cr):
        luts.append(None if m is None else m.to_lut_device(units="dB", **kwargs))
    key = (id(luts[0]), id(luts[1]), float(dsig_co))
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        co = cr = None
        if luts[0] is not None:
            l = luts[0]
            if l.phi is None:
                raise ValueError(f"model {model_co.name} has no phi dimension: not a co-pol model")
            co = (l.data, l.inc, l.wspd, l.phi)
        if luts[1] is not None:
            l = luts[1]
            if l.phi is not None:
                raise ValueError(f"model {model_cr.name} has a phi dimension: not a cross-pol model")
            cr = (l.data, l.inc, l.wspd)
        plan = dev.InversionPlan(co=co, cr=cr, dsig_co=dsig_co)
        plan._luts = luts  # keep the cached DeviceLut objects (and their ids) alive with the plan
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
            _PLAN_CACHE.popitem(last=False)[1].close()
    else:
        _PLAN_CACHE.move_to_end(key)
    return plan


def clear_plan_cache():
    while _PLAN_CACHE:
        _PLAN_CACHE.popitem()[1].close()


def _is_tensor(x):
    return type(x).__module__.split(".")[0] == "torch"


def _isnan(x):
    return x.isnan() if _is_tensor(x) else np.isnan(x)


def _values(x):
    """numpy view of a numpy / labelled input (None stays None; torch tensors pass through)."""
    if x is None or _is_tensor(x):
        return x
    return np.asarray(x.data if _xr.is_labelled(x) else x)


def _run_resident(plan, inc, s_co, s_cr, dsig_cr, anc, *, merge_dual, cr_abs):
    """Device-resident variant of `_run_device`: torch CUDA tensors in, torch CUDA tensors out, one `xs_invert` on the
    current stream, no host copies (an extension of the reference's container rule "output mirrors input",
    windspeed.py:333-388, to device arrays)."""
    torch = nat.torch_cuda()
    f32 = inc.dtype == torch.float32
    rdt, cdt = (torch.float32, torch.complex64) if f32 else (torch.float64, torch.complex128)

    def prep(x, dt):
        if x is None:
            return None
        x = x if _is_tensor(x) else torch.as_tensor(np.asarray(x))
        return x.to(device=inc.device, dtype=dt).expand(inc.shape).contiguous()

    dsig = dsig_cr if np.isscalar(dsig_cr) else prep(dsig_cr, rdt)
    oc, ox, _, _ = plan.invert(inc.contiguous(), prep(s_co, rdt), prep(s_cr, rdt), dsig, prep(anc, cdt), sigma0_db=False,
                               merge_dual=merge_dual, cr_abs=cr_abs)
    return oc, ox


def _run_device(plan, inc, s_co, s_cr, dsig_cr, anc, *, sigma0_db, merge_dual, cr_abs, mode=nat.MODE_FAST,
                need_co=False):
    """Host arrays in, host arrays out.  Streams row blocks: H2D on one side stream, xs_invert on the current
    stream, D2H on another side stream; two device slots so block k+1 uploads while block k is scanned."""
    torch = nat.torch_cuda()
    shape = inc.shape
    f32 = all(a is None or a.dtype in (np.float32, np.complex64) for a in (inc, s_co, s_cr, anc)) and (
        not isinstance(dsig_cr, np.ndarray) or dsig_cr.dtype == np.float32)
    rdt, cdt = (np.float32, np.complex64) if f32 else (np.float64, np.complex128)

    def flat(a, dt):
        return None if a is None else np.ascontiguousarray(np.broadcast_to(a, shape), dtype=dt).reshape(-1)

    h_inc, h_co, h_cr, h_anc = flat(inc, rdt), flat(s_co, rdt), flat(s_cr, rdt), flat(anc, cdt)
    h_dsig = flat(dsig_cr, rdt) if isinstance(dsig_cr, np.ndarray) and dsig_cr.ndim > 0 else None
    dsig_scalar = 0.1 if h_dsig is not None else float(dsig_cr)
    n = h_inc.size
    want_co = plan.co_grids is not None and h_co is not None
    out_co = torch.empty(n, dtype=torch.complex128, pin_memory=True) if (want_co or need_co) else None
    out_cr = torch.empty(n, dtype=torch.float64 if cr_abs else torch.complex128, pin_memory=True)
    if n == 0:
        return (None if out_co is None else out_co.numpy().reshape(shape)), out_cr.numpy().reshape(shape)

    blk = min(n, BLOCK_PIXELS)
    nslots = 1 if n <= blk else 2
    trdt, tcdt = (torch.float32, torch.complex64) if f32 else (torch.float64, torch.complex128)

    def dbuf(h, dt):
        return None if h is None else [torch.empty(blk, dtype=dt, device="cuda") for _ in range(nslots)]

    d_inc, d_co, d_cr, d_dsig, d_anc = dbuf(h_inc, trdt), dbuf(h_co, trdt), dbuf(h_cr, trdt), dbuf(h_dsig, trdt), dbuf(h_anc, tcdt)
    d_oco = None if out_co is None else [torch.empty(blk, dtype=torch.complex128, device="cuda") for _ in range(nslots)]
    d_ocr = [torch.empty(blk, dtype=out_cr.dtype, device="cuda") for _ in range(nslots)]
    cur = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_scan = [None] * nslots   # scan of the block that last used the slot's inputs
    ev_d2h = [None] * nslots    # download of the block that last used the slot's outputs
    for k, lo in enumerate(range(0, n, blk)):
        hi = min(lo + blk, n)
        m = hi - lo
        slot = k % nslots
        with torch.cuda.stream(s_in):
            if ev_scan[slot] is not None:
                s_in.wait_event(ev_scan[slot])
            for d, h in ((d_inc, h_inc), (d_co, h_co), (d_cr, h_cr), (d_dsig, h_dsig), (d_anc, h_anc)):
                if d is not None:
                    d[slot][:m].copy_(torch.from_numpy(h[lo:hi]), non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(s_in)
        cur.wait_event(ev_in)
        if ev_d2h[slot] is not None:
            cur.wait_event(ev_d2h[slot])
        sl = lambda d: None if d is None else d[slot][:m]
        plan.invert(sl(d_inc), sl(d_co), sl(d_cr), sl(d_dsig) if d_dsig is not None else dsig_scalar, sl(d_anc),
                    sigma0_db=sigma0_db, merge_dual=merge_dual, cr_abs=cr_abs, mode=mode,
                    out_co=sl(d_oco), out_cr=sl(d_ocr))
        ev_scan[slot] = torch.cuda.Event()
        ev_scan[slot].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_scan[slot])
            if out_co is not None:
                out_co[lo:hi].copy_(d_oco[slot][:m], non_blocking=True)
            out_cr[lo:hi].copy_(d_ocr[slot][:m], non_blocking=True)
            ev_d2h[slot] = torch.cuda.Event()
            ev_d2h[slot].record(s_out)
    s_out.synchronize()
    cur.synchronize()
    return (None if out_co is None else out_co.numpy().reshape(shape)), out_cr.numpy().reshape(shape)


def _invert_from_model_numpy(models, dsig_co, kwargs, np_inc, np_sigma0_co_db, np_sigma0_cr_db, np_dsig_cr,
                             np_ancillary_wind):
    """The numpy-level operator boundary of the reference (windspeed.py:132-134, gufunc
    "(n),(n),(n),(n),(n)->(n),(n)", f64 x4 + c128 -> c128 x2): sigma0 already in dB, an absent polarisation is an
    all-NaN raster, outputs are (wind_co, wind_dual) complex128.  Used directly by the dask path."""
    np_inc = np.asarray(np_inc)
    co_db, cr_db = np.asarray(np_sigma0_co_db), np.asarray(np_sigma0_cr_db)
    have_cr = not np.all(np.isnan(cr_db))                    # windspeed.py:170
    plan = _get_plan(models[0], models[1] if have_cr else None, dsig_co, kwargs)
    oc, ox = _run_device(plan, np_inc, co_db if models[0] is not None else None, cr_db if have_cr else None,
                         np.asarray(np_dsig_cr), np.asarray(np_ancillary_wind), sigma0_db=True, merge_dual=False,
                         cr_abs=False, need_co=True)
    if oc is None:
        oc = np.full(np_inc.shape, complex(np.nan, np.nan))
    return oc, ox


def invert_from_model(inc, sigma0, sigma0_dual=None, /, ancillary_wind=None, dsig_co=0.1, dsig_cr=0.1, model=None,
                      **kwargs):
    """Invert sigma0 to retrieve windspeed from model (lut or gmf).  Drop-in for
    xsarsea.windspeed.invert_from_model (windspeed.py:18-439).

    Parameters
    ----------
    inc : incidence angle (numpy, xarray or dask-backed xarray)
    sigma0 : sigma0 to be inverted (linear)
    sigma0_dual : sigma0 in cross pol for dual-pol inversion (optional)
    ancillary_wind : complex ancillary wind (e.g. ecmwf), antenna convention
    dsig_co : float, `Jsig_co = ((sigma0_gmf - sigma0) / dsig_co) ** 2`
    dsig_cr : float or array, `Jsig_cr = ((sigma0_gmf - sigma0) / dsig_cr) ** 2`
    model : str | Model | (model_co, model_cr)
    **kwargs : forwarded to `Model.to_lut` (resolution, inc_step, wspd_step, phi_step, ...)

    Returns
    -------
    co-pol only: complex128 wind (abs = m/s, angle = direction, antenna convention); cross-pol only: float64 wind
    speed; dual-pol: tuple (wind_co, wind_dual) with wind_dual merged as in windspeed.py:426-428.
    """
    models = model if isinstance(model, tuple) else (model, None)
    models = tuple(get_model(m) if m is not None else None for m in models)
    anc_given = ancillary_wind is not None

    if sigma0_dual is None:
        # mono-pol inversion (windspeed.py:86-115)
        try:
            pol = sigma0.pol.values.item()
        except AttributeError:
            pol = None
        model_pol = models[0].pol
        if pol is None:
            warnings.warn(f"Unable to check sigma0 pol. Assuming  {model_pol}")
        elif pol not in model_pol:
            raise ValueError(f"sigma0 pol is {pol}, and model {models[0].name} can only handle {model_pol}")
        if models[0].iscopol:
            sigma0_co, sigma0_cr = sigma0, None
            # copol needs valid ancillary wind
            assert anc_given and bool((~_isnan(_values(ancillary_wind))).any())
        elif models[0].iscrosspol:
            sigma0_co, sigma0_cr = None, sigma0
            if anc_given and not bool(_isnan(_values(ancillary_wind)).all()):
                warnings.warn("crosspol inversion is best without ancillary wind, but using it as requested.")
            models = (None, models[0])
    else:
        sigma0_co, sigma0_cr = sigma0, sigma0_dual

    template = sigma0_co if sigma0_co is not None else sigma0_cr
    dual = sigma0_dual is not None
    cross_only = sigma0_co is None

    if any(_xr.is_dask(v) for v in (inc, sigma0_co, sigma0_cr, ancillary_wind, dsig_cr) if v is not None):
        ws_co, ws_cr_or_dual = _invert_dask(models, dsig_co, kwargs, inc, sigma0_co, sigma0_cr, dsig_cr, ancillary_wind,
                                            template)
        if cross_only:
            ws_cr_or_dual = abs(ws_cr_or_dual)
        elif dual:
            import xarray as xr

            ws_cr_or_dual = xr.where((abs(ws_co) < 5) | (abs(ws_cr_or_dual) < 5), ws_co, ws_cr_or_dual)
    elif _is_tensor(inc):
        plan = _get_plan(models[0], models[1] if sigma0_cr is not None else None, dsig_co, kwargs)
        ws_co, ws_cr_or_dual = _run_resident(plan, inc, sigma0_co, sigma0_cr, dsig_cr, ancillary_wind if anc_given else None,
                                             merge_dual=dual, cr_abs=cross_only)
    else:
        plan = _get_plan(models[0], models[1] if sigma0_cr is not None else None, dsig_co, kwargs)
        dsig_in = dsig_cr if np.isscalar(dsig_cr) else _values(dsig_cr)
        oc, ox = _run_device(plan, _values(inc), _values(sigma0_co), _values(sigma0_cr), dsig_in,
                             _values(ancillary_wind) if anc_given else None, sigma0_db=False, merge_dual=dual,
                             cr_abs=cross_only)
        if _xr.is_labelled(template):
            ws_co = None if oc is None else _xr.like(template, oc, name="windspeed_gmf")
            ws_cr_or_dual = _xr.like(template, ox, name="windspeed_gmf")
        else:
            ws_co, ws_cr_or_dual = oc, ox

    # attrs and returns, windspeed.py:395-439
    if models[0] is not None and models[0].iscopol and hasattr(ws_co, "attrs"):
        ws_co.attrs["comment"] = f"wind speed and direction inverted from model {models[0].name} ({models[0].pol})"
        ws_co.attrs["model"] = models[0].name
    if not dual:
        if not cross_only:
            return ws_co
        if hasattr(ws_cr_or_dual, "attrs"):
            ws_cr_or_dual.attrs["comment"] = f"wind speed inverted from model {models[1].name} ({models[1].pol})"
            ws_cr_or_dual.attrs["model"] = models[1].name
            ws_cr_or_dual.attrs["units"] = "m/s"
        return ws_cr_or_dual
    if hasattr(ws_cr_or_dual, "attrs"):
        ws_cr_or_dual.attrs["comment"] = (f"wind speed and direction inverted from model {models[0].name} "
                                          f"({models[0].pol}) and {models[1].name} ({models[1].pol})")
        ws_cr_or_dual.attrs["model"] = f"{models[0].name} {models[1].name}"
    return ws_co, ws_cr_or_dual


def _invert_dask(models, dsig_co, kwargs, inc, sigma0_co, sigma0_cr, dsig_cr, ancillary_wind, template):
    """dask-backed inputs: lazy `da.apply_gufunc` over blocks with the sample axis as core dimension, like
    windspeed.py:356-364; every block goes through the numpy-level operator (and hence the GPU)."""
    import dask.array as da  # pragma: no cover - dask is absent from the build image
    import xarray as xr

    nan = template * np.nan
    sigma0_co = nan if sigma0_co is None else sigma0_co
    sigma0_cr = nan if sigma0_cr is None else sigma0_cr
    ancillary_wind = nan if ancillary_wind is None else ancillary_wind
    if np.isscalar(dsig_cr):
        dsig_cr = sigma0_cr * 0 + dsig_cr
    co_db = 10 * np.log10(sigma0_co + 1e-15)
    cr_db = 10 * np.log10(sigma0_cr + 1e-15)

    def block(i, c, x, d, a):
        return _invert_from_model_numpy(models, dsig_co, kwargs, i, c, x, d, a)

    data = [v.data if _xr.is_labelled(v) else v for v in (inc, co_db, cr_db, dsig_cr, ancillary_wind)]
    oc, ox = da.apply_gufunc(block, "(n),(n),(n),(n),(n)->(n),(n)", *data,
                             output_dtypes=(np.complex128, np.complex128))
    ws_co = xr.zeros_like(template, dtype=np.complex128)
    ws_co.name = "windspeed_gmf"
    ws_co.attrs.clear()
    ws_cr = ws_co.copy()
    ws_co.data, ws_cr.data = oc, ox
    return ws_co, ws_cr
