"""invert_from_model: wind inversion from sigma0 and a model -- counterpart of xsarsea/windspeed/windspeed.py.

Same signature, argument handling, warnings/errors, return containers and attrs as the reference
(windspeed.py:18-130 routing, :333-388 container dispatch, :395-439 attrs and returns).  The numeric part
(:132-331, the numba gufunc) is `xs_invert` on the GPU: the dB prologue (:126-128), the co-pol argmin, the
cross-pol/dual argmin and the dual-pol merge (:426-428) are all done on the device; rasters are streamed to the
GPU in row blocks on side streams so host<->device copies overlap the scan.
"""
from __future__ import annotations

import logging
import os
import queue
import threading
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
_PLAN_LOCK = threading.RLock()
BLOCK_PIXELS = int(os.environ.get("XS_BLOCK_PIXELS", 1 << 26))  # pixels per compute block of the host path (64 Mi px: 2.7 GB of f64 inputs, 2.1 GB of outputs, 5.6 GB of workspace)
STAGE_PIXELS = int(os.environ.get("XS_STAGE_PIXELS", 1 << 24))  # pixels per pinned staging chunk (16 Mi px: 0.27 GB per complex128 raster)


def _get_plan(model_co, model_cr, dsig_co, kwargs):
    """Device LUTs + scan image for a model pair (what windspeed.py:139-181 sets up on every call), cached.

    Thread-safe: LUT building, plan creation and cache eviction happen under one lock; an evicted plan is only dropped
    from the cache -- it is destroyed when the last caller still holding it lets go (`InversionPlan.__del__`/`close`
    defer while calls are in flight), so eviction can never pull a plan from under a running inversion."""
    with _PLAN_LOCK:
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
            plan = dev.InversionPlan(co=co, cr=cr, dsig_co=dsig_co)  # synchronises the stream the LUTs were built on
            plan._luts = luts  # keep the cached DeviceLut objects (and their ids) alive with the plan
            _PLAN_CACHE[key] = plan
            while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
                _PLAN_CACHE.popitem(last=False)  # dropped, not closed: users still holding it keep it alive
        else:
            _PLAN_CACHE.move_to_end(key)
        return plan


def clear_plan_cache():
    with _PLAN_LOCK:
        _PLAN_CACHE.clear()


def _is_tensor(x):
    return type(x).__module__.split(".")[0] == "torch"


def _isnan(x):
    return x.isnan() if _is_tensor(x) else np.isnan(x)


def _values(x):
    """numpy view of a numpy / labelled input (None stays None; torch tensors pass through)."""
    if x is None or _is_tensor(x):
        return x
    return np.asarray(x.data if _xr.is_labelled(x) else x)


def _run_resident(plan, inc, s_co, s_cr, dsig_cr, anc, *, merge_dual, cr_abs, speed_dir=False, ground_heading=None,
                  out_f32=False):
    """Device-resident variant of `_run_device`: torch CUDA tensors in, torch CUDA tensors out, one `xs_invert` on the
    current stream, no host copies (an extension of the reference's container rule "output mirrors input",
    windspeed.py:333-388, to device arrays).  float32 arithmetic inputs are used as such only when every raster is
    float32 / complex64; a mix is promoted to float64 / complex128 (like the host path and numpy's own promotion)."""
    torch = nat.torch_cuda()
    given = [x for x in (inc, s_co, s_cr, anc, ground_heading) + (() if np.isscalar(dsig_cr) else (dsig_cr,))
             if x is not None and hasattr(x, "dtype")]
    f32 = all(str(x.dtype).split(".")[-1] in ("float32", "complex64") for x in given)
    rdt, cdt = (torch.float32, torch.complex64) if f32 else (torch.float64, torch.complex128)

    def prep(x, dt):
        if x is None:
            return None
        x = x if _is_tensor(x) else torch.as_tensor(np.asarray(x))
        return x.to(device=inc.device, dtype=dt).expand(inc.shape).contiguous()

    dsig = dsig_cr if np.isscalar(dsig_cr) else prep(dsig_cr, rdt)
    gh = ground_heading if (ground_heading is None or np.isscalar(ground_heading)) else prep(ground_heading, rdt)
    oc, ox, _, _ = plan.invert(prep(inc, rdt), prep(s_co, rdt), prep(s_cr, rdt), dsig, prep(anc, cdt), sigma0_db=False,
                               merge_dual=merge_dual, cr_abs=cr_abs, speed_dir=speed_dir, ground_heading=gh, out_f32=out_f32)
    return oc, ox


class _PinnedPool:
    """Block-sized pinned staging buffers, reused across calls (and handed out per call, so concurrent calls never
    share one).  Only these are page-locked: inputs are read from, and results delivered into, ordinary pageable numpy
    arrays, so repeated calls on differently sized scenes do not accumulate locked RAM."""

    def __init__(self, keep=12):
        self._free, self._lock, self._keep = {}, threading.Lock(), keep

    def take(self, nbytes):
        torch = nat.torch_cuda()
        with self._lock:
            lst = self._free.get(nbytes)
            if lst:
                return lst.pop()
        return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)

    def give(self, buf):
        with self._lock:
            lst = self._free.setdefault(buf.numel(), [])
            if sum(len(v) for v in self._free.values()) < self._keep:
                lst.append(buf)


_POOL = _PinnedPool()
_COPY_THREADS = int(os.environ.get("XS_COPY_THREADS", 8))
_COPY_POOL = None


def _par_copy(dst, src):
    """dst[...] = src for large 1-D / [k, m] host arrays, split along the last axis between a few threads (numpy releases the
    GIL while it copies; one thread moves ~6-10 GB/s, and fresh pageable destinations also take their first-touch page
    faults here)."""
    global _COPY_POOL
    m = dst.shape[-1]
    if m < (1 << 20):
        np.copyto(dst, src)
        return
    if _COPY_POOL is None:
        with _PLAN_LOCK:
            if _COPY_POOL is None:
                from concurrent.futures import ThreadPoolExecutor

                _COPY_POOL = ThreadPoolExecutor(max_workers=4 * _COPY_THREADS, thread_name_prefix="xs-copy")
    cuts = [m * k // _COPY_THREADS for k in range(_COPY_THREADS + 1)]
    futs = [_COPY_POOL.submit(np.copyto, dst[..., a:b], src[..., a:b]) for a, b in zip(cuts[1:], cuts[2:])]
    np.copyto(dst[..., :cuts[1]], src[..., :cuts[1]])
    for f in futs:
        f.result()


MIN_BLOCK_PIXELS = 1 << 22  # a raster is split for overlap only into blocks of at least 4 Mi px


def _block_edges(n):
    """Compute blocks of the host path: BLOCK_PIXELS each, with a short first and last one (STAGE_PIXELS) so that the
    upload that cannot overlap anything (the first) and the download that cannot (the last) are small.  A raster below
    BLOCK_PIXELS is cut into four blocks (of at least MIN_BLOCK_PIXELS) so that its copies overlap its kernels too."""
    if n <= BLOCK_PIXELS:
        k = min(4, n // MIN_BLOCK_PIXELS)
        if k < 2:
            return [0, n]
        return [n * i // k for i in range(k)] + [n]
    short = min(STAGE_PIXELS, BLOCK_PIXELS)
    edges = [0, short]
    while n - edges[-1] > BLOCK_PIXELS + short:
        edges.append(edges[-1] + BLOCK_PIXELS)
    if n - edges[-1] > short:
        edges.append(n - short)
    edges.append(n)
    return edges


def _run_device(plan, inc, s_co, s_cr, dsig_cr, anc, *, sigma0_db, merge_dual, cr_abs, mode=nat.MODE_FAST,
                need_co=False, speed_dir=False, ground_heading=None, out_f32=False):
    """Host arrays in, host arrays out.

    The raster is inverted in compute blocks of BLOCK_PIXELS (large, because the co-pol scan shares work between pixels of
    equal incidence bin and sigma0: the more pixels a call sees, the more of it runs in the fast mode) held in two device
    slots: while block b is scanned, block b + 1 is uploaded (H2D on a side stream) and block b - 1 downloaded (D2H on
    another).  Host memory that is not page-locked is staged through STAGE_PIXELS-sized pinned buffers -- inputs by this
    thread, results by a helper thread that copies every finished chunk into the ordinary (pageable) output arrays while
    the GPU works on."""
    torch = nat.torch_cuda()
    shape = inc.shape
    rasters = (inc, s_co, s_cr, anc, ground_heading if isinstance(ground_heading, np.ndarray) else None)
    f32 = all(a is None or a.dtype in (np.float32, np.complex64) for a in rasters) and (
        not isinstance(dsig_cr, np.ndarray) or dsig_cr.dtype == np.float32)
    rdt, cdt = (np.float32, np.complex64) if f32 else (np.float64, np.complex128)

    def flat(a, dt):
        return None if a is None else np.ascontiguousarray(np.broadcast_to(a, shape), dtype=dt).reshape(-1)

    h_dsig = flat(dsig_cr, rdt) if isinstance(dsig_cr, np.ndarray) and dsig_cr.ndim > 0 else None
    h_gh = flat(ground_heading, rdt) if isinstance(ground_heading, np.ndarray) and ground_heading.ndim > 0 else None
    gh_scalar = None if (ground_heading is None or h_gh is not None) else float(ground_heading)
    h_in = [flat(inc, rdt), flat(s_co, rdt), flat(s_cr, rdt), h_dsig, flat(anc, cdt), h_gh]
    dsig_scalar = 0.1 if h_dsig is not None else float(dsig_cr)
    n = h_in[0].size
    want_co = plan.co_grids is not None and h_in[1] is not None
    pdt = np.float32 if out_f32 else np.float64
    planes = 2 if speed_dir else 1
    wind_dt = pdt if speed_dir else np.complex128
    out_co = np.empty((planes, n), dtype=wind_dt) if (want_co or need_co) else None
    out_cr = np.empty((1, n), dtype=np.float64) if cr_abs else np.empty((planes, n), dtype=wind_dt)

    def finish(o):
        return None if o is None else o.reshape(((2,) if o.shape[0] == 2 else ()) + tuple(shape))

    if n == 0:
        return finish(out_co), finish(out_cr)

    edges = _block_edges(n)
    nb = len(edges) - 1
    blk = max(b - a for a, b in zip(edges, edges[1:]))
    stg = min(blk, STAGE_PIXELS)
    nslots = 1 if nb == 1 else 2
    t_of = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
            np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}
    cur = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    # the device slots come from the caching allocator of `cur`: order the side streams behind whatever `cur` still has
    # in flight on recycled blocks (they are all idle again when this function returns: it ends with full synchronisation)
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    d_in = [None if h is None else [torch.empty(blk, dtype=t_of[h.dtype], device="cuda") for _ in range(nslots)] for h in h_in]
    d_oco = None if out_co is None else [torch.empty(out_co.shape[0] * blk, dtype=t_of[out_co.dtype], device="cuda") for _ in range(nslots)]
    d_ocr = [torch.empty(out_cr.shape[0] * blk, dtype=t_of[out_cr.dtype], device="cuda") for _ in range(nslots)]
    # pinned staging rings (2 chunks each): inputs that are not page-locked already, and both outputs
    taken = []

    def pinned(nelem, np_dt):
        buf = _POOL.take(int(nelem) * np.dtype(np_dt).itemsize)
        taken.append(buf)
        return buf.numpy().view(np_dt)

    in_stage = [None if (h is None or torch.from_numpy(h).is_pinned()) else [pinned(stg, h.dtype) for _ in range(2)] for h in h_in]
    oco_stage = None if out_co is None else [pinned(out_co.shape[0] * stg, out_co.dtype) for _ in range(2)]
    ocr_stage = [pinned(out_cr.shape[0] * stg, out_cr.dtype) for _ in range(2)]
    ev_stage_in = [None, None]            # H2D that last read the input staging chunk
    stage_free = [threading.Semaphore(1), threading.Semaphore(1)]   # output staging chunk has been drained
    ev_scan = [None] * nslots             # scan of the block that last used the device slot
    ev_d2h = [None] * nslots              # last download out of the device slot's outputs
    jobs: "queue.Queue" = queue.Queue()
    errors = []
    counters = dict(in_chunk=0, out_chunk=0)

    def drain():
        # the outputs are fresh pageable memory: the copy out of the staging chunk also takes the first-touch page faults,
        # so every finished chunk is split between a few threads (_par_copy)
        while True:
            job = jobs.get()
            if job is None:
                return
            ev, r, lo, m = job
            try:
                ev.synchronize()
                if out_co is not None:
                    k = out_co.shape[0]
                    _par_copy(out_co[:, lo:lo + m], oco_stage[r][:k * m].reshape(k, m))
                k = out_cr.shape[0]
                _par_copy(out_cr[:, lo:lo + m], ocr_stage[r][:k * m].reshape(k, m))
            except Exception as e:  # pragma: no cover
                errors.append(e)
            finally:
                stage_free[r].release()

    def upload(b):
        lo, hi = edges[b], edges[b + 1]
        slot = b % nslots
        with torch.cuda.stream(s_in):
            if ev_scan[slot] is not None:
                s_in.wait_event(ev_scan[slot])   # the slot's inputs are read by the scan of block b - 2 until then
            for d, h, st in zip(d_in, h_in, in_stage):
                if d is None:
                    continue
                if st is None:   # page-locked already: one direct copy
                    d[slot][:hi - lo].copy_(torch.from_numpy(h[lo:hi]), non_blocking=True)
                    continue
                for a in range(lo, hi, stg):
                    m = min(stg, hi - a)
                    r = counters["in_chunk"] % 2
                    counters["in_chunk"] += 1
                    if ev_stage_in[r] is not None:
                        ev_stage_in[r].synchronize()
                    _par_copy(st[r][:m], h[a:a + m])
                    d[slot][a - lo:a - lo + m].copy_(torch.from_numpy(st[r][:m]), non_blocking=True)
                    ev_stage_in[r] = torch.cuda.Event()
                    ev_stage_in[r].record(s_in)
            ev = torch.cuda.Event()
            ev.record(s_in)
        return ev

    def outs(b):
        """Contiguous views [planes * m] of the slot's output buffers for block b (planes are [planes][m])."""
        m = edges[b + 1] - edges[b]
        slot = b % nslots
        oc = None if d_oco is None else d_oco[slot][:out_co.shape[0] * m]
        return oc, d_ocr[slot][:out_cr.shape[0] * m]

    def invert(b, ev_in):
        m = edges[b + 1] - edges[b]
        slot = b % nslots
        cur.wait_event(ev_in)
        if ev_d2h[slot] is not None:
            cur.wait_event(ev_d2h[slot])         # results of block b - 2 have left the slot
        sl = lambda d: None if d is None else d[slot][:m]
        oc, ox = outs(b)
        shp = (lambda t, k: t if k == 1 else t.reshape(k, m))
        plan.invert(sl(d_in[0]), sl(d_in[1]), sl(d_in[2]), sl(d_in[3]) if d_in[3] is not None else dsig_scalar,
                    sl(d_in[4]), sigma0_db=sigma0_db, merge_dual=merge_dual, cr_abs=cr_abs, mode=mode,
                    out_co=None if oc is None else shp(oc, out_co.shape[0]), out_cr=shp(ox, out_cr.shape[0]),
                    speed_dir=speed_dir, out_f32=out_f32,
                    ground_heading=sl(d_in[5]) if d_in[5] is not None else gh_scalar)
        ev_scan[slot] = torch.cuda.Event()
        ev_scan[slot].record(cur)

    def download(b):
        lo, hi = edges[b], edges[b + 1]
        mb = hi - lo
        slot = b % nslots
        oc, ox = outs(b)
        for a in range(lo, hi, stg):
            m = min(stg, hi - a)
            r = counters["out_chunk"] % 2
            counters["out_chunk"] += 1
            stage_free[r].acquire()              # the chunk that last used the staging buffers is on the host
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_scan[slot])
                for dev_o, stage, host_o in ((oc, oco_stage, out_co), (ox, ocr_stage, out_cr)):
                    if dev_o is None:
                        continue
                    k = host_o.shape[0]
                    src = dev_o.reshape(k, mb)[:, a - lo:a - lo + m]
                    dst = torch.from_numpy(stage[r][:k * m]).reshape(k, m)
                    dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_out)
            jobs.put((ev, r, a, m))
        ev_d2h[slot] = ev

    worker = threading.Thread(target=drain, name="xs-d2h", daemon=True)
    worker.start()
    try:
        invert(0, upload(0))
        nxt = upload(1) if nb > 1 else None
        for b in range(nb):
            if b + 1 < nb:
                invert(b + 1, nxt)                # queued behind the scan of block b
            if b + 2 < nb:
                nxt = upload(b + 2)               # host-side staging overlaps the scans
            download(b)                           # blocks on the staging ring; the GPU is busy with block b + 1
    finally:
        jobs.put(None)
        worker.join()
        s_in.synchronize()
        s_out.synchronize()
        cur.synchronize()
        for buf in taken:
            _POOL.give(buf)
    if errors:
        raise errors[0]
    return finish(out_co), finish(out_cr)


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
    return _invert(inc, sigma0, sigma0_dual, ancillary_wind, dsig_co, dsig_cr, model, kwargs)


def invert_to_speed_dir(inc, sigma0, sigma0_dual=None, /, ancillary_wind=None, dsig_co=0.1, dsig_cr=0.1, model=None,
                        ground_heading=None, dtype=np.float64, **kwargs):
    """`invert_from_model` with the post-processing every caller applies to its result fused into the inversion
    kernels (docs/examples/windspeed_retrieval_L1.ipynb cell 33, detrend.py:114-130):

        windspeed = np.abs(wind)
        winddir   = np.angle(wind, deg=True)                          # ground_heading is None: antenna convention
        winddir   = (90 - np.angle(wind, deg=True) + ground_heading) % 360   # meteorological convention otherwise

    The device writes speed / direction planes (float64, or float32 with dtype=np.float32) instead of complex128, which
    halves (quarters) the device->host copy.  Same arguments as `invert_from_model` plus `ground_heading` (degrees;
    scalar or raster) and `dtype`.  Returns (windspeed, winddir) for a co-pol model, the wind speed for a cross-pol
    model (as `invert_from_model`), ((windspeed_co, winddir_co), (windspeed_dual, winddir_dual)) for dual-pol.
    """
    if np.dtype(dtype) not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("dtype must be float32 or float64")
    return _invert(inc, sigma0, sigma0_dual, ancillary_wind, dsig_co, dsig_cr, model, kwargs, speed_dir=True,
                   ground_heading=ground_heading, out_f32=np.dtype(dtype) == np.dtype(np.float32))


def _invert(inc, sigma0, sigma0_dual, ancillary_wind, dsig_co, dsig_cr, model, kwargs, speed_dir=False,
            ground_heading=None, out_f32=False):
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
    epi = dict(speed_dir=speed_dir, ground_heading=ground_heading, out_f32=out_f32)

    if any(_xr.is_dask(v) for v in (inc, sigma0_co, sigma0_cr, ancillary_wind, dsig_cr) if v is not None):
        ws_co, ws_cr_or_dual = _invert_dask(models, dsig_co, kwargs, inc, sigma0_co, sigma0_cr, dsig_cr, ancillary_wind,
                                            template)
        if cross_only:
            ws_cr_or_dual = abs(ws_cr_or_dual)
        elif dual:
            import xarray as xr

            ws_cr_or_dual = xr.where((abs(ws_co) < 5) | (abs(ws_cr_or_dual) < 5), ws_co, ws_cr_or_dual)
        if speed_dir:  # lazy containers: the callers' own expressions, evaluated block-wise by dask
            def planes(z):
                ang = np.angle(z, deg=True)
                d = ang if ground_heading is None else (90 - ang + ground_heading) % 360
                sp = abs(z)
                return (sp.astype(np.float32), d.astype(np.float32)) if out_f32 else (sp, d)

            ws_co = None if cross_only else planes(ws_co)
            ws_cr_or_dual = ws_cr_or_dual if cross_only else planes(ws_cr_or_dual)
    elif _is_tensor(inc):
        plan = _get_plan(models[0], models[1] if sigma0_cr is not None else None, dsig_co, kwargs)
        ws_co, ws_cr_or_dual = _run_resident(plan, inc, sigma0_co, sigma0_cr, dsig_cr, ancillary_wind if anc_given else None,
                                             merge_dual=dual, cr_abs=cross_only, **epi)
        if speed_dir:
            ws_co = None if ws_co is None else (ws_co[0], ws_co[1])
            ws_cr_or_dual = ws_cr_or_dual if cross_only else (ws_cr_or_dual[0], ws_cr_or_dual[1])
    else:
        plan = _get_plan(models[0], models[1] if sigma0_cr is not None else None, dsig_co, kwargs)
        dsig_in = dsig_cr if np.isscalar(dsig_cr) else _values(dsig_cr)
        gh_in = ground_heading if (ground_heading is None or np.isscalar(ground_heading)) else _values(ground_heading)
        epi["ground_heading"] = gh_in
        oc, ox = _run_device(plan, _values(inc), _values(sigma0_co), _values(sigma0_cr), dsig_in,
                             _values(ancillary_wind) if anc_given else None, sigma0_db=False, merge_dual=dual,
                             cr_abs=cross_only, **epi)
        lab = _xr.is_labelled(template)
        wrap = (lambda a, name: _xr.like(template, a, name=name)) if lab else (lambda a, name: a)
        if speed_dir:
            ws_co = None if oc is None else (wrap(oc[0], "windspeed"), wrap(oc[1], "winddir"))
            ws_cr_or_dual = wrap(ox, "windspeed") if cross_only else (wrap(ox[0], "windspeed"), wrap(ox[1], "winddir"))
        else:
            ws_co = None if oc is None else wrap(oc, "windspeed_gmf")
            ws_cr_or_dual = wrap(ox, "windspeed_gmf")

    # attrs and returns, windspeed.py:395-439
    def set_attrs(obj, **attrs):
        for o in (obj if isinstance(obj, tuple) else (obj,)):
            if hasattr(o, "attrs"):
                o.attrs.update(attrs)

    if models[0] is not None and models[0].iscopol:
        set_attrs(ws_co, comment=f"wind speed and direction inverted from model {models[0].name} ({models[0].pol})",
                  model=models[0].name)
    if not dual:
        if not cross_only:
            return ws_co
        set_attrs(ws_cr_or_dual, comment=f"wind speed inverted from model {models[1].name} ({models[1].pol})",
                  model=models[1].name, units="m/s")
        return ws_cr_or_dual
    set_attrs(ws_cr_or_dual, comment=(f"wind speed and direction inverted from model {models[0].name} "
                                      f"({models[0].pol}) and {models[1].name} ({models[1].pol})"),
              model=f"{models[0].name} {models[1].name}")
    return ws_co, ws_cr_or_dual


def _invert_dask(models, dsig_co, kwargs, inc, sigma0_co, sigma0_cr, dsig_cr, ancillary_wind, template):
    """dask-backed inputs: lazy `da.apply_gufunc` over blocks with the sample axis as core dimension, like
    windspeed.py:356-364; every block goes through the numpy-level operator (and hence the GPU).  dask and xarray are
    absent from the build image: tests/test_dask_stub.py executes this path with minimal stand-ins for both."""
    import dask.array as da
    import xarray as xr

    if not _xr.is_labelled(template):  # the result mirrors the first labelled input
        template = next(v for v in (sigma0_co, sigma0_cr, inc, ancillary_wind, dsig_cr) if v is not None and _xr.is_labelled(v))
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
