"""Device-side operators of the hot path: thin Python over the C ABI (xsarsea_b200/_native.py).

Everything here takes/returns torch CUDA tensors (float64 unless stated); the host only computes the small
grid vectors (np.linspace, np.cos/np.sin of the phi grid) so that they carry numpy's rounding, exactly as
the reference does.
"""
from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _native as nat


def _t():
    return nat.torch_cuda()


def to_device(a, dtype=None):
    torch = _t()
    if isinstance(a, torch.Tensor):
        t = a.cuda()
        return t.to(dtype).contiguous() if dtype is not None else t.contiguous()
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a).cuda()
    return t.to(dtype) if dtype is not None else t


def gmf_eval(model_id: int, inc, wspd, phi=None):
    """Element-wise GMF over already-broadcast device tensors (reference K3, gmfs.py:210-214).
    inc/wspd float64 or float32 (both the same); phi float64 or None (cross-pol)."""
    torch = _t()
    L = nat.load()
    assert inc.shape == wspd.shape and inc.dtype == wspd.dtype
    dt = nat.XS_F64 if inc.dtype == torch.float64 else nat.XS_F32
    inc, wspd = inc.contiguous(), wspd.contiguous()
    if phi is not None:
        phi = phi.to(torch.float64).contiguous()
        assert phi.shape == inc.shape
    out = torch.empty_like(inc)
    nat.check(L.xs_gmf_eval(model_id, dt, nat.dptr(inc), nat.dptr(wspd), nat.dptr(phi), nat.dptr(out), inc.numel(),
                            nat.stream_ptr()), "xs_gmf_eval")
    return out


def lut_build(model_id: int, inc_grid, wspd_grid, phi_grid=None):
    """Outer-product LUT [n_inc, n_wspd(, n_phi)] float64 on device (reference K2, gmfs.py:218-230)."""
    torch = _t()
    L = nat.load()
    gi, gw = nat.host_f64(inc_grid), nat.host_f64(wspd_grid)
    gp = None if phi_grid is None else nat.host_f64(phi_grid)
    shape = (gi.size, gw.size) + (() if gp is None else (gp.size,))
    out = torch.empty(shape, dtype=torch.float64, device="cuda")
    nat.check(L.xs_lut_build(model_id, nat.hptr(gi), gi.size, nat.hptr(gw), gw.size, nat.hptr(gp),
                             0 if gp is None else gp.size, nat.dptr(out), nat.stream_ptr()), "xs_lut_build")
    return out


def lut_interp_axis(lut, axis: int, x_src, x_dst):
    """scipy interp1d(kind='linear', bounds_error=True) along `axis` on device (reference K5, models.py:167)."""
    torch = _t()
    L = nat.load()
    lut = lut.contiguous()
    xs, xd = nat.host_f64(x_src), nat.host_f64(x_dst)
    assert lut.shape[axis] == xs.size
    outer = int(np.prod(lut.shape[:axis], dtype=np.int64))
    inner = int(np.prod(lut.shape[axis + 1:], dtype=np.int64))
    out = torch.empty(tuple(lut.shape[:axis]) + (xd.size,) + tuple(lut.shape[axis + 1:]), dtype=torch.float64,
                      device="cuda")
    nat.check(L.xs_lut_interp_axis(nat.dptr(lut), outer, xs.size, inner, nat.hptr(xs), nat.hptr(xd), xd.size,
                                   nat.dptr(out), nat.stream_ptr()), "xs_lut_interp_axis")
    return out


def lut_to_db(lut):
    torch = _t()
    out = torch.empty_like(lut)
    nat.check(nat.load().xs_lut_to_db(nat.dptr(lut.contiguous()), nat.dptr(out), lut.numel(), nat.stream_ptr()),
              "xs_lut_to_db")
    return out


def lut_to_linear(lut):
    torch = _t()
    out = torch.empty_like(lut)
    nat.check(nat.load().xs_lut_to_linear(nat.dptr(lut.contiguous()), nat.dptr(out), lut.numel(), nat.stream_ptr()),
              "xs_lut_to_linear")
    return out


def detrend(sigma0, gmf_line):
    """out[l, s] = sigma0[l, s] / (gmf_line[s] / nanmean(gmf_line)) (reference detrend.py:63-64)."""
    torch = _t()
    sigma0 = sigma0.contiguous()
    assert sigma0.dim() == 2 and gmf_line.numel() == sigma0.shape[1]
    dt = nat.XS_F64 if sigma0.dtype == torch.float64 else nat.XS_F32
    gmf_line = gmf_line.to(torch.float64).contiguous()
    out = torch.empty_like(sigma0)
    nat.check(nat.load().xs_detrend(nat.dptr(sigma0), nat.dptr(gmf_line), sigma0.shape[0], sigma0.shape[1], dt,
                                    nat.dptr(out), nat.stream_ptr()), "xs_detrend")
    return out


def _same_real_dtype(*ts):
    """Common raster dtype code of device tensors: float32 only if all are float32, else everything is made float64."""
    torch = _t()
    if all(t.dtype == torch.float32 for t in ts):
        return nat.XS_F32, [t.contiguous() for t in ts]
    return nat.XS_F64, [t.to(torch.float64).contiguous() for t in ts]


def dsig(dsig_id: int, inc, sigma0_cr, nesz_cr):
    """dsig_cr raster on device (reference get_dsig, windspeed/utils.py:47-91); inputs already broadcast, float64 out."""
    torch = _t()
    ts = [sigma0_cr, nesz_cr] + ([inc] if inc is not None else [])
    assert all(t.shape == sigma0_cr.shape for t in ts)
    dt, ts = _same_real_dtype(*ts)
    out = torch.empty(sigma0_cr.shape, dtype=torch.float64, device="cuda")
    nat.check(nat.load().xs_dsig(dsig_id, dt, nat.dptr(ts[2]) if inc is not None else None, nat.dptr(ts[0]),
                                 nat.dptr(ts[1]), nat.dptr(out), out.numel(), nat.stream_ptr()), "xs_dsig")
    return out


def dsig_wspd(dsig_wspd_id: int, u_crosspol, snr_cr):
    """Co/cross blending weight on device (reference get_dsig_wspd, windspeed/utils.py:18-44)."""
    torch = _t()
    assert u_crosspol.shape == snr_cr.shape
    dt, (u, s) = _same_real_dtype(u_crosspol, snr_cr)
    out = torch.empty(u.shape, dtype=torch.float64, device="cuda")
    nat.check(nat.load().xs_dsig_wspd(dsig_wspd_id, dt, nat.dptr(u), nat.dptr(s), nat.dptr(out), out.numel(),
                                      nat.stream_ptr()), "xs_dsig_wspd")
    return out


def nesz_flatten(noise, inc):
    """Per-line order-1 flattening of the noise in dB (reference nesz_flattening, windspeed/utils.py:94-163)."""
    torch = _t()
    L = nat.load()
    assert noise.dim() == 2 and inc.shape == noise.shape
    dt, (noise, inc) = _same_real_dtype(noise, inc)
    h, w = noise.shape
    out = torch.empty((h, w), dtype=torch.float64, device="cuda")
    if h == 0 or w == 0:
        return out
    need = int(L.xs_nesz_flatten_workspace_bytes(h, w))
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    nat.check(L.xs_nesz_flatten(nat.dptr(noise), nat.dptr(inc), h, w, dt, nat.dptr(out), nat.dptr(ws), need,
                                nat.stream_ptr()), "xs_nesz_flatten")
    return out


def local_gradients(image):
    """(G2 complex128, G3 float64, c float64), each [H//2, W//2], of a 2-D device image (reference local_gradients,
    gradients.py:588-634)."""
    torch = _t()
    L = nat.load()
    assert image.dim() == 2
    dt, (image,) = _same_real_dtype(image)
    h, w = image.shape
    h2, w2 = h // 2, w // 2
    g2 = torch.empty((h2, w2), dtype=torch.complex128, device="cuda")
    g3 = torch.empty((h2, w2), dtype=torch.float64, device="cuda")
    c = torch.empty((h2, w2), dtype=torch.float64, device="cuda")
    if h2 == 0 or w2 == 0:
        return g2, g3, c
    need = int(L.xs_local_gradients_workspace_bytes(h, w))
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    nat.check(L.xs_local_gradients(nat.dptr(image), h, w, dt, nat.dptr(g2), nat.dptr(g3), nat.dptr(c), nat.dptr(ws), need,
                                   nat.stream_ptr()), "xs_local_gradients")
    return g2, g3, c


class ScanTimer:
    """An `xs_timer` (CUDA events recorded around the co-pol scan of one xs_invert call)."""

    def __init__(self):
        self._h = ctypes.c_void_p()
        nat.check(nat.load().xs_timer_create(ctypes.byref(self._h)), "xs_timer_create")

    def elapsed_ms(self):
        """(k_scan_co, k_refine_easy) device times in ms; waits for the kernels to finish."""
        ms = (ctypes.c_float * 2)()
        nat.check(nat.load().xs_timer_elapsed_ms(self._h, ms), "xs_timer_elapsed_ms")
        return float(ms[0]), float(ms[1])

    def __del__(self):
        try:
            if self._h:
                nat.load().xs_timer_destroy(self._h)
                self._h = None
        except Exception:
            pass


class InversionPlan:
    """Owns an xs_plan: the device LUTs of one (co-pol, cross-pol) model pair plus the scan image.

    co = (lut_db [n_inc, n_wspd, n_phi] device f64, inc_grid, wspd_grid, phi_grid) or None
    cr = (lut_db [n_inc_cr, n_wspd_cr] device f64, inc_grid, wspd_grid) or None
    Mirrors the per-call setup of windspeed.py:139-181 (done once here and cached by the caller).

    Re-entrant (dask's threaded scheduler calls the operator concurrently): the native plan is
    immutable after creation; every `invert` call allocates its own workspace, counters and (optionally) timer, and the
    "last call" bookkeeping is per host thread.  `close()` is deferred while calls are in flight.
    """

    def __init__(self, co=None, cr=None, dsig_co=0.1):
        torch = _t()
        L = nat.load()
        self._handle = None
        self._lock = threading.Lock()
        self._users = 0
        self._close_pending = False
        self._tls = threading.local()
        d = nat.PlanDesc()
        keep = []
        self.co_grids = self.cr_grids = None
        if co is not None:
            lut, gi, gw, gp = co
            lut = lut.to(torch.float64).contiguous()
            gi, gw, gp = nat.host_f64(gi), nat.host_f64(gw), nat.host_f64(gp)
            assert tuple(lut.shape) == (gi.size, gw.size, gp.size), (tuple(lut.shape), gi.size, gw.size, gp.size)
            cphi, sphi = np.cos(np.radians(gp)), np.sin(np.radians(gp))  # windspeed.py:167-168
            keep += [lut, gi, gw, gp, cphi, sphi]
            d.co_lut_db_dev = lut.data_ptr()
            d.inc_grid_host, d.wspd_grid_host, d.phi_grid_host = gi.ctypes.data, gw.ctypes.data, gp.ctypes.data
            d.cos_phi_host, d.sin_phi_host = cphi.ctypes.data, sphi.ctypes.data
            d.n_inc, d.n_wspd, d.n_phi = gi.size, gw.size, gp.size
            self.co_lut = lut
            self.co_grids = (gi, gw, gp)
        if cr is not None:
            lut, gi, gw = cr
            lut = lut.to(torch.float64).contiguous()
            gi, gw = nat.host_f64(gi), nat.host_f64(gw)
            assert tuple(lut.shape) == (gi.size, gw.size)
            keep += [lut, gi, gw]
            d.cr_lut_db_dev = lut.data_ptr()
            d.inc_cr_grid_host, d.wspd_cr_grid_host = gi.ctypes.data, gw.ctypes.data
            d.n_inc_cr, d.n_wspd_cr = gi.size, gw.size
            self.cr_lut = lut
            self.cr_grids = (gi, gw)
        d.dsig_co = float(dsig_co)
        self._keep = keep
        h = ctypes.c_void_p()
        nat.check(L.xs_plan_create(ctypes.byref(d), nat.stream_ptr(), ctypes.byref(h)), "xs_plan_create")
        self._handle = h

    def close(self):
        """Free the native plan -- after the calls in flight on other threads have returned."""
        with self._lock:
            if self._users > 0:
                self._close_pending = True
                return
            h, self._handle = self._handle, None
        if h is not None:
            nat.load().xs_plan_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _acquire(self):
        with self._lock:
            if self._handle is None or self._close_pending:
                raise nat.NativeError("InversionPlan is closed")
            self._users += 1
            return self._handle

    def _release(self):
        with self._lock:
            self._users -= 1
            last = self._users == 0 and self._close_pending
            if last:
                self._close_pending = False
        if last:
            self.close()

    def workspace_bytes(self, n_px: int, flags: int = 0) -> int:
        return int(nat.load().xs_invert_workspace_bytes(self._handle, int(n_px), int(flags)))

    def invert(self, inc, sigma0_co=None, sigma0_cr=None, dsig_cr=0.1, ancillary=None, *, sigma0_db=False,
               merge_dual=False, cr_abs=False, mode=nat.MODE_FAST, want_idx=False, out_co=None, out_cr=None,
               need_co=False, cr_full_scan=False, speed_dir=False, ground_heading=None, out_f32=False, timed=False,
               no_prune=False):
        """Run K1 on device tensors (all the same shape; float64/complex128 or float32/complex64).

        Returns (wind_co complex128 | None, wind_cr complex128 (float64 if cr_abs) | None, idx_co, idx_cr).
        speed_dir=True (row F2 epilogue): the winds are [2, *shape] planes (speed m/s, direction deg; float32 with
        out_f32) instead of complex128 -- direction = np.angle(wind, deg=True), or (90 - angle + ground_heading) % 360
        when `ground_heading` (device raster or scalar, degrees) is given.
        """
        torch = _t()
        L = nat.load()
        inc = inc.contiguous()
        shape, n = inc.shape, inc.numel()
        f32 = inc.dtype == torch.float32
        rdt, cdt = (torch.float32, torch.complex64) if f32 else (torch.float64, torch.complex128)

        def prep(x, dt):
            if x is None:
                return None
            x = x.contiguous()
            if x.dtype != dt or x.shape != shape:
                raise TypeError(f"raster of dtype {x.dtype}/shape {tuple(x.shape)}; expected {dt}/{tuple(shape)}")
            return x

        s_co, s_cr, anc = prep(sigma0_co, rdt), prep(sigma0_cr, rdt), prep(ancillary, cdt)
        a = nat.InvertArgs()
        a.inc, a.sigma0_co, a.sigma0_cr = inc.data_ptr(), nat.dptr(s_co), nat.dptr(s_cr)
        a.ancillary = nat.dptr(anc)
        dsig_t = gh_t = None
        if hasattr(dsig_cr, "shape") and getattr(dsig_cr, "ndim", 0) > 0:
            dsig_t = prep(dsig_cr, rdt)
            a.dsig_cr = dsig_t.data_ptr()
        else:
            a.dsig_cr_scalar = float(dsig_cr)
        a.dtype = nat.XS_F32 if f32 else nat.XS_F64
        flags = (nat.FLAG_SIGMA0_DB if sigma0_db else 0) | (nat.FLAG_MERGE_DUAL if merge_dual else 0) | (
            nat.FLAG_CR_ABS if cr_abs else 0) | (nat.FLAG_CR_FULL_SCAN if cr_full_scan else 0) | (
            nat.FLAG_NO_PRUNE if no_prune else 0)
        if speed_dir:
            flags |= nat.FLAG_OUT_SPEED_DIR | (nat.FLAG_OUT_F32 if out_f32 else 0)
            if ground_heading is not None:
                flags |= nat.FLAG_DIR_METEO
                if hasattr(ground_heading, "shape") and getattr(ground_heading, "ndim", 0) > 0:
                    gh_t = prep(ground_heading, rdt)
                    a.ground_heading = gh_t.data_ptr()
                else:
                    a.ground_heading_scalar = float(ground_heading)
        elif ground_heading is not None or out_f32:
            raise ValueError("ground_heading / out_f32 need speed_dir=True")
        a.flags = flags
        a.mode = mode
        a.n_px = n
        has_co = self.co_grids is not None and s_co is not None
        pdt = torch.float32 if out_f32 else torch.float64
        if out_co is None and (has_co or need_co):
            out_co = torch.empty((2,) + tuple(shape), dtype=pdt, device="cuda") if speed_dir else torch.empty(
                shape, dtype=torch.complex128, device="cuda")
        if out_cr is None:
            if cr_abs:
                out_cr = torch.empty(shape, dtype=torch.float64, device="cuda")
            elif speed_dir:
                out_cr = torch.empty((2,) + tuple(shape), dtype=pdt, device="cuda")
            else:
                out_cr = torch.empty(shape, dtype=torch.complex128, device="cuda")
        a.out_co, a.out_cr = nat.dptr(out_co), nat.dptr(out_cr)
        idx_co = idx_cr = None
        if want_idx:
            idx_co = torch.empty(shape, dtype=torch.int32, device="cuda")
            idx_cr = torch.empty(shape, dtype=torch.int32, device="cuda")
            a.idx_co, a.idx_cr = idx_co.data_ptr(), idx_cr.data_ptr()
        handle = self._acquire()
        try:
            # everything mutable belongs to this call: workspace (stream-ordered allocation on the current stream),
            # counters, timer
            need = int(L.xs_invert_workspace_bytes(handle, n, flags))
            workspace = torch.empty(max(need, 256), dtype=torch.uint8, device="cuda")
            counters = torch.zeros(nat.N_COUNTERS, dtype=torch.int64, device="cuda")
            timer = ScanTimer() if timed else None
            a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel()
            a.counters_dev = counters.data_ptr()
            a.scan_timer = timer._h if timer is not None else None
            nat.check(L.xs_invert(handle, ctypes.byref(a), nat.stream_ptr()), "xs_invert")
        finally:
            self._release()
        self._tls.counters, self._tls.timer = counters, timer
        return out_co, out_cr, idx_co, idx_cr

    def last_scan_ms(self):
        """Device times (k_scan_co, k_refine_easy) in ms of this thread's last `invert(..., timed=True)`."""
        timer = getattr(self._tls, "timer", None)
        if timer is None:
            raise nat.NativeError("no timed invert() call on this thread")
        return timer.elapsed_ms()

    def debug_counters(self):
        """Raw device counters of this thread's last invert() (layout: include/xsarsea_b200.h)."""
        c = getattr(self._tls, "counters", None)
        if c is None:
            raise nat.NativeError("no invert() call on this thread")
        return [int(v) for v in c.cpu().tolist()]

    def last_stats(self):
        c = self.debug_counters()
        return dict(scan_pixels=c[2], fp64_chunks=c[3], exhaustive_pixels=c[1], tiles=c[0], fp64_pixels=c[11], many_lane_pixels=c[12],
                    shared_mode_positions=c[13], cross_listed_pixels=c[4], chunks_streamed=c[5], warp_chunk_phi=c[6])
