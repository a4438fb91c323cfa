"""CPU ORACLE -- test infrastructure, not product code.

Restates the reference's wind-inversion hot path on the CPU so that the CUDA path can be checked
against it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; xsarsea_b200/ never does.

  oracle/c/xs_oracle.c   plain-C restatement (GMF formulas, LUT build, scipy-style linear interpolation,
                         dB conversion, per-pixel inversion incl. NaN semantics, detrend)
  oracle/numba_port.py   the reference's own program shape (per-pixel whole-array numpy expressions
                         compiled with the identical numba.guvectorize arguments) -- the timed CPU baseline
  oracle/lut.py          host-side LUT recipes (np.linspace grids, resolution decision table)

Pinning: tests/golden/ holds outputs of the reference's real numba kernels (made by
tests/golden/make_golden.py where /root/reference is mounted).  The interpolation sub-step follows
scipy.interpolate.interp1d because xarray cannot be imported here: parity unpinned by the reference for
that sub-step.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libxs_oracle.so")

MODEL_IDS = {
    "gmf_cmod5": 0,
    "gmf_cmod5n": 1,
    "gmf_cmod5n_pr_zhangA": 2,
    "gmf_cmod5n_pr_mouche1": 3,
    "gmf_cmodifr2": 4,
    "gmf_rs2_v2": 5,
    "gmf_s1_v2": 6,
    "gmf_rcm_noaa": 7,
    "gmf_s1_v3_ew_rec": 8,
    "gmf_rs2_v3": 9,
    "gmf_rcm_v3": 10,
    "gmf_rcm_v4": 11,
    "gmf_rs2_v4": 12,
}
COPOL_MODELS = [n for n, i in MODEL_IDS.items() if i <= 4]
CROSSPOL_MODELS = [n for n, i in MODEL_IDS.items() if i >= 5]

_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/c/xs_oracle.c with gcc (recipe: oracle/c/Makefile)."""
    src = os.path.join(_HERE, "c", "xs_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(_HERE, "c")], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        d, i32, i64, p = ctypes.c_double, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p
        L.xso_gmf_scalar.restype = d
        L.xso_gmf_scalar.argtypes = [i32, d, d, d]
        L.xso_gmf_eval.restype = None
        L.xso_gmf_eval.argtypes = [i32, p, p, p, p, i64]
        L.xso_lut_build.restype = None
        L.xso_lut_build.argtypes = [i32, p, i32, p, i32, p, i32, p]
        L.xso_interp_axis.restype = i32
        L.xso_interp_axis.argtypes = [p, i64, i32, i64, p, p, i32, p]
        L.xso_to_db.restype = None
        L.xso_to_db.argtypes = [p, p, i64]
        L.xso_to_linear.restype = None
        L.xso_to_linear.argtypes = [p, p, i64]
        L.xso_invert.restype = None
        L.xso_invert.argtypes = [p, p, i32, p, i32, p, i32, p, p, i32, p, p, i32, p, i32, d] + [p] * 5 + [i64] + [p] * 4
        L.xso_detrend.restype = None
        L.xso_detrend.argtypes = [p, p, i64, i64, p]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def gmf_scalar(model: str, inc: float, wspd: float, phi: float = float("nan")) -> float:
    return lib().xso_gmf_scalar(MODEL_IDS[model], float(inc), float(wspd), float(phi))


def gmf_eval(model: str, inc, wspd, phi=None) -> np.ndarray:
    """Element-wise GMF on broadcast inputs (reference K3, gmfs.py:210-214)."""
    if phi is None:
        inc, wspd = np.broadcast_arrays(inc, wspd)
        ph = None
    else:
        inc, wspd, ph = np.broadcast_arrays(inc, wspd, phi)
        ph = _f64(ph)
    shape = inc.shape
    inc, wspd = _f64(inc), _f64(wspd)
    out = np.empty(inc.size, dtype=np.float64)
    lib().xso_gmf_eval(MODEL_IDS[model], _ptr(inc), _ptr(wspd), _ptr(ph), _ptr(out), inc.size)
    return out.reshape(shape)


def lut_build(model: str, inc, wspd, phi=None) -> np.ndarray:
    """Outer-product LUT [inc][wspd][phi] (or [inc][wspd] for cross-pol); reference K2, gmfs.py:218-230."""
    inc, wspd = _f64(inc), _f64(wspd)
    ph = None if phi is None else _f64(phi)
    shape = (inc.size, wspd.size) + (() if ph is None else (ph.size,))
    out = np.empty(shape, dtype=np.float64)
    lib().xso_lut_build(MODEL_IDS[model], _ptr(inc), inc.size, _ptr(wspd), wspd.size, _ptr(ph),
                        0 if ph is None else ph.size, _ptr(out))
    return out


def interp_axis(src: np.ndarray, axis: int, x_src, x_dst) -> np.ndarray:
    """scipy interp1d(kind='linear', bounds_error=True) along `axis` (reference K5 sub-step, models.py:167)."""
    src = _f64(src)
    x_src, x_dst = _f64(x_src), _f64(x_dst)
    outer = int(np.prod(src.shape[:axis], dtype=np.int64))
    inner = int(np.prod(src.shape[axis + 1:], dtype=np.int64))
    dst = np.empty(src.shape[:axis] + (x_dst.size,) + src.shape[axis + 1:], dtype=np.float64)
    rc = lib().xso_interp_axis(_ptr(src), outer, src.shape[axis], inner, _ptr(x_src), _ptr(x_dst), x_dst.size, _ptr(dst))
    if rc != 0:
        raise ValueError("A value in x_new is outside the interpolation range.")  # scipy's bounds_error text
    return dst


def to_db(lut: np.ndarray) -> np.ndarray:
    lut = _f64(lut)
    out = np.empty_like(lut)
    lib().xso_to_db(_ptr(lut), _ptr(out), lut.size)
    return out


def to_linear(lut: np.ndarray) -> np.ndarray:
    lut = _f64(lut)
    out = np.empty_like(lut)
    lib().xso_to_linear(_ptr(lut), _ptr(out), lut.size)
    return out


def wind_tables(wspd_grid, phi_grid):
    """np_wspd_lut_co_antenna / _azi (windspeed.py:166-168) and cos/sin(deg2rad(phi)) with numpy's rounding."""
    phi_lut, wspd_lut = np.meshgrid(_f64(phi_grid), _f64(wspd_grid))
    return (np.ascontiguousarray(wspd_lut * np.cos(np.radians(phi_lut))),
            np.ascontiguousarray(wspd_lut * np.sin(np.radians(phi_lut))))


def phi_is_180(phi_grid) -> bool:
    """windspeed.py:152-156"""
    phi_grid = np.asarray(phi_grid)
    return bool((180 - (phi_grid[-1] - phi_grid[0])) < 2)


def invert(inc, s0_co_db, s0_cr_db, dsig_cr, anc, *, co_lut=None, inc_grid=None, wspd_grid=None, phi_grid=None,
           cr_lut=None, inc_cr_grid=None, wspd_cr_grid=None, dsig_co=0.1, threads=None):
    """Reference K1 (windspeed.py:183-282) on dB inputs.

    co_lut [n_inc][n_wspd][n_phi] dB, cr_lut [n_inc_cr][n_wspd_cr] dB (model-native layouts).
    Returns (wind_co c128, wind_dual c128, idx_co i32, idx_cr i32) shaped like `inc`.
    """
    inc = _f64(inc)
    shape = inc.shape
    n = inc.size
    s0_co_db = _f64(np.broadcast_to(s0_co_db, shape)).ravel()
    s0_cr_db = _f64(np.broadcast_to(s0_cr_db, shape)).ravel()
    dsig_cr = _f64(np.broadcast_to(dsig_cr, shape)).ravel()
    anc = np.ascontiguousarray(np.broadcast_to(anc, shape), dtype=np.complex128).ravel()
    incf = inc.ravel()
    if co_lut is not None:
        co_lut = _f64(co_lut)
        inc_grid, wspd_grid, phi_grid = _f64(inc_grid), _f64(wspd_grid), _f64(phi_grid)
        assert co_lut.shape == (inc_grid.size, wspd_grid.size, phi_grid.size)
        wcos, wsin = wind_tables(wspd_grid, phi_grid)
        p180 = phi_is_180(phi_grid)
        n_inc, n_wspd, n_phi = co_lut.shape
    else:
        inc_grid = wspd_grid = phi_grid = wcos = wsin = None
        p180, n_inc, n_wspd, n_phi = False, 0, 0, 0
    if cr_lut is not None:
        cr_lut = _f64(cr_lut)
        inc_cr_grid, wspd_cr_grid = _f64(inc_cr_grid), _f64(wspd_cr_grid)
        assert cr_lut.shape == (inc_cr_grid.size, wspd_cr_grid.size)
        n_inc_cr, n_wspd_cr = cr_lut.shape
    else:
        inc_cr_grid = wspd_cr_grid = None
        n_inc_cr, n_wspd_cr = 0, 0
    out_co = np.empty(n, dtype=np.complex128)
    out_cr = np.empty(n, dtype=np.complex128)
    idx_co = np.empty(n, dtype=np.int32)
    idx_cr = np.empty(n, dtype=np.int32)
    L = lib()

    def run(lo, hi):
        def off(a, itemsize):
            return ctypes.c_void_p(a.ctypes.data + lo * itemsize)

        L.xso_invert(_ptr(co_lut), _ptr(inc_grid), n_inc, _ptr(wspd_grid), n_wspd, _ptr(phi_grid), n_phi,
                     _ptr(wcos), _ptr(wsin), int(p180), _ptr(cr_lut), _ptr(inc_cr_grid), n_inc_cr,
                     _ptr(wspd_cr_grid), n_wspd_cr, float(dsig_co), off(incf, 8), off(s0_co_db, 8),
                     off(s0_cr_db, 8), off(dsig_cr, 8), off(anc, 16), hi - lo, off(out_co, 16), off(out_cr, 16),
                     off(idx_co, 4), off(idx_cr, 4))

    threads = threads or min(os.cpu_count() or 1, 32)
    if n < 4 * threads or threads == 1:
        run(0, n)
    else:
        edges = np.linspace(0, n, threads * 4 + 1).astype(np.int64)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda k: run(int(edges[k]), int(edges[k + 1])), range(threads * 4)))
    return out_co.reshape(shape), out_cr.reshape(shape), idx_co.reshape(shape), idx_cr.reshape(shape)


def detrend(sigma0, gmf_line):
    """detrend.py:63-64 given the GMF profile of the first line."""
    sigma0 = _f64(sigma0)
    H, W = sigma0.shape
    gmf_line = _f64(gmf_line)
    out = np.empty_like(sigma0)
    lib().xso_detrend(_ptr(sigma0), _ptr(gmf_line), H, W, _ptr(out))
    return out
