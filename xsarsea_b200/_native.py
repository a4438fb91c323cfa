"""ctypes binding of libxsarsea_b200.so (the C ABI declared in include/xsarsea_b200.h).

This module is the only place the package touches native code.  There is no CPU fallback: if the shared
library is missing, or no CUDA device is present, calls raise -- loudly.
torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libxsarsea_b200.so")
CSRC = os.path.join(_HERE, "csrc")

XS_F64, XS_F32 = 0, 1
FLAG_SIGMA0_DB, FLAG_MERGE_DUAL, FLAG_CR_ABS, FLAG_CR_FULL_SCAN = 1, 2, 4, 8
FLAG_OUT_SPEED_DIR, FLAG_DIR_METEO, FLAG_OUT_F32, FLAG_NO_PRUNE = 16, 32, 64, 128
ABI_VERSION = 2
N_COUNTERS = 16
MODE_FAST, MODE_FP64 = 0, 1

GMF_IDS = {
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

EXPORTS = [
    "xs_abi_version", "xs_last_error", "xs_launch_count", "xs_bench_fp32_peak", "xs_gmf_eval", "xs_lut_build", "xs_lut_interp_axis",
    "xs_lut_to_db", "xs_lut_to_linear", "xs_plan_create", "xs_plan_destroy", "xs_invert_workspace_bytes",
    "xs_invert", "xs_timer_create", "xs_timer_destroy", "xs_timer_elapsed_ms", "xs_detrend",
    "xs_dsig", "xs_dsig_wspd", "xs_nesz_flatten_workspace_bytes", "xs_nesz_flatten",
    "xs_local_gradients_workspace_bytes", "xs_local_gradients",
]

DSIG_IDS = {"gmf_s1_v2": 0, "gmf_rs2_v2": 1, "sarwing_lut_cmodms1ahw": 2, "nc_lut_cmodms1ahw": 2}
DSIG_WSPD_IDS = {"dsig_wspd_rs2_v3": 0, "dsig_wspd_s1_ew_rec_v3": 1, "dsig_wspd_rcm_v3": 2}


class NativeError(RuntimeError):
    pass


class PlanDesc(ctypes.Structure):
    _fields_ = [
        ("co_lut_db_dev", ctypes.c_void_p),
        ("inc_grid_host", ctypes.c_void_p),
        ("wspd_grid_host", ctypes.c_void_p),
        ("phi_grid_host", ctypes.c_void_p),
        ("cos_phi_host", ctypes.c_void_p),
        ("sin_phi_host", ctypes.c_void_p),
        ("n_inc", ctypes.c_int32),
        ("n_wspd", ctypes.c_int32),
        ("n_phi", ctypes.c_int32),
        ("cr_lut_db_dev", ctypes.c_void_p),
        ("inc_cr_grid_host", ctypes.c_void_p),
        ("wspd_cr_grid_host", ctypes.c_void_p),
        ("n_inc_cr", ctypes.c_int32),
        ("n_wspd_cr", ctypes.c_int32),
        ("dsig_co", ctypes.c_double),
    ]


class InvertArgs(ctypes.Structure):
    _fields_ = [
        ("inc", ctypes.c_void_p),
        ("sigma0_co", ctypes.c_void_p),
        ("sigma0_cr", ctypes.c_void_p),
        ("dsig_cr", ctypes.c_void_p),
        ("ancillary", ctypes.c_void_p),
        ("dsig_cr_scalar", ctypes.c_double),
        ("dtype", ctypes.c_int32),
        ("flags", ctypes.c_uint32),
        ("mode", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("n_px", ctypes.c_int64),
        ("out_co", ctypes.c_void_p),
        ("out_cr", ctypes.c_void_p),
        ("idx_co", ctypes.c_void_p),
        ("idx_cr", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p),
        ("workspace_bytes", ctypes.c_size_t),
        ("counters_dev", ctypes.c_void_p),
        ("scan_timer", ctypes.c_void_p),
        ("ground_heading", ctypes.c_void_p),
        ("ground_heading_scalar", ctypes.c_double),
    ]


_lib = None
_lock = threading.Lock()


def build(force: bool = False) -> str:
    """Compile the CUDA sources for sm_100a with nvcc (xsarsea_b200/csrc/Makefile) if the library is stale."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "xsarsea_b200.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        r = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
        if r.returncode != 0:
            raise NativeError("building libxsarsea_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


def load():
    """Load the shared library (no CUDA call is made here, so this works on a box without a GPU)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. Run `python -c 'import "
                f"__graft_entry__ as g; g.build()'` (or `make -C xsarsea_b200/csrc`). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        i32, i64, dbl, vp, sz = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t
        L.xs_abi_version.restype = i32
        L.xs_abi_version.argtypes = []
        L.xs_last_error.restype = ctypes.c_char_p
        L.xs_last_error.argtypes = []
        L.xs_launch_count.restype = i64
        L.xs_launch_count.argtypes = []
        L.xs_bench_fp32_peak.restype = i32
        L.xs_bench_fp32_peak.argtypes = [ctypes.POINTER(dbl), vp]
        L.xs_gmf_eval.restype = i32
        L.xs_gmf_eval.argtypes = [i32, i32, vp, vp, vp, vp, i64, vp]
        L.xs_lut_build.restype = i32
        L.xs_lut_build.argtypes = [i32, vp, i32, vp, i32, vp, i32, vp, vp]
        L.xs_lut_interp_axis.restype = i32
        L.xs_lut_interp_axis.argtypes = [vp, i64, i32, i64, vp, vp, i32, vp, vp]
        L.xs_lut_to_db.restype = i32
        L.xs_lut_to_db.argtypes = [vp, vp, i64, vp]
        L.xs_lut_to_linear.restype = i32
        L.xs_lut_to_linear.argtypes = [vp, vp, i64, vp]
        L.xs_plan_create.restype = i32
        L.xs_plan_create.argtypes = [ctypes.POINTER(PlanDesc), vp, ctypes.POINTER(vp)]
        L.xs_plan_destroy.restype = None
        L.xs_plan_destroy.argtypes = [vp]
        L.xs_invert_workspace_bytes.restype = sz
        L.xs_invert_workspace_bytes.argtypes = [vp, i64, ctypes.c_uint32]
        L.xs_invert.restype = i32
        L.xs_invert.argtypes = [vp, ctypes.POINTER(InvertArgs), vp]
        L.xs_timer_create.restype = i32
        L.xs_timer_create.argtypes = [ctypes.POINTER(vp)]
        L.xs_timer_destroy.restype = None
        L.xs_timer_destroy.argtypes = [vp]
        L.xs_timer_elapsed_ms.restype = i32
        L.xs_timer_elapsed_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
        L.xs_detrend.restype = i32
        L.xs_detrend.argtypes = [vp, vp, i64, i64, i32, vp, vp]
        L.xs_dsig.restype = i32
        L.xs_dsig.argtypes = [i32, i32, vp, vp, vp, vp, i64, vp]
        L.xs_dsig_wspd.restype = i32
        L.xs_dsig_wspd.argtypes = [i32, i32, vp, vp, vp, i64, vp]
        L.xs_nesz_flatten_workspace_bytes.restype = sz
        L.xs_nesz_flatten_workspace_bytes.argtypes = [i64, i64]
        L.xs_nesz_flatten.restype = i32
        L.xs_nesz_flatten.argtypes = [vp, vp, i64, i64, i32, vp, vp, sz, vp]
        L.xs_local_gradients_workspace_bytes.restype = sz
        L.xs_local_gradients_workspace_bytes.argtypes = [i64, i64]
        L.xs_local_gradients.restype = i32
        L.xs_local_gradients.argtypes = [vp, i64, i64, i32, vp, vp, vp, vp, sz, vp]
        if L.xs_abi_version() != ABI_VERSION:
            raise NativeError("libxsarsea_b200.so ABI version mismatch")
        _lib = L
        return L


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().xs_last_error().decode(errors="replace")
        if rc == -3:
            raise ValueError(msg)  # scipy's bounds_error text (models.py:167)
        raise NativeError(f"{what} failed ({rc}): {msg}")


def launch_count() -> int:
    return int(load().xs_launch_count())


# ---- device helpers (torch = memory + streams) ----------------------------------------------------------------

def torch_cuda():
    """Return the torch module after checking that a CUDA device is usable; never falls back to CPU."""
    import torch

    if not torch.cuda.is_available():
        raise NativeError("xsarsea_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists.")
    return torch


def stream_ptr():
    torch = torch_cuda()
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def host_f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def hptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def dptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())
