"""Minimal stand-ins for `dask.array` and `xarray` (neither is installable in the build image) -- just enough of their
interfaces to EXECUTE the container paths of `invert_from_model` that the reference runs through them
(windspeed.py:333-388: `da.apply_gufunc` over row blocks, `xr.zeros_like`, `xr.where`).  Test infrastructure only.

`Array` mimics a row-chunked dask array (eager numpy inside, but typed as `dask.array.core.Array` and only touched through
ufuncs / `apply_gufunc` / `compute`), `DataArray` a labelled wrapper around it.  `install(monkeypatch)` registers both as
importable modules for the duration of a test.
"""
import sys
import types

import numpy as np
from numpy.lib.mixins import NDArrayOperatorsMixin


class Array(NDArrayOperatorsMixin):
    def __init__(self, data, chunk_rows):
        self._data = np.asarray(data)
        self.chunk_rows = int(chunk_rows)

    shape = property(lambda self: self._data.shape)
    dtype = property(lambda self: self._data.dtype)
    ndim = property(lambda self: self._data.ndim)

    def compute(self):
        return self._data

    def __array__(self, dtype=None, copy=None):
        return self._data if dtype is None else self._data.astype(dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__":
            return NotImplemented
        rows = max(x.chunk_rows for x in inputs if isinstance(x, Array))
        with np.errstate(all="ignore"):
            res = ufunc(*[x._data if isinstance(x, Array) else x for x in inputs], **kwargs)
        return Array(res, rows)


Array.__module__ = "dask.array.core"
CALLS = {"blocks": 0}


def from_array(a, chunks):
    return Array(a, chunks[0] if isinstance(chunks, (tuple, list)) else chunks)


def apply_gufunc(func, signature, *args, output_dtypes=None, **kwargs):
    """Core dimension = the last axis; loops over blocks of `chunk_rows` lines like dask does over its chunks."""
    assert signature == "(n),(n),(n),(n),(n)->(n),(n)"
    rows = max(a.chunk_rows for a in args if isinstance(a, Array))
    arrs = [np.asarray(a) for a in args]
    n = arrs[0].shape[0]
    outs = None
    for lo in range(0, n, rows):
        res = func(*[a[lo:lo + rows] for a in arrs])
        CALLS["blocks"] += 1
        outs = [[r] for r in res] if outs is None else [o + [r] for o, r in zip(outs, res)]
    return tuple(Array(np.concatenate(o, axis=0), rows) for o in outs)


class DataArray(NDArrayOperatorsMixin):
    def __init__(self, data, dims=None, coords=None, attrs=None, name=None):
        self.data = data
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{i}" for i in range(np.ndim(data)))
        self.coords = dict(coords or {})
        self.attrs = dict(attrs or {})
        self.name = name

    shape = property(lambda self: self.data.shape)
    dtype = property(lambda self: self.data.dtype)
    ndim = property(lambda self: self.data.ndim)
    values = property(lambda self: np.asarray(self.data))

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data) if dtype is None else np.asarray(self.data).astype(dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__":
            return NotImplemented
        res = getattr(ufunc, method)(*[x.data if isinstance(x, DataArray) else x for x in inputs], **kwargs)
        first = next(x for x in inputs if isinstance(x, DataArray))
        return DataArray(res, first.dims, first.coords, {}, first.name)   # xarray drops attrs in arithmetic by default

    def copy(self):
        return DataArray(self.data, self.dims, self.coords, dict(self.attrs), self.name)

    def astype(self, dtype):
        d = self.data
        return DataArray(Array(d._data.astype(dtype), d.chunk_rows) if isinstance(d, Array) else np.asarray(d).astype(dtype),
                         self.dims, self.coords, dict(self.attrs), self.name)

    def compute(self):
        d = self.data
        return DataArray(d.compute() if isinstance(d, Array) else d, self.dims, self.coords, dict(self.attrs), self.name)


def zeros_like(obj, dtype=None):
    d = obj.data
    z = np.zeros(d.shape, dtype=dtype or d.dtype)
    return DataArray(Array(z, d.chunk_rows) if isinstance(d, Array) else z, obj.dims, obj.coords, dict(obj.attrs), obj.name)


def where(cond, a, b):
    first = next(x for x in (cond, a, b) if isinstance(x, DataArray))
    raw = [np.asarray(x.data) if isinstance(x, DataArray) else x for x in (cond, a, b)]
    res = np.where(*raw)
    d = first.data
    return DataArray(Array(res, d.chunk_rows) if isinstance(d, Array) else res, first.dims, first.coords, {}, first.name)


def install(monkeypatch):
    """Make `import dask.array as da` and `import xarray as xr` resolve to the stand-ins."""
    dask = types.ModuleType("dask")
    da = types.ModuleType("dask.array")
    da.Array, da.apply_gufunc, da.from_array = Array, apply_gufunc, from_array
    dask.array = da
    xr = types.ModuleType("xarray")
    xr.DataArray, xr.zeros_like, xr.where = DataArray, zeros_like, where
    for name, mod in (("dask", dask), ("dask.array", da), ("xarray", xr)):
        monkeypatch.setitem(sys.modules, name, mod)
    CALLS["blocks"] = 0
    return da, xr
