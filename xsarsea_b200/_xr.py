"""Labelled-array plumbing.

The reference returns xarray.DataArray objects (LUTs, GMF outer products, inversion results).  xarray is an
optional dependency here (it is absent from the build image): when it is importable real DataArrays are
produced, otherwise `DataArrayLite`, a small stand-in exposing the subset of the DataArray interface this
package and its callers rely on (data/values/dims/coords/attrs/name, transpose, isel, squeeze, copy, item,
coordinate access by attribute).  Everything in the package treats labelled inputs by duck typing
(`.dims`, `.data`, `.coords`, `.attrs`), so real DataArrays, dask-backed ones included, pass through.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - not installed in the build image
    import xarray as _xarray

    HAVE_XARRAY = True
except Exception:  # ImportError or a broken install
    _xarray = None
    HAVE_XARRAY = False


class DataArrayLite:
    """Minimal stand-in for xarray.DataArray (see module docstring)."""

    __array_priority__ = 50

    def __init__(self, data, dims=None, coords=None, attrs=None, name=None):
        self.data = data if hasattr(data, "shape") else np.asarray(data)
        nd = self.data.ndim
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{i}" for i in range(nd))
        if len(self.dims) != nd:
            raise ValueError(f"dims {self.dims} do not match data of rank {nd}")
        self.coords = {}
        for k, v in (coords or {}).items():
            v = np.asarray(v.data if isinstance(v, DataArrayLite) else v)
            self.coords[k] = v
        self.attrs = dict(attrs or {})
        self.name = name

    # -- numpy protocol ------------------------------------------------------------------------------------
    @property
    def values(self):
        return np.asarray(self.data)

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self.data)
        return a.astype(dtype) if dtype is not None else a

    shape = property(lambda self: self.data.shape)
    ndim = property(lambda self: self.data.ndim)
    dtype = property(lambda self: self.data.dtype)
    size = property(lambda self: int(np.prod(self.data.shape)))

    def __len__(self):
        return self.data.shape[0]

    def __getattr__(self, item):
        coords = self.__dict__.get("coords", {})
        if item in coords:
            return DataArrayLite(coords[item], dims=(item,), coords={item: coords[item]}, name=item)
        raise AttributeError(item)

    def __getitem__(self, key):
        if isinstance(key, str):
            return getattr(self, key)
        return np.asarray(self.data)[key]

    def __repr__(self):
        shape = ", ".join(f"{d}: {n}" for d, n in zip(self.dims, self.shape))
        return f"<DataArrayLite {self.name or ''}({shape}) attrs={self.attrs}>"

    # -- the DataArray subset ---------------------------------------------------------------------------------
    def copy(self, data=None, deep=True):
        d = (np.array(self.data) if deep else self.data) if data is None else data
        return DataArrayLite(d, self.dims, self.coords, self.attrs, self.name)

    def astype(self, dtype):
        return self.copy(data=np.asarray(self.data).astype(dtype))

    def item(self):
        return np.asarray(self.data).item()

    def transpose(self, *dims):
        if not dims:
            dims = self.dims[::-1]
        order = [self.dims.index(d) for d in dims]
        return DataArrayLite(np.transpose(np.asarray(self.data), order), dims, self.coords, self.attrs, self.name)

    def isel(self, **indexers):
        data = np.asarray(self.data)
        dims = list(self.dims)
        coords = dict(self.coords)
        for d, idx in indexers.items():
            ax = dims.index(d)
            data = np.take(data, idx, axis=ax)
            if d in coords and coords[d].ndim == 1:
                coords[d] = np.take(coords[d], idx)
            if np.ndim(idx) == 0:
                dims.pop(ax)
                coords.pop(d, None)
        return DataArrayLite(data, dims, {k: v for k, v in coords.items() if k in dims}, self.attrs, self.name)

    def squeeze(self, dim=None):
        dims = [dim] if isinstance(dim, str) else (list(dim) if dim is not None else
                                                   [d for d, n in zip(self.dims, self.shape) if n == 1])
        data = np.asarray(self.data)
        keep = [d for d in self.dims if d not in dims]
        data = data.reshape([n for d, n in zip(self.dims, self.shape) if d not in dims])
        return DataArrayLite(data, keep, {k: v for k, v in self.coords.items() if k in keep}, self.attrs, self.name)

    def drop_vars(self, names):
        names = [names] if isinstance(names, str) else list(names)
        return DataArrayLite(self.data, self.dims, {k: v for k, v in self.coords.items() if k not in names},
                             self.attrs, self.name)


class DatasetLite(dict):
    """Minimal stand-in for xarray.Dataset: a dict of variables with attribute access (`ds.G2`, `ds["G2"]`)."""

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError:
            raise AttributeError(item) from None

    @property
    def data_vars(self):
        return self


def make_dataset(variables):
    """xarray.Dataset of named DataArrays when xarray is installed, DatasetLite otherwise."""
    if HAVE_XARRAY:
        return _xarray.merge([v.rename(k) if getattr(v, "name", None) != k else v for k, v in variables.items()])
    return DatasetLite(variables)


def is_labelled(x) -> bool:
    """True for xarray.DataArray and DataArrayLite (duck typed)."""
    return hasattr(x, "dims") and hasattr(x, "data") and hasattr(x, "attrs")


def is_dask(x) -> bool:
    d = x.data if is_labelled(x) else x
    return type(d).__module__.split(".")[0] == "dask"


def make_dataarray(data, dims, coords=None, attrs=None, name=None):
    """xarray.DataArray when xarray is installed, DataArrayLite otherwise."""
    if HAVE_XARRAY:
        return _xarray.DataArray(data, dims=dims, coords=coords, attrs=attrs, name=name)
    return DataArrayLite(data, dims, coords, attrs, name)


def like(template, data, name=None, attrs=None):
    """A labelled array shaped like `template` (same dims/coords) holding `data`; attrs replaced."""
    if HAVE_XARRAY and isinstance(template, _xarray.DataArray):
        out = _xarray.DataArray(data, dims=template.dims, coords=template.coords, name=name)
    else:
        coords = {k: v for k, v in getattr(template, "coords", {}).items()
                  if k in template.dims and np.ndim(v) == 1}
        out = DataArrayLite(data, template.dims, coords, None, name)
    out.attrs.update(attrs or {})
    return out
