"""Executes the dask / xarray container paths of `invert_from_model` (reference windspeed.py:333-388, ours
`_invert_dask` + the `xr.where` merge and `abs` of the lazy branch) with minimal stand-ins for the two packages
(tests/_stubs.py; neither is installable here).  CPU variant: the numeric core (`_invert_from_model_numpy`, i.e. the GPU) is
replaced by the oracle, so what is checked is the container logic -- dB prologue, all-NaN rasters for an absent polarisation,
row blocks, merge, attrs.  GPU variant: the real operator, against the plain numpy call."""
import warnings

import numpy as np
import pytest

import _stubs
import oracle
import xsarsea_b200  # noqa: F401  (imported BEFORE the stand-ins are installed: the package must not mistake them for xarray)
from oracle import lut as olut

KW = dict(inc_step_lr=2.0, wspd_step_lr=1.0, phi_step_lr=10.0, inc_step=1.0, wspd_step=0.5, phi_step=5.0)


def scene(shape=(23, 40), seed=4):
    rng = np.random.default_rng(seed)
    inc = rng.uniform(18, 48, shape)
    w, p = rng.uniform(2, 25, shape), rng.uniform(0, 360, shape)
    s_co = oracle.gmf_eval("gmf_cmod5n", inc, w, p) * np.exp(rng.normal(0, 0.05, shape))
    s_cr = oracle.gmf_eval("gmf_s1_v2", inc, w) * np.exp(rng.normal(0, 0.05, shape))
    anc = (w + rng.normal(0, 2, shape)) * np.exp(1j * np.deg2rad(p + rng.normal(0, 20, shape)))
    s_co[rng.uniform(size=shape) < 0.03] = np.nan
    return inc, s_co, s_cr, anc


def oracle_operator(models, dsig_co, kwargs, np_inc, co_db, cr_db, np_dsig_cr, np_anc):
    """Same signature and contract as `_invert_from_model_numpy` (windspeed.py:132-134), computed by the oracle."""
    kw = {}
    if models[0] is not None:
        lut, (gi, gw, gp) = olut.to_lut(models[0].name, units="dB", **kwargs)
        kw.update(co_lut=lut, inc_grid=gi, wspd_grid=gw, phi_grid=gp)
    if models[1] is not None:
        lut, (gi, gw, _) = olut.to_lut(models[1].name, units="dB", **kwargs)
        kw.update(cr_lut=lut, inc_cr_grid=gi, wspd_cr_grid=gw)
    with np.errstate(all="ignore"):
        oc, ox, _, _ = oracle.invert(np_inc, co_db, cr_db, np_dsig_cr, np_anc, dsig_co=dsig_co, **kw)
    return oc, ox


def lazy(da, xr, a, rows=5):
    return xr.DataArray(da.from_array(a, chunks=(rows, a.shape[1])), dims=("line", "sample"),
                        coords={"line": np.arange(a.shape[0]), "sample": np.arange(a.shape[1])}, attrs={"units": "linear"})


def run_all(ws, da, xr, numpy_call):
    inc, s_co, s_cr, anc = scene()
    L = lambda a: lazy(da, xr, a)
    model = ("gmf_cmod5n", "gmf_s1_v2")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # dual-pol: lazy containers in -> labelled lazy containers out, merged like windspeed.py:426-428
        co, dual = ws.invert_from_model(L(inc), L(s_co), L(s_cr), ancillary_wind=L(anc), dsig_cr=0.1, model=model, **KW)
        assert isinstance(co, xr.DataArray) and isinstance(co.data, da.Array) and co.dims == ("line", "sample")
        assert co.name == "windspeed_gmf" and co.attrs["model"] == "gmf_cmod5n" and "units" not in co.attrs
        assert dual.attrs["model"] == "gmf_cmod5n gmf_s1_v2"
        assert _stubs.CALLS["blocks"] == 5          # 23 lines in blocks of 5: the operator ran once per row block
        w_co, w_dual = numpy_call(inc, s_co, s_cr, anc, model)
        assert np.array_equal(np.asarray(co.data), w_co, equal_nan=True)
        assert np.array_equal(np.asarray(dual.data), w_dual, equal_nan=True)
        # mono co-pol and mono cross-pol (abs of the lazy result, units attr); only one input lazy is enough
        mono = ws.invert_from_model(L(inc), s_co, ancillary_wind=anc, model="gmf_cmod5n", **KW)
        assert np.array_equal(np.asarray(mono.data), w_co, equal_nan=True)
        x = ws.invert_from_model(L(inc), L(s_cr), model="gmf_s1_v2", **KW)
        assert x.attrs["units"] == "m/s" and np.asarray(x.data).dtype == np.float64
        w_x = numpy_call(inc, None, s_cr, None, "gmf_s1_v2")
        assert np.array_equal(np.asarray(x.data), w_x, equal_nan=True)
        # the fused-epilogue API falls back to the callers' own expressions on lazy containers
        (sp, dr), _ = ws.invert_to_speed_dir(L(inc), L(s_co), L(s_cr), ancillary_wind=L(anc), model=model, ground_heading=30.0, **KW)
        assert np.allclose(np.asarray(sp.data), np.abs(w_co), equal_nan=True)
        assert np.allclose(np.asarray(dr.data), (90 - np.angle(w_co, deg=True) + 30.0) % 360, equal_nan=True)


def test_dask_xarray_container_paths_with_the_oracle_as_operator(monkeypatch):
    da, xr = _stubs.install(monkeypatch)
    from xsarsea_b200 import windspeed as ws
    from xsarsea_b200.windspeed import windspeed as impl

    monkeypatch.setattr(impl, "_invert_from_model_numpy", oracle_operator)

    def numpy_call(inc, s_co, s_cr, anc, model):
        models = tuple(ws.get_model(m) for m in model) if isinstance(model, tuple) else (
            (ws.get_model(model), None) if s_co is not None else (None, ws.get_model(model)))
        nan = np.full(inc.shape, np.nan)
        with np.errstate(all="ignore"):
            co_db = 10 * np.log10(s_co + 1e-15) if s_co is not None else nan
            cr_db = 10 * np.log10(s_cr + 1e-15) if s_cr is not None else nan
            oc, ox = oracle_operator(models, 0.1, KW, inc, co_db, cr_db, np.full(inc.shape, 0.1), nan * 1j if anc is None else anc)
            if s_co is None:
                return np.abs(ox)
            if s_cr is None:
                return oc
            return oc, np.where((np.abs(oc) < 5) | (np.abs(ox) < 5), oc, ox)

    run_all(ws, da, xr, numpy_call)


@pytest.mark.gpu
def test_dask_xarray_container_paths_on_the_gpu(monkeypatch):
    da, xr = _stubs.install(monkeypatch)
    from xsarsea_b200 import windspeed as ws

    from xsarsea_b200.windspeed import windspeed as impl

    def numpy_call(inc, s_co, s_cr, anc, model):
        if s_co is None:
            return ws.invert_from_model(inc, s_cr, model=model, **KW)
        if s_cr is None:
            return ws.invert_from_model(inc, s_co, ancillary_wind=anc, model=model, **KW)
        # dual-pol: the lazy branch merges on the host (xr.where of numpy's abs) like the reference, the eager numpy branch
        # merges in the kernel; where a wind speed sits exactly on the 5 m/s node the two abs() may round differently
        # (DESIGN.md section 7 item 1), so the expectation is built from the unmerged operator output
        with np.errstate(all="ignore"):
            oc, ox = impl._invert_from_model_numpy(tuple(ws.get_model(m) for m in model), 0.1, dict(KW), inc, 10 * np.log10(s_co + 1e-15),
                                                   10 * np.log10(s_cr + 1e-15), np.full(inc.shape, 0.1), anc)
            return oc, np.where((np.abs(oc) < 5) | (np.abs(ox) < 5), oc, ox)

    run_all(ws, da, xr, numpy_call)
