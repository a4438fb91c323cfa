"""The public Python API on the GPU against the oracle: same calls a user of xsarsea makes (test/test_xsarsea.py
of the reference checks only types; values are pinned here through the oracle).
Tolerance of BASELINE.json: wind speed <= 1e-3 m/s, direction <= 0.1 deg, indices exact except documented near-ties
(DESIGN.md): device-built LUT values differ from the host's by ~1e-16 relative."""
import warnings

import numpy as np
import pytest

import oracle
from oracle import lut as olut

pytestmark = pytest.mark.gpu

KW = dict(inc_step_lr=2.0, wspd_step_lr=1.0, phi_step_lr=10.0, inc_step=0.5, wspd_step=0.25, phi_step=2.5)


@pytest.fixture(scope="module")
def ws():
    import torch

    assert torch.cuda.is_available()
    from xsarsea_b200 import windspeed

    return windspeed


def reset_steps(m):
    """`to_lut(**kwargs)` stores its steps on the model (reference behaviour, gmfs.py:367-379); back to the defaults."""
    for k, v in dict(inc_step=0.1, wspd_step=0.1, phi_step=1.0, inc_step_lr=1.0, wspd_step_lr=0.2, phi_step_lr=2.5).items():
        setattr(m, k, v)
    return m


def synth(n, seed=0, shape=None):
    rng = np.random.default_rng(seed)
    inc = rng.uniform(17, 49, n)
    w = rng.uniform(2, 25, n)
    phi = rng.uniform(0, 360, n)
    s_co = oracle.gmf_eval("gmf_cmod5n", inc, w, phi) * np.exp(rng.normal(0, 0.05, n))
    s_cr = oracle.gmf_eval("gmf_s1_v2", inc, w) * np.exp(rng.normal(0, 0.05, n))
    anc = (w + rng.normal(0, 2, n)) * np.exp(1j * np.deg2rad(phi + rng.normal(0, 20, n)))
    s_co[rng.uniform(size=n) < 0.01] = np.nan
    arrs = [inc, s_co, s_cr, anc]
    if shape:
        arrs = [a.reshape(shape) for a in arrs]
    return arrs


def oracle_dual(inc, s_co, s_cr, anc, dsig_cr=0.1, kw=KW):
    co_lut, (gi, gw, gp) = olut.to_lut("gmf_cmod5n", units="dB", **kw)
    cr_lut, (gic, gwc, _) = olut.to_lut("gmf_s1_v2", units="dB", **kw)
    with np.errstate(all="ignore"):
        co_db = 10 * np.log10(s_co + 1e-15) if s_co is not None else np.full(inc.shape, np.nan)
        cr_db = 10 * np.log10(s_cr + 1e-15) if s_cr is not None else np.full(inc.shape, np.nan)
        anc = anc if anc is not None else np.full(inc.shape, np.nan + 0j)
        return oracle.invert(inc, co_db, cr_db, dsig_cr, anc, co_lut=co_lut, inc_grid=gi, wspd_grid=gw, phi_grid=gp,
                             cr_lut=cr_lut, inc_cr_grid=gic, wspd_cr_grid=gwc)


def wind_close(got, want, frac=0.003):
    """NaN pattern equal; |speed| within 1e-3 m/s and direction within 0.1 deg for all but `frac` of the pixels."""
    got, want = np.asarray(got), np.asarray(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    g, w = got[ok], want[ok]
    if np.iscomplexobj(w):
        dspd = np.abs(np.abs(g) - np.abs(w))
        ddir = np.abs(np.angle(g * np.conj(w), deg=True))
        ddir[(np.abs(w) == 0) | (np.abs(g) == 0)] = 0
        bad = (dspd > 1e-3) | (ddir > 0.1)
    else:
        bad = np.abs(g - w) > 1e-3
    assert bad.mean() <= frac, f"{bad.sum()} of {bad.size} pixels outside tolerance"


def test_dual_pol_numpy(ws):
    inc, s_co, s_cr, anc = synth(6000, shape=(60, 100))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        co, dual = ws.invert_from_model(inc, s_co, s_cr, ancillary_wind=anc, dsig_cr=0.1, model=("gmf_cmod5n", "gmf_s1_v2"), **KW)
    assert isinstance(co, np.ndarray) and co.dtype == np.complex128 and co.shape == inc.shape
    assert isinstance(dual, np.ndarray) and dual.dtype == np.complex128
    o_co, o_du, _, _ = oracle_dual(inc, s_co, s_cr, anc)
    with np.errstate(invalid="ignore"):
        merged = np.where((np.abs(o_co) < 5) | (np.abs(o_du) < 5), o_co, o_du)   # windspeed.py:426-428
    wind_close(co, o_co)
    wind_close(dual, merged)


def test_device_resident_tensors_in_tensors_out(ws):
    """torch CUDA tensors in -> torch CUDA tensors out, identical to the host-array call (no copies in between)."""
    import torch

    inc, s_co, s_cr, anc = synth(6000, shape=(60, 100))
    model = ("gmf_cmod5n", "gmf_s1_v2")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        co, dual = ws.invert_from_model(inc, s_co, s_cr, ancillary_wind=anc, dsig_cr=0.1, model=model, **KW)
        t = lambda a: torch.from_numpy(a).cuda()
        tco, tdual = ws.invert_from_model(t(inc), t(s_co), t(s_cr), ancillary_wind=t(anc), dsig_cr=0.1, model=model, **KW)
        assert tco.is_cuda and tdual.is_cuda and tco.dtype == torch.complex128 and tuple(tco.shape) == inc.shape
        assert np.array_equal(tco.cpu().numpy(), co, equal_nan=True)
        assert np.array_equal(tdual.cpu().numpy(), dual, equal_nan=True)
        # mono cross-pol: float64 wind speed; dsig_cr as a device raster produced by get_dsig
        dsig = ws.get_dsig("gmf_s1_v2", t(inc), t(s_cr), 10 ** -3.2)
        x_dev = ws.invert_from_model(t(inc), t(s_cr), dsig_cr=dsig, model="gmf_s1_v2", **KW)
        x_host = ws.invert_from_model(inc, s_cr, dsig_cr=dsig.cpu().numpy(), model="gmf_s1_v2", **KW)
        assert x_dev.is_cuda and x_dev.dtype == torch.float64
        assert np.array_equal(x_dev.cpu().numpy(), x_host, equal_nan=True)


def test_index_exact_with_the_models_own_lut(ws):
    """Feeding the oracle the device-built LUT removes the libm difference: answers must then agree exactly
    (up to the fused log10 prologue's ulp, i.e. measure-zero near-ties)."""
    inc, s_co, s_cr, anc = synth(5000, seed=3)
    m = ws.get_model("gmf_cmod5n")
    lut = m.to_lut(units="dB", **KW)
    co = ws.invert_from_model(inc, s_co, ancillary_wind=anc, model="gmf_cmod5n", **KW)
    with np.errstate(all="ignore"):
        o_co, _, _, _ = oracle.invert(inc, 10 * np.log10(s_co + 1e-15), np.nan, 0.1, anc, co_lut=np.asarray(lut),
                                      inc_grid=np.asarray(lut.incidence), wspd_grid=np.asarray(lut.wspd),
                                      phi_grid=np.asarray(lut.phi))
    assert np.array_equal(np.isnan(co), np.isnan(o_co))
    ok = ~np.isnan(o_co)
    assert (np.abs(co[ok] - o_co[ok]) > 1e-9).mean() < 5e-4


def test_mono_pol_return_contract(ws):
    inc, s_co, s_cr, anc = synth(3000, seed=1)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        co = ws.invert_from_model(inc, s_co, ancillary_wind=anc, model="cmod5n", **KW)     # alias
    assert any("Unable to check sigma0 pol" in str(w.message) for w in rec)                 # windspeed.py:96
    assert co.dtype == np.complex128
    o_co, o_du, _, _ = oracle_dual(inc, s_co, None, anc)
    wind_close(co, o_co)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cr = ws.invert_from_model(inc, s_cr, model="gmf_s1_v2", **KW)
    assert cr.dtype == np.float64                                                            # windspeed.py:422-423
    _, o_x, _, _ = oracle_dual(inc, None, s_cr, None)
    wind_close(cr, np.abs(o_x))
    with pytest.raises(AssertionError):                                                      # windspeed.py:107
        ws.invert_from_model(inc, s_co, model="gmf_cmod5n", **KW)
    with pytest.raises(KeyError):
        ws.invert_from_model(inc, s_co, ancillary_wind=anc, model="gmf_nope")


def test_labelled_inputs_and_pol_check(ws):
    from xsarsea_b200._xr import DataArrayLite, is_labelled

    inc, s_co, s_cr, anc = synth(2400, seed=2, shape=(40, 60))
    coords = dict(line=np.arange(40), sample=np.arange(60))
    L = lambda a: DataArrayLite(a, ("line", "sample"), coords, attrs=dict(units="x"))
    co, dual = ws.invert_from_model(L(inc), L(s_co), L(s_cr), ancillary_wind=L(anc), dsig_cr=L(np.full(inc.shape, 0.1)),
                                    model=("gmf_cmod5n", "gmf_s1_v2"), **KW)
    assert is_labelled(co) and is_labelled(dual) and co.dims == ("line", "sample") and co.name == "windspeed_gmf"
    assert co.attrs == {"comment": "wind speed and direction inverted from model gmf_cmod5n (VV)", "model": "gmf_cmod5n"}
    assert dual.attrs["model"] == "gmf_cmod5n gmf_s1_v2"
    o_co, o_du, _, _ = oracle_dual(inc, s_co, s_cr, anc)
    wind_close(co.values, o_co)

    class Pol:  # sigma0.pol.values.item() as on an xsar dataset
        def __init__(self, p):
            self.values = np.array(p)

    s = L(s_co)
    s.pol = Pol("HH")
    with pytest.raises(ValueError, match="can only handle"):
        ws.invert_from_model(L(inc), s, ancillary_wind=L(anc), model="gmf_cmod5n", **KW)


def test_streamed_blocks_equal_single_block(ws):
    from xsarsea_b200.windspeed import windspeed as impl

    inc, s_co, s_cr, anc = synth(5000, seed=4)
    dsig = np.full(5000, 0.1)
    dsig[::7] = 0.3
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = ws.invert_from_model(inc, s_co, s_cr, ancillary_wind=anc, dsig_cr=dsig, model=("gmf_cmod5n", "gmf_s1_v2"), **KW)
        old = impl.BLOCK_PIXELS
        impl.BLOCK_PIXELS = 700
        try:
            b = ws.invert_from_model(inc, s_co, s_cr, ancillary_wind=anc, dsig_cr=dsig, model=("gmf_cmod5n", "gmf_s1_v2"), **KW)
        finally:
            impl.BLOCK_PIXELS = old
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)


def test_user_gmf_and_nc_lut_roundtrip(ws, tmp_path):
    @ws.GmfModel.register(pol="VH", units="linear", defer=False, inc_range=[17.0, 50.0], wspd_range=[3.0, 80.0])
    def gmf_dummy(inc, wspd, phi=None):       # the docstring example of gmfs.py:45-56 / test_xsarsea.py:8-21
        a = 0.00013106836021008122 + -4.530598283705591e-06 * inc + 4.429277425062766e-08 * inc ** 2
        b = 1.3925444179360706 + 0.004157838450541205 * inc + 3.4735809771069953e-05 * inc ** 2
        return a * wspd ** b

    m = ws.get_model("gmf_dummy")
    assert m.iscrosspol and m.phi_range is None and "gmf_dummy" in ws.available_models().index
    got = m(np.arange(20, 22), np.arange(10, 12))
    np.testing.assert_allclose(np.asarray(got), [[0.00179606, 0.00207004], [0.0017344, 0.00200004]], rtol=2e-6)
    assert np.isscalar(m(20.0, 10.0))
    # persist as a LUT file in the reference's schema and read it back as an NcLutModel
    path = tmp_path / "nc_lut_dummy.nc"
    m.to_netcdf(str(path))
    ws.register_nc_luts(str(tmp_path))
    nc = ws.get_model("nc_lut_dummy")
    assert nc.pol == "VH" and nc.short_name == "dummy" and type(nc).__name__ == "NcLutModel"
    a, b = m.to_lut(units="dB", resolution="high"), nc.to_lut(units="dB")
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=0, atol=1e-12)
    # both invert identically (cross-pol only, no ancillary)
    rng = np.random.default_rng(0)
    inc = rng.uniform(18, 49, 2000)
    s = np.asarray([gmf_dummy(i, w) for i, w in zip(inc, rng.uniform(4, 60, 2000))])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        w1 = ws.invert_from_model(inc, s, model="gmf_dummy", resolution="high")
        w2 = ws.invert_from_model(inc, s, model="nc_lut_dummy")
    assert (np.abs(w1 - w2) > 1e-9).mean() < 1e-3
    with pytest.raises(NotImplementedError):
        nc(np.zeros((2, 2)), np.zeros((2, 2)))     # models.py:320-329
    v = nc(np.array([20.0, 21.0]), np.array([10.0, 11.0]), units="linear")
    np.testing.assert_allclose(np.asarray(v), [[0.00179606, 0.00207004], [0.0017344, 0.00200004]], rtol=1e-3)


def test_model_call_shapes(ws):
    m = ws.get_model("gmf_cmod5n")
    assert abs(m(35.0, 10.0, 45.0) - 0.05376709128885202) < 1e-15
    out = m(np.array([30.0, 35.0]), np.array([5.0, 10.0, 15.0]), np.array([0.0, 45.0, 90.0, 135.0]))
    assert out.dims == ("incidence", "wspd", "phi") and out.shape == (2, 3, 4) and out.attrs["units"] == "linear"
    want = oracle.lut_build("gmf_cmod5n", [30.0, 35.0], [5.0, 10.0, 15.0], [0.0, 45.0, 90.0, 135.0])
    np.testing.assert_allclose(np.asarray(out), want, rtol=1e-12)
    inc2d = np.linspace(20, 40, 12).reshape(3, 4)
    b = m(inc2d, 10.0 + 0 * inc2d, 45.0 + 0 * inc2d)
    np.testing.assert_allclose(b, oracle.gmf_eval("gmf_cmod5n", inc2d, 10.0, 45.0), rtol=1e-12)
    x = ws.get_model("gmf_rs2_v3")
    o = x(np.array([30.0, 35.0]), np.array([5.0, 10.0, 15.0]))
    assert o.dims == ("incidence", "wspd")
    np.testing.assert_allclose(np.asarray(o), oracle.lut_build("gmf_rs2_v3", [30.0, 35.0], [5.0, 10.0, 15.0]), rtol=1e-12)


def test_sigma0_detrend_api(ws):
    import xsarsea_b200
    from xsarsea_b200._xr import DataArrayLite

    rng = np.random.default_rng(0)
    h, w = 50, 400
    inc = np.broadcast_to(np.linspace(30, 45, w), (h, w)).copy()
    s0 = rng.uniform(0.01, 0.3, (h, w))
    L = lambda a: DataArrayLite(a, ("line", "sample"), dict(line=np.arange(h), sample=np.arange(w)))
    out = xsarsea_b200.sigma0_detrend(L(s0), L(inc))
    prof = oracle.gmf_eval("gmf_cmod5n", inc[0], 10.0, 45.0)
    want = s0 / (prof / np.nanmean(prof))
    np.testing.assert_allclose(out.values, want, rtol=1e-12)
    assert out.attrs["comment"] == "detrended with model gmf_cmod5n" and out.dims == ("line", "sample")
    with pytest.raises(AttributeError):
        xsarsea_b200.sigma0_detrend(s0, inc)                           # numpy inc has no .isel (detrend.py:55)
    with pytest.raises(ValueError):
        xsarsea_b200.sigma0_detrend(L(s0), L(inc), wind_speed_gmf=np.array([5.0, 10.0]))


def test_config1_full_size_fast_equals_fp64(ws):
    """BASELINE.json configs[0] at full size (1000 x 1000, cmod5n co-pol, default 501x499x181 LUT): the FP32 scan with
    FP64 refinement must give exactly the indices of the exhaustive FP64 kernel on every pixel."""
    import torch

    from xsarsea_b200 import _device as D
    from xsarsea_b200 import _native as nat
    from xsarsea_b200.windspeed import windspeed as impl

    m = reset_steps(ws.get_model("gmf_cmod5n"))
    plan = impl._get_plan(m, None, 0.1, {})
    assert plan.co_lut.shape == (501, 499, 181)
    g = torch.Generator(device="cuda").manual_seed(0)
    H = W = 1000
    f64 = dict(device="cuda", dtype=torch.float64)
    inc = (17.5 + 32 * torch.arange(W, **f64) / (W - 1)).expand(H, W).contiguous()
    w = 2 + 23 * torch.rand(H, W, generator=g, **f64)
    p = 360 * torch.rand(H, W, generator=g, **f64)
    s_co = D.gmf_eval(nat.GMF_IDS["gmf_cmod5n"], inc, w, p) * torch.exp(0.05 * torch.randn(H, W, generator=g, **f64))
    anc = torch.polar((w + 2 * torch.randn(H, W, generator=g, **f64)).abs(), torch.deg2rad(p + 20 * torch.randn(H, W, generator=g, **f64)))
    a, _, ia, _ = plan.invert(inc, s_co, None, 0.1, anc, want_idx=True)
    st = plan.last_stats()
    b, _, ib, _ = plan.invert(inc, s_co, None, 0.1, anc, want_idx=True, mode=nat.MODE_FP64)
    assert torch.equal(ia, ib) and torch.equal(torch.view_as_real(a), torch.view_as_real(b))
    assert st["scan_pixels"] + st["exhaustive_pixels"] == H * W and st["exhaustive_pixels"] < 0.02 * H * W


def test_full_iw_scene_properties(ws):
    """BASELINE.json configs[2] at full size (16700 x 25000 dual-pol): size-independent properties -- the scene is a
    1670-line block repeated 10 times, so every repeat must give bit-identical output, and the first lines must equal
    a separate small inversion (independence of pixels / of the tiling and binning)."""
    import torch

    import bench
    from xsarsea_b200.windspeed import windspeed as impl

    plan = impl._get_plan(reset_steps(ws.get_model("gmf_cmod5n")), reset_steps(ws.get_model("gmf_s1_v2")), 0.1, {})
    assert plan.co_lut.shape == (501, 499, 181) and plan.cr_lut.shape == (501, 771)
    rep, blk = 10, 1670
    inc, s_co, s_cr, anc = bench.synth_scene_device(blk, 25000, 1)
    tile = lambda t: t.repeat(rep, 1)
    co, du, _, _ = plan.invert(tile(inc), tile(s_co), tile(s_cr), 0.1, tile(anc), merge_dual=True)
    co_r = torch.view_as_real(co).view(rep, blk, 25000, 2)
    du_r = torch.view_as_real(du).view(rep, blk, 25000, 2)
    for r in range(1, rep):
        assert torch.equal(co_r[0].nan_to_num(-1.0), co_r[r].nan_to_num(-1.0))
        assert torch.equal(du_r[0].nan_to_num(-1.0), du_r[r].nan_to_num(-1.0))
    k = 40
    co2, du2, _, _ = plan.invert(inc[:k].contiguous(), s_co[:k].contiguous(), s_cr[:k].contiguous(), 0.1, anc[:k].contiguous(),
                                 merge_dual=True)
    assert torch.equal(torch.view_as_real(co2).nan_to_num(-1.0), co_r[0, :k].nan_to_num(-1.0))
    assert torch.equal(torch.view_as_real(du2).nan_to_num(-1.0), du_r[0, :k].nan_to_num(-1.0))
    w = co.abs()
    assert torch.isnan(w).float().mean().item() < 0.02 and 2 < torch.nanmean(w).item() < 30


def _sample_vs_oracle(idx, inc, s_co_db, s_cr_db, dsig, anc, co, cr, got_co, got_x):
    """Compare the pixels `idx` of a large device result with the oracle run on the same LUTs."""
    h = lambda t: None if t is None else t.reshape(-1)[idx].cpu().numpy()
    kw = {}
    if co is not None:
        kw.update(co_lut=co[0], inc_grid=co[1], wspd_grid=co[2], phi_grid=co[3])
    if cr is not None:
        kw.update(cr_lut=cr[0], inc_cr_grid=cr[1], wspd_cr_grid=cr[2])
    n = idx.numel()
    nan = np.full(n, np.nan)
    with np.errstate(all="ignore"):
        o_co, o_x, _, _ = oracle.invert(h(inc), nan if s_co_db is None else h(s_co_db), nan if s_cr_db is None else h(s_cr_db),
                                        h(dsig) if hasattr(dsig, "reshape") else dsig, nan + 0j if anc is None else h(anc), **kw)
    if got_co is not None:
        assert np.allclose(h(got_co), o_co, rtol=0, atol=1e-9, equal_nan=True)
    return h(got_x), o_x


def test_config4_cross_pol_only_full_size(ws):
    """BASELINE.json configs[3] at full size (10 000 x 10 000 cross-pol only, NetCDF-style LUT 331 x 771, no ancillary
    wind): a 30 000-pixel sample equals the oracle (dB inputs, uploaded LUT -> exact), the run is deterministic, and the
    wind speeds are LUT grid values."""
    import torch

    import bench
    from xsarsea_b200 import _device as D

    gi, gwc = np.linspace(17.0, 50.0, 331), np.linspace(3.0, 80.0, 771)
    cr = 10 * np.log10(oracle.lut_build("gmf_s1_v2", gi, gwc) + 1e-15)
    plan = D.InversionPlan(cr=(D.to_device(cr), gi, gwc))
    g = torch.Generator(device="cuda").manual_seed(4)
    H = W = 10000
    f64 = dict(device="cuda", dtype=torch.float64)
    inc = (20 + 29 * torch.arange(W, **f64) / (W - 1)).expand(H, W).contiguous()
    s_db = -38 + 30 * torch.rand(H, W, generator=g, **f64)          # spans below / inside / above the LUT range
    s_db[torch.rand(H, W, generator=g, device="cuda") < 0.01] = float("nan")
    a, _, _, _ = None, None, None, None
    _, w1, _, _ = plan.invert(inc, None, s_db, 0.1, None, sigma0_db=True, cr_abs=True)
    _, w2, _, _ = plan.invert(inc, None, s_db, 0.1, None, sigma0_db=True, cr_abs=True)
    assert torch.equal(w1.nan_to_num(-1.0), w2.nan_to_num(-1.0))
    idx = torch.randint(0, H * W, (30000,), generator=g, device="cuda")
    got, want = _sample_vs_oracle(idx, inc, None, s_db, 0.1, None, None, (cr, gi, gwc), None, w1)
    assert np.allclose(got, np.abs(want), rtol=0, atol=1e-12, equal_nan=True)
    vals = w1[~torch.isnan(w1)][:200000].cpu().numpy()
    assert np.isin(vals, gwc).all()


def test_config5_ew_scene_dual_pol_with_dsig_raster_and_detrend(ws):
    """BASELINE.json configs[4], one of the 8 EW scenes (10 000 x 10 400, inc 19-47 deg): dual-pol inversion with the
    dsig_cr raster of windspeed/utils.py:82-86 ((1.25/(sigma0_cr/nesz))**4, spanning 1e-9 .. 1e3) checked on a sample
    against the oracle, and sigma0_detrend's round trip out * ratio == sigma0."""
    import torch

    from xsarsea_b200 import _device as D
    from xsarsea_b200 import _native as nat
    from xsarsea_b200.windspeed import windspeed as impl

    m_co, m_cr = reset_steps(ws.get_model("gmf_cmod5n")), reset_steps(ws.get_model("gmf_s1_v2"))
    plan = impl._get_plan(m_co, m_cr, 0.1, {})
    H, W = 10000, 10400
    g = torch.Generator(device="cuda").manual_seed(5)
    f64 = dict(device="cuda", dtype=torch.float64)
    inc = (19 + 28 * torch.arange(W, **f64) / (W - 1)).expand(H, W).contiguous()
    w = 2 + 23 * torch.rand(H, W, generator=g, **f64)
    p = 360 * torch.rand(H, W, generator=g, **f64)
    s_co = D.gmf_eval(nat.GMF_IDS["gmf_cmod5n"], inc, w, p) * torch.exp(0.05 * torch.randn(H, W, generator=g, **f64))
    s_cr = D.gmf_eval(nat.GMF_IDS["gmf_s1_v2"], inc, w, None) * torch.exp(0.05 * torch.randn(H, W, generator=g, **f64))
    anc = torch.polar((w + 2 * torch.randn(H, W, generator=g, **f64)).abs(), torch.deg2rad(p + 20 * torch.randn(H, W, generator=g, **f64)))
    dsig = (1.25 / (s_cr / 10 ** -3.2)) ** 4.0
    co, du, _, _ = plan.invert(inc, s_co, s_cr, dsig, anc)
    # oracle on a sample, with the device-built LUTs downloaded (removes the libm difference), dB computed by numpy
    idx = torch.randint(0, H * W, (8000,), generator=g, device="cuda")
    co_l, cr_l = plan.co_lut.cpu().numpy(), plan.cr_lut.cpu().numpy()
    hs = lambda t: t.reshape(-1)[idx].cpu().numpy()
    with np.errstate(all="ignore"):
        o_co, o_du, _, _ = oracle.invert(hs(inc), 10 * np.log10(hs(s_co) + 1e-15), 10 * np.log10(hs(s_cr) + 1e-15), hs(dsig), hs(anc),
                                         co_lut=co_l, inc_grid=plan.co_grids[0], wspd_grid=plan.co_grids[1], phi_grid=plan.co_grids[2],
                                         cr_lut=cr_l, inc_cr_grid=plan.cr_grids[0], wspd_cr_grid=plan.cr_grids[1])
    assert (np.abs(hs(co) - o_co) > 1e-9).mean() < 1e-3       # fused log10 prologue: ulp-level near-ties only
    assert (np.abs(hs(du) - o_du) > 1e-9).mean() < 1e-3
    # detrend round trip on the same raster
    prof = D.gmf_eval(nat.GMF_IDS["gmf_cmod5n"], inc[0].contiguous(), torch.full((W,), 10.0, **f64), torch.full((W,), 45.0, **f64))
    out = D.detrend(s_co, prof)
    ratio = prof / prof.mean()
    assert torch.allclose(out * ratio, s_co, rtol=1e-14, atol=0)


def test_device_resident_chain_noise_flattening_to_gradients(ws):
    """The whole chain a dual-pol user runs (docs/examples/windspeed_retrieval_L1.ipynb cells 26-33 + the streaks
    notebook), device-resident from end to end: nesz_flattening -> get_dsig -> invert_from_model (dual-pol, dsig_cr
    raster) and sigma0_detrend -> local_gradients, against the same chain computed by the oracles on the host."""
    import torch

    import xsarsea_b200
    from oracle import dsig as od
    from oracle import gradients as og
    from xsarsea_b200 import _xr

    h, w = 120, 260
    rng = np.random.default_rng(17)
    inc = np.broadcast_to(np.linspace(20.0, 45.0, w), (h, w)).copy()
    wspd, phi = rng.uniform(3, 22, (h, w)), rng.uniform(0, 360, (h, w))
    s_co = oracle.gmf_eval("gmf_cmod5n", inc, wspd, phi) * np.exp(rng.normal(0, 0.05, (h, w)))
    s_cr = oracle.gmf_eval("gmf_s1_v2", inc, wspd) * np.exp(rng.normal(0, 0.05, (h, w)))
    anc = (wspd + rng.normal(0, 2, (h, w))) * np.exp(1j * np.deg2rad(phi + rng.normal(0, 20, (h, w))))
    nesz = 10 ** ((-33.0 + 0.1 * (inc - 20) + rng.normal(0, 0.2, (h, w))) / 10)
    nesz[rng.random((h, w)) < 0.01] = np.nan
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    model = ("gmf_cmod5n", "gmf_s1_v2")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # device chain
        flat_d = ws.nesz_flattening(t(nesz), t(inc))
        dsig_d = ws.get_dsig("gmf_s1_v2", t(inc), t(s_cr), flat_d)
        co_d, dual_d = ws.invert_from_model(t(inc), t(s_co), t(s_cr), ancillary_wind=t(anc), dsig_cr=dsig_d, model=model, **KW)
        assert all(x.is_cuda for x in (flat_d, dsig_d, co_d, dual_d))
        # host chain through the oracles
        flat_o = od.nesz_flattening(nesz, inc)
        dsig_o = od.get_dsig("gmf_s1_v2", inc, s_cr, flat_o)
        o_co, o_du, _, _ = oracle_dual(inc, s_co, s_cr, anc, dsig_cr=dsig_o)
        with np.errstate(invalid="ignore"):
            merged = np.where((np.abs(o_co) < 5) | (np.abs(o_du) < 5), o_co, o_du)
    np.testing.assert_allclose(flat_d.cpu().numpy(), flat_o, rtol=1e-10)
    np.testing.assert_allclose(dsig_d.cpu().numpy(), dsig_o, rtol=1e-9)
    wind_close(co_d.cpu().numpy(), o_co)
    wind_close(dual_d.cpu().numpy(), merged)
    # detrend (labelled inputs: the reference's sigma0_detrend needs .isel) -> local_gradients
    lab = lambda a: _xr.make_dataarray(a, dims=("line", "sample"), coords={"line": np.arange(h), "sample": np.arange(w)})
    det = xsarsea_b200.sigma0_detrend(lab(s_co), lab(inc), model="gmf_cmod5n")
    prof = oracle.gmf_eval("gmf_cmod5n", inc[0], np.full(w, 10.0), np.full(w, 45.0))
    det_o = s_co / (prof / np.nanmean(prof))
    np.testing.assert_allclose(np.asarray(det.data), det_o, rtol=1e-12)
    det_t = xsarsea_b200.sigma0_detrend(t(s_co), t(inc), model="gmf_cmod5n")      # CUDA tensors in -> CUDA tensor out
    assert det_t.is_cuda and np.array_equal(det_t.cpu().numpy(), np.asarray(det.data), equal_nan=True)
    ds = xsarsea_b200.gradients.local_gradients(det)
    g2, g3, c = og.local_gradients(det_o)
    for got, want in ((ds.G2, g2), (ds.G3, g3), (ds.c, c)):
        got = np.asarray(got.data)
        assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
