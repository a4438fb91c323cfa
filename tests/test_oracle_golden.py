"""The oracle (oracle/) against the golden vectors produced by the reference's own numba kernels
(tests/golden/make_golden.py).  CPU only.  Bar: bit-exact (the oracle restates the same FP64 operations
in the same order and calls the same libm)."""
import numpy as np
import pytest

import oracle
from oracle import lut as olut
from oracle import numba_port


def same(a, b):
    """Equal where both finite/inf, NaN positions equal (payload/sign of NaN not compared: SURVEY A.5)."""
    a, b = np.asarray(a), np.asarray(b)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(a[~na], b[~nb])


@pytest.mark.parametrize("name", list(oracle.MODEL_IDS))
def test_gmf_points_bit_exact(golden, name):
    g = golden("gmf_points")
    got = oracle.gmf_eval(name, g[name + "/inc"], g[name + "/wspd"], g[name + "/phi"])
    assert np.array_equal(got, g[name + "/sigma0"])
    phi = g.get(name + "/lut_phi")
    lut = oracle.lut_build(name, g[name + "/lut_inc"], g[name + "/lut_wspd"], phi)
    assert np.array_equal(lut, g[name + "/lut"])


def test_known_answers():
    # value of the reference's own code, and the docstring example of gmfs.py:60-63
    assert oracle.gmf_scalar("gmf_cmod5n", 35.0, 10.0, 45.0) == 0.05376709128885202


def _run(d, co=True, cr=True, **over):
    kw = {}
    if co:
        kw.update(co_lut=d["co_lut_db"], inc_grid=d["inc_grid"], wspd_grid=d["wspd_grid"], phi_grid=d["phi_grid"])
    if cr:
        kw.update(cr_lut=d["cr_lut_db"], inc_cr_grid=d["inc_grid"], wspd_cr_grid=d["wspd_cr_grid"])
    kw.update(over)
    return kw


@pytest.mark.parametrize("case", ["inv_small", "inv_slabs"])
def test_inversion_dual_bit_exact(golden, case):
    d = golden(case)
    co, dual, ic, ix = oracle.invert(d["inc"], d["s0_co_db"], d["s0_cr_db"], d["dsig_cr"], d["anc"], **_run(d))
    assert same(co, d["out_co"])
    assert same(dual, d["out_cr"])
    # indices agree with the returned grid values
    ok = ic >= 0
    n_phi = d["phi_grid"].size
    assert np.array_equal(np.abs(d["out_co"][ok]).round(9), d["wspd_grid"][ic[ok] // n_phi].round(9))
    okx = ix >= 0
    assert np.allclose(np.abs(d["out_cr"][okx]), d["wspd_cr_grid"][ix[okx]], rtol=1e-14, atol=0)


def test_inversion_mono_variants(golden):
    d = golden("inv_small")
    nanr = np.full(d["inc"].shape, np.nan)
    co, dual, _, _ = oracle.invert(d["inc"], d["s0_co_db"], nanr, 0.1, d["anc"], **_run(d, cr=False))
    assert same(co, d["co_only"]) and same(dual, d["co_only_cr"])
    co, dual, _, _ = oracle.invert(d["inc"], nanr, d["s0_cr_db"], d["dsig_cr"], nanr + 0j, **_run(d, co=False))
    assert same(co, d["cr_only_co"]) and same(dual, d["cr_only"])


def test_inversion_nan_lut_and_unmirrored_phi(golden):
    d = golden("inv_ifr2")
    nanr = np.full(d["inc"].shape, np.nan)
    assert np.isnan(d["co_lut_db"]).any()
    co, dual, ic, _ = oracle.invert(d["inc"], d["s0_co_db"], nanr, 0.1, d["anc"], co_lut=d["co_lut_db"],
                                    inc_grid=d["inc_grid"], wspd_grid=d["wspd_grid"], phi_grid=d["phi_grid"])
    assert same(co, d["out_co"]) and same(dual, d["out_cr"])
    assert not oracle.phi_is_180(d["phi_grid2"])
    co2, _, _, _ = oracle.invert(d["inc"], d["s0_co_db"], nanr, 0.1, d["anc"], co_lut=d["co2_lut_db"],
                                 inc_grid=d["inc_grid"], wspd_grid=d["wspd_grid"], phi_grid=d["phi_grid2"],
                                 dsig_co=float(d["dsig_co2"]))
    assert same(co2, d["out2_co"])


def test_numba_port_equals_c_oracle(golden):
    d = golden("inv_small")
    n = 500
    sl = slice(0, n)
    f = numba_port.make_inverter(d["co_lut_db"], d["inc_grid"], d["wspd_grid"], d["phi_grid"], d["cr_lut_db"],
                                 d["inc_grid"], d["wspd_cr_grid"], parallel=False)
    with np.errstate(all="ignore"):
        co, dual = f(d["inc"][sl], d["s0_co_db"][sl], d["s0_cr_db"][sl], d["dsig_cr"][sl], d["anc"][sl])
    assert same(co, d["out_co"][sl]) and same(dual, d["out_cr"][sl])


def test_interp_matches_scipy_interp1d():
    from scipy.interpolate import interp1d

    lut, (gi, gw, gp), res, steps = olut.raw_lut("gmf_cmod5n", inc_step_lr=5.0, wspd_step_lr=2.0, phi_step_lr=15.0)
    ti, tw, tp = olut.grid([16.0, 66.0], 0.5), olut.grid([0.2, 50.0], 0.3), olut.grid([0.0, 180.0], 4.0)
    want = lut
    for ax, (xs, xd) in enumerate([(gi, ti), (gw, tw), (gp, tp)]):
        want = interp1d(xs, want, kind="linear", axis=ax, bounds_error=True)(xd)
    got = lut
    for ax, (xs, xd) in enumerate([(gi, ti), (gw, tw), (gp, tp)]):
        got = oracle.interp_axis(got, ax, xs, xd)
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        oracle.interp_axis(lut, 0, gi, np.array([15.0, 20.0]))


def test_to_lut_decision_table():
    # SURVEY appendix A.1
    lut, (gi, gw, gp) = olut.to_lut("gmf_cmod5n", units="dB", inc_step_lr=5.0, wspd_step_lr=1.0, phi_step_lr=10.0,
                                    inc_step=1.0, wspd_step=0.5, phi_step=5.0)
    assert lut.shape == (51, 101, 37)           # low-res evaluated then interpolated to the "high" steps
    hi, (hi_i, hi_w, hi_p) = olut.to_lut("gmf_cmod5n", units="dB", resolution="high", inc_step=1.0, wspd_step=0.5,
                                         phi_step=5.0)
    assert hi.shape == lut.shape and not np.array_equal(hi, lut)   # direct evaluation differs from interpolation
    assert 0.01 < np.abs(hi - lut).max() < 6.0
    lo, grids = olut.to_lut("gmf_cmod5n", resolution="low")
    assert lo.shape == (51, 250, 73)
    x, gx = olut.to_lut("gmf_s1_v2", units="dB", resolution=None)
    assert x.shape == (501, 771) and gx[2] is None
    np.testing.assert_array_equal(gx[1], np.linspace(3.0, 80.0, 771))


def test_detrend_oracle():
    rng = np.random.default_rng(0)
    s0 = rng.uniform(0.01, 0.2, (7, 33))
    inc = np.linspace(30, 45, 33)
    prof = oracle.gmf_eval("gmf_cmod5n", inc, 10.0, 45.0)
    prof[3] = np.nan
    want = s0 / (prof / np.nanmean(prof))
    got = oracle.detrend(s0, prof)
    np.testing.assert_allclose(got, want, rtol=1e-14, equal_nan=True)


def test_dsig_oracle_bit_exact_against_reference_outputs(golden):
    """oracle/dsig.py against outputs of the reference's own windspeed/utils.py (tests/golden/make_golden.py::dsig_utils):
    same numpy calls in the same order -> bit-identical, NaN positions included."""
    import warnings

    from oracle import dsig as od

    g = golden("dsig_utils")
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for key in g:
            if key.startswith("dsig__"):
                got = od.get_dsig(key[6:], g["inc"], g["sigma0_cr"], g["nesz_cr"])
            elif key.startswith("wspd__"):
                got = od.get_dsig_wspd(key[6:], g["u_crosspol"], g["snr_cr"])
            else:
                continue
            assert np.array_equal(got, g[key], equal_nan=True), key
        assert np.array_equal(od.nesz_flattening(g["noise"], g["inc2d"]), g["noise_flat"], equal_nan=True)
        flat = od.nesz_flattening(np.full((3, 16), np.nan), g["inc2d"][:3, :16])
        assert np.isnan(flat).all() and np.isnan(g["noise_allnan_flat"]).all()
    with pytest.raises(ValueError):
        od.get_dsig("other", 1.0, 1.0, 1.0)


def test_gradients_oracle_against_golden_and_an_independent_restatement(golden):
    """oracle/gradients.py reproduces tests/golden/gradients.npz bit for bit (same cv2 / scipy calls), and agrees with
    a restatement that uses neither library (explicit Scharr stencil with reflect-101 borders, separable binomial
    filters with symmetric borders) to rounding -- which pins the border rules and the filter taps independently."""
    from oracle import gradients as og

    g = golden("gradients")
    for tag in ("odd", "even", "tiny", "thin"):
        img = g[tag + "/image"]
        g2, g3, c = og.local_gradients(img)
        assert same(g2, g[tag + "/G2"]) and same(g3, g[tag + "/G3"]) and same(c, g[tag + "/c"]), tag

    img = g["even/image"]
    p = np.pad(img, 1, mode="reflect")                       # cv2 BORDER_REFLECT_101
    dx, dy = p[:, 2:] - p[:, :-2], p[2:, :] - p[:-2, :]
    gr = 10 * dx[1:-1] + 3 * (dx[:-2] + dx[2:])
    gi = 10 * dy[:, 1:-1] + 3 * (dy[:, :-2] + dy[:, 2:])
    z = (gr + 1j * gi) ** 2

    def binom(a, taps):
        r = len(taps) // 2
        for ax in (0, 1):
            q = np.pad(a, [(r, r) if k == ax else (0, 0) for k in (0, 1)], mode="symmetric")   # scipy 'symm'
            a = sum(t * np.take(q, np.arange(i, i + a.shape[ax]), axis=ax) for i, t in enumerate(taps))
        return a

    def r2(a):
        a = binom(a, np.array([1, 4, 6, 4, 1]) / 16)
        a = a[:a.shape[0] // 2 * 2, :a.shape[1] // 2 * 2].reshape(a.shape[0] // 2, 2, a.shape[1] // 2, 2).mean(axis=(1, 3))
        return binom(a, np.array([1, 2, 1]) / 4)

    grad2, grad3 = r2(z), r2(np.abs(z))
    np.testing.assert_allclose(np.sqrt(grad2), g["even/G2"], rtol=0, atol=1e-12 * np.abs(g["even/G2"]).max())
    np.testing.assert_allclose(grad3, g["even/G3"], rtol=1e-12)
    cq = np.abs(grad2) / (grad3 + 1e-5)
    np.testing.assert_allclose(np.where(cq <= 1, cq, 0), g["even/c"], rtol=0, atol=1e-12)


def test_nesz_line_fit_algorithm_matches_polyfit():
    """The per-line fit of xs_nesz_flatten (one pass, sums shifted by the line's first finite point, centred normal
    equations; equal abscissae -> the minimum-norm solution a = ybar/(2x), b = ybar/2 that np.polyfit's scaled lstsq
    returns) transcribed to numpy, against np.polyfit itself -- pins the algorithm the kernel's comments state."""
    import warnings

    rng = np.random.default_rng(9)

    def device_fit(x, y):
        x0, y0 = x[0], y[0]
        dx, dy = x - x0, y - y0
        n, sx, sy, sxx, sxy = x.size, dx.sum(), dy.sum(), (dx * dx).sum(), (dx * dy).sum()
        mx, my = sx / n, sy / n
        cxx, cxy = sxx - sx * mx, sxy - sx * my
        if cxx > 0 and np.isfinite(cxx):
            a = cxy / cxx
            return a, (y0 + my) - a * (x0 + mx)
        xb, yb = x0 + mx, y0 + my
        return yb / (2 * xb), yb / 2

    for n in (2, 3, 50, 5000):
        x = np.sort(rng.uniform(19, 47, n))
        y = -30 + 0.15 * x + rng.normal(0, 0.3, n)
        a, b = device_fit(x, y)
        pa, pb = np.polyfit(x, y, 1)
        np.testing.assert_allclose([a, b], [pa, pb], rtol=1e-9)
        np.testing.assert_allclose(a * x + b, pa * x + pb, rtol=0, atol=1e-10)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # RankWarning: that is the point
        for n in (1, 4):
            x, y = np.full(n, 33.25), rng.uniform(-40, -20, n)
            a, b = device_fit(x, y)
            pa, pb = np.polyfit(x, y, 1)
            np.testing.assert_allclose([a, b], [pa, pb], rtol=1e-12)
