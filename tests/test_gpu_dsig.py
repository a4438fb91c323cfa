"""GPU parity of the dsig_cr pre-processors (xs_dsig, xs_dsig_wspd, xs_nesz_flatten through
the C ABI and through the reference-named API) against the golden outputs of the reference's own windspeed/utils.py and
against oracle/dsig.py on larger seeded inputs.

Tolerances (floating point; stated here as the contract):
  * element-wise formulas: rtol 2e-13 -- same FP64 operations, device exp/pow differ from glibc by <= 2 ulp and the
    exponents (8, 4, c(inc) <= 2.9) amplify an ulp of the ratio by at most that factor;
  * nesz_flattening: rtol 1e-10 -- np.polyfit solves the least-squares problem by SVD of the scaled Vandermonde matrix,
    the device by centred normal equations in FP64 (condition ~1e2-1e3), different summation order in the column means.
NaN positions must match exactly."""
import warnings

import numpy as np
import pytest

from oracle import dsig as od

pytestmark = pytest.mark.gpu
RTOL_ELEM = 2e-13
RTOL_FLAT = 1e-10


@pytest.fixture(scope="module")
def ws():
    import torch

    assert torch.cuda.is_available()
    from xsarsea_b200 import windspeed

    return windspeed


def close(got, want, rtol):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape and got.dtype == np.float64
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=rtol, atol=0)


def test_get_dsig_vs_reference_golden(ws, golden):
    g = golden("dsig_utils")
    for name in ("gmf_s1_v2", "gmf_rs2_v2", "nc_lut_cmodms1ahw", "sarwing_lut_cmodms1ahw"):
        close(ws.get_dsig(name, g["inc"], g["sigma0_cr"], g["nesz_cr"]), g["dsig__" + name], RTOL_ELEM)


def test_get_dsig_wspd_vs_reference_golden(ws, golden):
    g = golden("dsig_utils")
    for name in ("dsig_wspd_rs2_v3", "dsig_wspd_s1_ew_rec_v3", "dsig_wspd_rcm_v3"):
        got = ws.get_dsig_wspd(name, g["u_crosspol"], g["snr_cr"])
        want = g["wspd__" + name]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        # the weight is a product of two sigmoids: absolute tolerance at the 1-ulp-of-1 level plus the relative one
        np.testing.assert_allclose(got[ok], want[ok], rtol=RTOL_ELEM, atol=1e-300)
        assert (got[ok] >= 0).all() and (got[ok] <= 1).all()


def test_nesz_flattening_vs_reference_golden(ws, golden):
    g = golden("dsig_utils")
    close(ws.nesz_flattening(g["noise"], g["inc2d"]), g["noise_flat"], RTOL_FLAT)
    flat = ws.nesz_flattening(np.full((3, 16), np.nan), g["inc2d"][:3, :16].copy())
    assert flat.shape == (3, 16) and np.isnan(flat).all()


def test_containers_scalars_broadcast_and_float32(ws):
    import torch

    from xsarsea_b200 import _xr

    rng = np.random.default_rng(5)
    inc = rng.uniform(20, 45, (7, 33))
    s = 10 ** rng.uniform(-4, -1, (7, 33))
    # scalar nesz broadcasts (the usual call: a flattened NESZ raster or one number)
    close(ws.get_dsig("nc_lut_cmodms1ahw", inc, s, 10 ** -3.2), od.get_dsig("nc_lut_cmodms1ahw", inc, s, 10 ** -3.2),
          RTOL_ELEM)
    # 0-d in -> numpy scalar out
    v = ws.get_dsig("gmf_s1_v2", 33.0, 2e-3, 6e-4)
    assert np.ndim(v) == 0 and abs(v / od.get_dsig("gmf_s1_v2", 33.0, 2e-3, 6e-4) - 1) < RTOL_ELEM
    # labelled in -> labelled out with the template's dims
    lab = _xr.make_dataarray(s, dims=("line", "sample"), coords={"line": np.arange(7), "sample": np.arange(33)})
    r = ws.get_dsig("gmf_rs2_v2", inc, lab, np.full_like(s, 1e-3))
    assert _xr.is_labelled(r) and tuple(r.dims) == ("line", "sample")
    close(np.asarray(r.data), od.get_dsig("gmf_rs2_v2", inc, s, 1e-3), RTOL_ELEM)
    # float32 rasters are promoted on load; result is float64 of the promoted values
    s32, n32 = s.astype(np.float32), np.full(s.shape, 1e-3, np.float32)
    close(ws.get_dsig("gmf_rs2_v2", inc.astype(np.float32), s32, n32),
          od.get_dsig("gmf_rs2_v2", None, s32.astype(np.float64), n32.astype(np.float64)), RTOL_ELEM)
    # CUDA tensors in -> CUDA tensor out (device-resident pipeline)
    t = ws.get_dsig("nc_lut_cmodms1ahw", torch.from_numpy(inc).cuda(), torch.from_numpy(s).cuda(), 10 ** -3.2)
    assert t.is_cuda and t.dtype == torch.float64
    close(t.cpu().numpy(), od.get_dsig("nc_lut_cmodms1ahw", inc, s, 10 ** -3.2), RTOL_ELEM)
    # empty
    assert ws.get_dsig("gmf_rs2_v2", np.zeros((0, 4)), np.zeros((0, 4)), np.zeros((0, 4))).shape == (0, 4)
    assert ws.nesz_flattening(np.zeros((0, 4)), np.zeros((0, 4))).shape == (0, 4)


@pytest.mark.parametrize("shape", [(1, 1), (1, 257), (65, 1), (300, 1000), (2049, 515)])
def test_nesz_flattening_shapes_vs_oracle(ws, shape):
    h, w = shape
    rng = np.random.default_rng(h * 7919 + w)
    inc = np.broadcast_to(np.linspace(19.0, 47.0, w) if w > 1 else np.array([30.0]), (h, w)) + rng.normal(0, 1e-3, (h, w))
    noise = 10 ** ((-30.0 - 0.1 * (inc - 19.0) + rng.normal(0, 0.2, (h, w))) / 10.0)
    noise[rng.random((h, w)) < 0.03] = np.nan
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        want = od.nesz_flattening(noise, inc)
    got = ws.nesz_flattening(noise, inc)
    if w == 1:
        # a single abscissa: np.polyfit is rank deficient (minimum-norm solution); the prediction at that abscissa is
        # the mean of the ordinates either way
        close(got, want, 1e-9)
    else:
        close(got, want, RTOL_FLAT)


def test_exactly_log_linear_profile(ws):
    """Property: a noise profile that is exactly linear in dB is reproduced shifted by -1 dB (utils.py:158)."""
    inc = np.broadcast_to(np.linspace(20, 45, 640), (50, 640)).copy()
    noise = 10 ** ((-30 + 0.2 * inc) / 10)
    noise[1, 5] = np.nan
    flat = ws.nesz_flattening(noise, inc)
    np.testing.assert_allclose(10 * np.log10(flat), -31 + 0.2 * inc, atol=1e-10)


def test_dsig_raster_feeds_the_inversion_on_device(ws, golden):
    """get_dsig(CUDA tensors) -> InversionPlan.invert without leaving the device gives the same indices as the raster
    computed by the oracle on the host."""
    import torch

    import oracle
    from xsarsea_b200 import _device as D

    g = golden("inv_small")
    rng = np.random.default_rng(11)
    n = 4000
    gi, gwc = g["inc_grid"], g["wspd_cr_grid"]
    inc = rng.uniform(gi[0], gi[-1], n)
    cr_db = g["cr_lut_db"]
    s_cr = 10 ** (cr_db[rng.integers(0, cr_db.shape[0], n), rng.integers(0, cr_db.shape[1], n)] / 10) * np.exp(
        rng.normal(0, 0.05, n))
    nesz = np.full(n, 10 ** -3.2)
    dsig_host = od.get_dsig("nc_lut_cmodms1ahw", inc, s_cr, nesz)
    dsig_dev = ws.get_dsig("nc_lut_cmodms1ahw", *(torch.from_numpy(a).cuda() for a in (inc, s_cr, nesz)))
    plan = D.InversionPlan(cr=(D.to_device(cr_db), gi, gwc))
    t = lambda a: torch.from_numpy(a).cuda()
    _, o1, _, i1 = plan.invert(t(inc), None, t(s_cr), dsig_dev, None, cr_abs=True, want_idx=True)
    _, o2, _, i2 = plan.invert(t(inc), None, t(s_cr), t(dsig_host), None, cr_abs=True, want_idx=True)
    # dsig differs by <= 2e-13 relative between the two: the argmin can only move at an exact cost tie
    assert (i1 != i2).float().mean().item() < 1e-3
    assert torch.equal(o1[i1 == i2], o2[i1 == i2])
    plan.close()
