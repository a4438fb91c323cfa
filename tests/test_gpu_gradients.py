"""GPU parity of the local-gradient stage (xs_local_gradients through the C ABI and through
`xsarsea_b200.gradients.local_gradients`) against the golden vectors (cv2.Scharr + scipy.signal.convolve2d, the
libraries the reference delegates to) and against oracle/gradients.py on larger seeded images.

Tolerance (floating point, stated as the contract): every output within 1e-12 x max|output| absolutely -- the device
evaluates the 5x5 and 3x3 binomial filters separably while scipy sums the 25 / 9 products directly, so results differ
by rounding (a few 1e-16 of the largest term); the complex square root and the quotient `c` amplify that by at most
~1e3 on the seeded images (documented near-tie: sqrt's branch cut -- a squared gradient on the negative real axis with
an imaginary part that is zero up to rounding -- cannot occur for noisy data and is excluded from the golden images).
NaN positions and the zeros of `c` (c > 1 or NaN -> 0) must match exactly."""
import numpy as np
import pytest

from oracle import gradients as og

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import torch

    assert torch.cuda.is_available()
    from xsarsea_b200 import gradients

    return gradients


def close(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape and got.dtype == want.dtype, what
    assert np.array_equal(np.isnan(got), np.isnan(want)), what
    ok = ~np.isnan(want)
    if ok.any():
        scale = np.abs(want[ok]).max()
        assert np.abs(got[ok] - want[ok]).max() <= 1e-12 * scale + 1e-300, (what, np.abs(got[ok] - want[ok]).max(), scale)


@pytest.mark.parametrize("tag", ["odd", "even", "tiny", "thin"])
def test_local_gradients_vs_golden(G, golden, tag):
    g = golden("gradients")
    ds = G.local_gradients(g[tag + "/image"])
    for name in ("G2", "G3", "c"):
        close(ds[name], g[f"{tag}/{name}"], f"{tag}/{name}")
    c = np.asarray(ds.c)
    assert ((c >= 0) & (c <= 1)).all()


@pytest.mark.parametrize("shape", [(16, 64), (17, 65), (250, 1000), (1001, 333), (2, 7), (1, 50), (40, 1)])
def test_shapes_vs_oracle(G, shape):
    h, w = shape
    rng = np.random.default_rng(h * 31 + w)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 0.05 + 0.02 * np.sin(0.3 * xx - 0.17 * yy) + 0.01 * rng.standard_normal((h, w))
    if h > 20 and w > 20:
        img[rng.random((h, w)) < 0.002] = np.nan
    ds = G.local_gradients(img)
    if h < 2 or w < 2:
        assert ds.G2.shape == (h // 2, w // 2) and ds.G2.size == 0
        return
    g2, g3, c = og.local_gradients(img)
    close(ds.G2, g2, "G2")
    close(ds.G3, g3, "G3")
    close(ds.c, c, "c")


def test_containers_and_float32(G):
    import torch

    from xsarsea_b200 import _xr

    rng = np.random.default_rng(2)
    img = rng.uniform(0.01, 0.2, (60, 90))
    lab = _xr.make_dataarray(img, dims=("line", "sample"), coords={"line": np.arange(60) * 10.0, "sample": np.arange(90) * 10.0})
    ds = G.local_gradients(lab)
    assert _xr.is_labelled(ds.G2) and tuple(ds.G2.dims) == ("line", "sample")
    # coarsen().mean() averages the coordinates of each pair
    np.testing.assert_array_equal(np.asarray(ds.G2.coords["line"] if isinstance(ds.G2.coords, dict) else ds.G2.line)[:3],
                                  [5.0, 25.0, 45.0])
    plain = G.local_gradients(img)
    np.testing.assert_array_equal(np.asarray(ds.G3.data), plain.G3)
    # CUDA tensor in -> CUDA tensors out
    t = G.local_gradients(torch.from_numpy(img).cuda())
    assert t.G2.is_cuda and t.G2.dtype == torch.complex128
    np.testing.assert_array_equal(t.c.cpu().numpy(), plain.c)
    # float32 image: promoted on load, same as the float64 path on the promoted values
    f = G.local_gradients(img.astype(np.float32))
    want = G.local_gradients(img.astype(np.float32).astype(np.float64))
    np.testing.assert_array_equal(f.G3, want.G3)
    with pytest.raises(ValueError):
        G.local_gradients(np.zeros((2, 3, 4)))


def test_properties_full_size(G):
    """Size-independent properties on an EW-sized raster (10 000 x 10 400): a pure plane wave gives gradients aligned
    with its wave vector (angle of G2 modulo pi) with quality c ~ 1 in the interior; a constant image gives zeros;
    scaling the image by k scales G2 by k and G3 by k**2."""
    import torch

    h, w = 10000, 10400
    yy = torch.arange(h, device="cuda", dtype=torch.float64)[:, None]
    xx = torch.arange(w, device="cuda", dtype=torch.float64)[None, :]
    kx, ky = 0.05, 0.03
    img = 1.0 + 0.5 * torch.sin(kx * xx + ky * yy)   # strong enough for the 1e-5 regulariser of c not to matter
    ds = G.local_gradients(img)
    g2 = ds.G2[50:-50, 50:-50]
    ang = torch.angle(g2)
    expect = np.arctan2(ky, kx)
    d = torch.remainder(ang - expect + np.pi / 2, np.pi) - np.pi / 2
    strong = g2.abs() > 0.2 * g2.abs().max()   # away from the crests, where the gradient vanishes
    assert d[strong].abs().max().item() < 0.05
    assert ds.c[50:-50, 50:-50][strong].min().item() > 0.99
    ds3 = G.local_gradients(3.0 * img)
    assert torch.allclose(ds3.G2, 3.0 * ds.G2, rtol=1e-9, atol=1e-15)
    assert torch.allclose(ds3.G3, 9.0 * ds.G3, rtol=1e-9, atol=1e-15)
    z = G.local_gradients(torch.full((512, 512), 0.3, device="cuda", dtype=torch.float64))
    assert z.G2.abs().max().item() == 0 and z.G3.abs().max().item() == 0 and z.c.max().item() == 0
