"""CPU check of the error bounds of the centred FP32 scan (`k_scan_co`, both modes).

The kernel's exactness rests on `E >= |J''_fp32 - J''_exact|` for every candidate that can still win.  This test
re-creates the kernel's FP32 operation sequence in numpy (every FP32 operation is evaluated exactly in float64 -- a
product of two floats and one addend fit -- and rounded once to float32), evaluates the exact shifted cost in extended
precision, mirrors the kernel's formula for E, and checks on seeded adversarial pixels (sigma0 on LUT nodes, ancillary
wind on candidates, strong winds, wide spreads of sigma0 inside a warp) that
  * every candidate with |L/dsig - s/dsig| <= D has an FP32 cost within E of the exact one,
  * the exact argmin is inside the band {J_fp32 <= m32 + 2E}, i.e. the FP64 refinement sees it.
No GPU and nothing of the product is involved: this pins the mathematics the kernel's comments and DESIGN.md state."""
import numpy as np
import pytest

import oracle

U = 2.0 ** -24
F32 = np.float32


def f32(x):
    return np.asarray(x, dtype=np.float64).astype(F32).astype(np.float64)   # one rounding to float32, held as float64


def fma32(a, b, c):
    """fl32(a*b + c) for float32-valued a, b, c (a*b is exact in float64; the sum is rounded to float64 first, which can
    differ from a true fused rounding by a double-rounding ulp of float32 in ~1e-9 of the cases -- irrelevant for a
    magnitude check)."""
    return f32(a * b + c)


@pytest.fixture(scope="module")
def slabs():
    gi = np.array([20.0, 30.0, 37.5, 45.0, 60.0])
    gw, gp = np.linspace(0.2, 50, 499), np.linspace(0, 180, 181)
    lut_db = 10 * np.log10(oracle.lut_build("gmf_cmod5n", gi, gw, gp) + 1e-15)
    return lut_db, gw, gp


def kernel_E(m32, sc_abs, amag, lmax, wmax, shared=False):
    """Mirror of the band section of k_scan_co (float arithmetic there, float64 here): bound of the full centred form."""
    A, W = amag, wmax * 1.0000002
    T = W * A + 0.25 * W * W
    SC = sc_abs
    R0 = np.sqrt(np.maximum(m32, 0.0) + 0.25 * A * A + 1.0)
    D = np.sqrt(np.maximum(m32 + SC * SC * 1.0000002, 0.0) + 0.25 * A * A + 1.0 + (2.0 * SC * R0 * 1.000001 if shared else 0.0))
    Lam = D + SC
    E = 5.9604645e-8 * 1.5 * (2 * Lam * lmax + 2 * SC * lmax + 4 * Lam * Lam + 6 * SC * Lam + SC * SC + D * D + 4 * T)
    return E, D


@pytest.mark.parametrize("spread_db", [0.0, 0.3, 3.0, 15.0])
def test_centred_scan_error_bound(slabs, spread_db):
    lut_db, gw, gp = slabs
    dsig = 0.1
    rng = np.random.default_rng(int(spread_db * 10) + 1)
    cphi, sphi = np.cos(np.radians(gp)), np.sin(np.radians(gp))
    nwh32, w2q32 = f32(-0.5 * gw), f32(0.25 * gw * gw)
    worst = 0.0
    for b in range(lut_db.shape[0]):
        Ls = lut_db[b] / dsig                       # [n_wspd, n_phi] float64 (the kernel: (float)(L/dsig))
        L32 = f32(Ls)
        lmax = np.abs(L32).max() * 1.0000002
        for trial in range(24):
            kind = trial % 6
            iw, ip = rng.integers(0, gw.size), rng.integers(0, gp.size)
            s = lut_db[b, iw, ip] if kind in (0, 1) else rng.uniform(-35.0, 5.0)      # on a node / anywhere
            if kind == 4:
                s = rng.uniform(-60.0, 20.0)                                          # outside the LUT range
            a_w = gw[iw] if kind in (0, 2) else rng.uniform(0.0, 60.0)
            a_p = np.radians(gp[ip]) if kind in (0, 2) else rng.uniform(0.0, np.pi)
            qa, qb = a_w * np.cos(a_p), a_w * np.sin(a_p)                             # qb = |Im| (mirrored phi grid)
            if kind == 5:
                qa = qb = 0.0
            # warp centre: somewhere within +- spread of the pixel's own sigma0 (the kernel: middle of the warp's 8 pixels)
            cs = float(F32((s + rng.uniform(-spread_db, spread_db)) / dsig))
            sc = s / dsig - cs
            k32 = float(F32(-2.0 * sc))
            g64 = qa * cphi + qb * sphi
            g32 = f32(g64)
            # the kernel's operation sequence
            Lc = f32(L32 - cs)
            M = fma32(Lc, Lc, w2q32[:, None])
            aa = fma32(k32, Lc, M)
            J32 = fma32(nwh32[:, None], g32[None, :], aa)
            # exact shifted cost J'' = J' - sc^2 in extended precision
            lam = Ls.astype(np.longdouble) - np.longdouble(cs)
            Jx = lam * lam - 2 * np.longdouble(sc) * lam + (np.longdouble(0.25) * gw * gw)[:, None] - (
                np.longdouble(0.5) * gw)[:, None] * g64.astype(np.longdouble)[None, :]
            err = np.abs(J32 - Jx.astype(np.float64))
            m32 = J32.min()
            amag = float(F32(np.hypot(qa, qb))) * 1.0000002
            E, D = kernel_E(m32, abs(sc) * 1.0000002, amag, lmax, gw.max())
            if not (E < 0.25):          # the kernel sends such pixels to the exhaustive FP64 kernel
                continue
            d = np.abs(Ls - s / dsig)
            can_win = d <= D
            assert can_win.any()
            assert err[can_win].max() <= E, (b, trial, kind, err[can_win].max(), E)
            worst = max(worst, err[can_win].max() / E)
            # the exact argmin (first minimum) is inside the band the refinement examines, and candidates outside the
            # |d| <= D cone are outside the band
            ix = np.unravel_index(np.argmin(Jx), Jx.shape)
            assert J32[ix] <= m32 + 2 * E
            assert (J32[~can_win] > m32 + 2 * E).all()
    assert 0 < worst <= 1.0


def test_shared_sigma0_mode_band_and_second_filter(slabs):
    """Shared-sigma0 mode of k_scan_co: the scanned cost leaves k lambda out (J_a = fma(-w/2, g, M)) and the band is widened
    by 2 |sigma| Lam; k_refine_easy then keeps only the members within 2 efp of the smallest FULL FP32 cost.  Checked: the
    exact argmin (first minimum, ties included) is a member of the wide band, survives the second filter, and every
    candidate of the wide band has |lambda| <= Lam (which is what both bounds assume)."""
    lut_db, gw, gp = slabs
    dsig = 0.1
    rng = np.random.default_rng(21)
    cphi, sphi = np.cos(np.radians(gp)), np.sin(np.radians(gp))
    nwh32, w2q32 = f32(-0.5 * gw), f32(0.25 * gw * gw)
    n_checked = n_outside = 0
    for b in range(lut_db.shape[0]):
        Ls = lut_db[b] / dsig
        L32 = f32(Ls)
        lmax = np.abs(L32).max() * 1.0000002
        for trial in range(30):
            kind = trial % 5
            iw, ip = rng.integers(0, gw.size), rng.integers(0, gp.size)
            s = lut_db[b, iw, ip] if kind == 0 else rng.uniform(-35.0, 5.0)
            if kind == 4:
                s = rng.choice([rng.uniform(-60.0, -40.0), rng.uniform(8.0, 25.0)])   # outside the LUT range: large lambda, small spread
            a_w, a_p = (gw[iw], np.radians(gp[ip])) if kind == 0 else (rng.uniform(0.0, 40.0), rng.uniform(0.0, np.pi))
            qa, qb = a_w * np.cos(a_p), a_w * np.sin(a_p)
            sc = rng.uniform(-1.0, 1.0) * 10.0 ** rng.uniform(-4, -1.5)       # |sigma| of a sorted list: 1e-4 .. 3e-2
            cs = float(F32(s / dsig - sc))
            sc = s / dsig - cs
            g64 = qa * cphi + qb * sphi
            g32 = f32(g64)
            Lc = f32(L32 - cs)
            M = fma32(Lc, Lc, w2q32[:, None])
            Ja = fma32(nwh32[:, None], g32[None, :], M)                       # the shared-mode scan
            kfull = float(F32(-2.0 * sc))
            Jf = fma32(nwh32[:, None], g32[None, :], fma32(kfull, Lc, M))     # the refinement's full FP32 cost
            lam = Ls.astype(np.longdouble) - np.longdouble(cs)
            Jx = lam * lam - 2 * np.longdouble(sc) * lam + (np.longdouble(0.25) * gw * gw)[:, None] - (
                np.longdouble(0.5) * gw)[:, None] * g64.astype(np.longdouble)[None, :]
            m32 = Ja.min()
            amag = float(F32(np.hypot(qa, qb))) * 1.000001
            SC = abs(sc) * 1.0000002
            efp, D = kernel_E(m32, SC, amag, lmax, gw.max(), shared=True)
            Lam = D + SC
            sp = float(F32(s / dsig))
            dist = max(float(L32.min()) - sp, sp - float(L32.max()))      # sigma0 outside the slab's value range?
            dmin = max(dist * 0.999999 - 1e-6 * lmax - 1e-6 * abs(sp), 0.0)
            dl = max(Lam - dmin, 0.0) if dmin > 0.0 else 2.0 * Lam
            E = efp + SC * dl * 1.000001
            if not (E < 0.25):      # the kernel sends such pixels to the exhaustive FP64 kernel
                continue
            band = Ja <= m32 + 2 * E
            ix = np.unravel_index(np.argmin(Jx), Jx.shape)
            assert band[ix], (b, trial, kind)
            assert (np.abs(Lc)[band] <= Lam).all()
            assert np.abs(Jf - Jx.astype(np.float64))[band].max() <= efp
            assert Jf[ix] <= Jf[band].min() + 2 * efp
            ties = Jx == Jx[ix]                                               # every exact tie must survive too
            assert (band & (Jf <= Jf[band].min() + 2 * efp))[ties].all()
            n_checked += 1
            n_outside += dmin > 0.0
    assert n_checked > 100 and n_outside > 15
