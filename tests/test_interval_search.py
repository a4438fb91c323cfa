"""CPU check of the exact interval search of the cross-pol pass (`cross_interval_search`, xs_invert.cu).

A statement-by-statement Python transcription of the device function (float64 arithmetic is IEEE on both sides, no FMA
involved) against the brute-force first-minimum argmin of the reference's cost (windspeed.py:254-269) on seeded
adversarial rows: strictly increasing, plateaus (runs of equal LUT values -> runs of equal costs), two-valued rows,
sigma0 on nodes and midpoints, |wind_co| on nodes and midpoints, dsig from 1e-8 to 1e4, sigma0 far outside the row.
This pins the monotonicity argument (the candidate set is an index interval; its ends are found from guessed brackets --
inverse index of the row, uniform-grid guess -- verified and widened when wrong, and by galloping outwards from the valley
bottoms) independently of the GPU; the GPU test `test_cross_pol_filter_adversarial` pins the implementation."""
import math

import numpy as np


def cost(col, wg, s, dsig, mag, hc, w):
    ts = (col[w] - s) / dsig
    J = ts * ts
    if hc:
        tw = (wg[w] - mag) * 0.5
        J = J + tw * tw
    return J


NB = 1024   # kCrInvBuckets


def inverse_index(col):
    """k_build_cr_tables: inv[b] = first w with col[w] >= vlo + b / scale (b < NB), inv[NB] = n."""
    n = len(col)
    vlo, vhi = col[0], col[-1]
    usable = vhi > vlo and math.isfinite(vhi - vlo)
    scale = NB / (vhi - vlo) if usable else 0.0
    inv, w = [], 0
    for b in range(NB):
        if usable:
            edge = vlo + b / scale
            while w < n and col[w] < edge:
                w += 1
        inv.append(w if usable else 0)
    inv.append(n)
    return (vlo if usable else 0.0), scale, inv


def interval_search(col, wg, s, dsig, mag, hc, tables=None, wspd_uniform=None):
    n = len(col)
    num = lambda w: col[w] - s
    tw_at = lambda w: (wg[w] - mag) * 0.5

    def first_ge(lo, hi, ge):
        lo0, hi0 = lo, hi
        while lo < hi:
            mid = (lo + hi) >> 1
            if not ge(mid):
                lo = mid + 1
            else:
                hi = mid
        if (lo == lo0 and lo > 0 and ge(lo - 1)) or (lo == hi0 and lo < n and not ge(lo)):
            lo, hi = 0, n
            while lo < hi:
                mid = (lo + hi) >> 1
                if not ge(mid):
                    lo = mid + 1
                else:
                    hi = mid
        return lo

    vlo, scale, inv = tables if tables is not None else inverse_index(col)
    t = (s - vlo) * scale
    b = int(min(t, NB - 1)) if t > 0.0 else 0
    k = first_ge(inv[b], inv[b + 1], lambda w: not (num(w) < 0.0))
    if not hc:   # cross-pol only: decided on |L - s| of the nodes around the sign change (no division)
        xa = abs(num(k - 1)) if k > 0 else math.inf
        xb = abs(num(k)) if k < n else math.inf
        xm = min(xa, xb)
        if 1e-100 <= dsig <= 1e100 and 1e-100 * dsig <= xm <= 1e100 * dsig:
            if xa > xb * (1.0 + 4e-16):
                return k, 1
            if xa <= xb:
                if k - 1 == 0:
                    return 0, 1
                if abs(num(k - 2)) > xa * (1.0 + 4e-16):
                    return k - 1, 1
    m0 = math.inf
    if k < n:
        m0 = cost(col, wg, s, dsig, mag, hc, k)
    if k > 0:
        m0 = min(m0, cost(col, wg, s, dsig, mag, hc, k - 1))
    j = 0
    if hc:
        lo, hi = 0, n
        if wspd_uniform is not None:
            g0, inv_step = wspd_uniform
            t = (mag - g0) * inv_step
            g = 0 if t <= 0.0 else (n if t >= n else int(t))
            lo, hi = max(g - 1, 0), min(g + 2, n)
        j = first_ge(lo, hi, lambda w: not (tw_at(w) < 0.0))
        if j < n:
            m0 = min(m0, cost(col, wg, s, dsig, mag, hc, j))
        if j > 0:
            m0 = min(m0, cost(col, wg, s, dsig, mag, hc, j - 1))
    if not (m0 < math.inf):
        return -1
    q = math.sqrt(m0) * dsig
    cheap = 1e-290 <= m0 <= 1e290 and 1e-290 <= q <= 1e290
    q_hi, q_lo = q * (1.0 + 1e-14), q * (1.0 - 1e-14)

    def a_gt_m0(w):
        v = num(w)
        an = abs(v)
        if cheap and an > q_hi:
            return True
        if cheap and an < q_lo:
            return False
        t = v / dsig
        return t * t > m0

    def b_gt_m0(w):
        t = tw_at(w)
        return t * t > m0

    def left_end(c, gt):
        ok, bad, step, pos = c, -1, 1, c - 1
        while pos >= 0:
            if gt(pos):
                bad = pos
                break
            ok = pos
            pos -= step
            step <<= 1
        lo, hi = bad + 1, ok
        while lo < hi:
            mid = (lo + hi) >> 1
            if gt(mid):
                lo = mid + 1
            else:
                hi = mid
        return lo

    def right_end(c, gt):
        ok, bad, step, pos = c - 1, n, 1, c
        while pos < n:
            if gt(pos):
                bad = pos
                break
            ok = pos
            pos += step
            step <<= 1
        lo, hi = ok + 1, bad
        while lo < hi:
            mid = (lo + hi) >> 1
            if not gt(mid):
                lo = mid + 1
            else:
                hi = mid
        return lo

    first, last = left_end(k, a_gt_m0), right_end(k, a_gt_m0)
    if hc:
        first = max(first, left_end(j, b_gt_m0))
        last = min(last, right_end(j, b_gt_m0))
    best, res = math.inf, -1
    for w in range(first, last):
        J = cost(col, wg, s, dsig, mag, hc, w)
        if J < best:
            best, res = J, w
    return res, last - first


def test_interval_search_equals_brute_force():
    rng = np.random.default_rng(42)
    wg = np.linspace(3.0, 80.0, 771)
    base = -38.0 + 30.0 * (1 - np.exp(-wg / 25.0)) + 0.05 * wg          # a cross-pol-like monotone row (dB)
    rows = {
        "smooth": base,
        "plateau": np.round(base * 2) / 2,                               # runs of equal values
        "two_valued": np.where(wg < 40, -30.0, -20.0),
        "flat": np.full_like(wg, -25.0),
        "steep": np.sort(rng.uniform(-45, -5, wg.size)),                 # monotone, irregular steps
    }
    checked, widths = 0, []
    for name, col in rows.items():
        col = np.ascontiguousarray(col, dtype=np.float64)
        assert (np.diff(col) >= 0).all()
        for trial in range(1500):
            kind = trial % 6
            k = int(rng.integers(0, wg.size - 1))
            s = float(rng.uniform(-50, 0))
            if kind == 0:
                s = float(col[k])
            elif kind == 1:
                s = float(0.5 * (col[k] + col[k + 1]))
            elif kind == 2:
                s = float(rng.uniform(-120, 40))
            dsig = float(10.0 ** rng.uniform(-8, 4)) if kind != 3 else 0.1
            hc = bool(trial % 2)
            mag = float(rng.uniform(0, 90))
            if kind == 4:
                mag = float(wg[k])
            elif kind == 5:
                mag = float(0.5 * (wg[k] + wg[k + 1]))
            with np.errstate(all="ignore"):
                ts = (col - s) / dsig
                J = ts * ts
                if hc:
                    tw = (wg - mag) * 0.5
                    J = J + tw * tw
            want = int(np.argmin(J))                                      # first minimum, like the reference
            hint = (float(wg[0]), 1.0 / ((wg[-1] - wg[0]) / (len(wg) - 1))) if trial % 4 < 2 else None   # the uniform-grid guess
            got = interval_search(col, wg, s, dsig, mag, hc, wspd_uniform=hint)
            if got == -1:                                                 # every cost overflowed: the kernel falls back
                assert not np.isfinite(J).any() or not np.isfinite(J.min())
                continue
            assert got[0] == want, (name, trial, got, want, s, dsig, mag, hc)
            widths.append(got[1])
            checked += 1
    assert checked > 7000
    assert np.median(widths) <= 8          # a handful of candidates are evaluated instead of 771
