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


# ---- k_cross_only: the cross-pol-only argmin as a step function of sigma0 (linear domain, no log10 per pixel) ----------
PROBE = 6   # kStepProbe


def step_tables(col):
    """k_build_cr_tables: midpoints in dB and in the linear domain, and whether the row is eligible (cr_finite bit 2)."""
    n = len(col)
    ok = bool(np.all(np.isfinite(col)))
    sdb, slin = [0.0], [0.0]
    for w in range(1, n):
        a, b = col[w - 1], col[w]
        ok = ok and (b - a > 1e-9 * (abs(a) + abs(b) + 1.0))
        mid = 0.5 * a + 0.5 * b
        lin = 10.0 ** (mid * 0.1)
        ok = ok and abs(mid) < 3000.0 and 1e-290 < lin < 1e290
        sdb.append(mid)
        slin.append(lin)
    return ok, sdb, slin


def step_search(col, x, dsig, db, tables, steps):
    """Transcription of k_cross_only's per-pixel decision; -1 = left to k_cross."""
    n = len(col)
    vlo, vscale, inv = tables
    ok_row, sdb_t, slin_t = steps
    y = x if db else x + 1e-15
    if not (math.isfinite(y) and (db or y > 0.0) and 1e-100 <= dsig <= 1e100) or not ok_row:
        return -1
    row = sdb_t if db else slin_t
    if db:
        sdb = y
    else:
        yf = float(np.float32(y))
        # __log2f: float log2 with ~2^-22 absolute error; denormal / zero input -> -inf
        sdb = float(np.float32(3.0102999566) * np.float32(math.log2(yf))) if yf >= 1.18e-38 and math.isfinite(yf) else \
            (-math.inf if yf < 1.18e-38 else math.inf)
    t = (sdb - vlo) * vscale
    b = int(min(t, NB - 1)) if t > 0.0 else 0
    k_lo, k_hi = max(inv[b] - 1, 0), min(inv[b + 1], n - 1)
    w0, w1 = max(k_lo, 1), min(k_hi + 1, n - 1)
    if not (w1 - w0 < PROBE):
        return -1
    c_hi = c_lo = 0
    first_below, last_below = True, False
    for q in range(PROBE):
        on = w0 + q <= w1
        v = row[w0 + q] if on else math.inf
        below_hi = ((v + 1e-10) if db else v * (1.0 + 1e-11)) < y
        below_lo = ((v - 1e-10) if db else v * (1.0 - 1e-11)) < y
        c_hi += below_hi
        c_lo += below_lo
        if q == 0:
            first_below = below_hi
        if on and w0 + q == w1:
            last_below = below_lo
    ok_lo = w0 == 1 or first_below
    ok_hi = (k_hi == n - 1 or not last_below) if w1 == n - 1 else (not last_below)
    if w1 < w0:
        return 0
    if c_hi == c_lo and ok_lo and ok_hi:
        return w0 - 1 + c_hi
    return -1


def test_cross_only_step_function_equals_brute_force():
    """Whenever the step-function kernel settles a pixel, its index is np.argmin of the reference cost evaluated the
    reference's way (dB conversion included); sigma0 on / next to midpoints and nodes must either agree or be declined."""
    rng = np.random.default_rng(20261018)
    n_settled = n_declined = n_random = n_random_declined = 0
    for case in range(60):
        n = int(rng.choice([1, 2, 3, 17, 200, 771]))
        kind = case % 4
        if kind == 0:      # GMF-like: dB of a power law
            w = np.linspace(3, 80, n) if n > 1 else np.array([3.0])
            col = 10 * np.log10(1e-4 * (w / 10.0) ** rng.uniform(1.0, 2.5))
        elif kind == 1:    # irregular steps, some tiny
            col = np.cumsum(rng.choice([1e-7, 1e-3, 0.05, 2.0], n)) - 40.0
        elif kind == 2:    # a plateau: the row is not eligible at all
            col = np.sort(rng.uniform(-40, -10, n))
            if n > 2:
                col[n // 2] = col[n // 2 - 1]
        else:              # clustered nodes: many nodes in one bucket of the inverse index
            col = np.sort(np.concatenate([rng.uniform(-30.0, -29.99, n // 2), rng.uniform(-45, -5, n - n // 2)]))
        col = np.asarray(col, dtype=np.float64)
        tables = inverse_index(col)
        steps = step_tables(col)
        for db in (False, True):
            mids = 0.5 * col[:-1] + 0.5 * col[1:] if n > 1 else np.array([col[0]])
            base = np.concatenate([mids, col, rng.uniform(col[0] - 15, col[-1] + 15, 40)])
            is_random = np.concatenate([np.zeros(len(mids) + len(col), bool), np.ones(40, bool)])
            for s_t, rnd in zip(base, is_random):
                for rel in (0.0, 1e-16, -1e-16, 3e-13, -3e-13, 1e-9, -1e-9):
                    dsig = float(rng.choice([1e-8, 0.1, 1.0, 37.5, 1e4]))
                    if db:
                        x = float(s_t + rel * max(abs(s_t), 1.0))
                        s = x
                    else:
                        x = float(10.0 ** (s_t / 10.0) * (1.0 + rel) - 1e-15)
                        with np.errstate(invalid="ignore", divide="ignore"):
                            s = float(10.0 * np.log10(np.float64(x) + 1e-15))
                    got = step_search(col, x, dsig, db, tables, steps)
                    n_random += rnd and rel == 0.0
                    if got < 0:
                        n_declined += 1
                        n_random_declined += rnd and rel == 0.0
                        continue
                    ts = (col - s) / dsig
                    want = int(np.argmin(ts * ts))
                    assert got == want, (case, db, x, dsig, got, want)
                    n_settled += 1
    assert n_settled > 20000
    # pixels away from midpoints are settled (only rows that are not eligible, or clustered beyond the probe, decline)
    assert n_random_declined < 0.45 * n_random, (n_random_declined, n_random)
