"""CPU check of the exact chunk pruning of the co-pol scan (`k_tile_plan`).

The scan skips a 16-row chunk of the slab for a pixel when a lower bound of the reference's cost over the whole chunk,
    LB = (dist(s, [lo_c, hi_c]) / dsig)^2 + (dist(|anc|, [wlo_c, whi_c]) / 2)^2,
exceeds the cost U of a seed candidate (with a 1e-6 margin).  This test transcribes the kernel's rule in numpy -- the
per-chunk ranges, the seed (row where the slab crosses sigma0 on a strided subset of the phi nodes, cheapest by an FP32
estimate, evaluated with the reference's operations) and the comparison -- and checks against the brute-force cost of
windspeed.py:220-225 (oracle restatement) on seeded adversarial pixels that
  * the chunk of the reference's argmin, and of every candidate tying with it, is never skipped,
  * on the benchmark recipe most of the slab is skipped (the rule is worth having).
No GPU and nothing of the product is involved."""
import numpy as np
import pytest

import oracle

CHUNK = 16
SEED_MAX = 64


@pytest.fixture(scope="module")
def slabs():
    gi = np.array([20.0, 33.0, 45.0])
    gw, gp = np.linspace(0.2, 50, 499), np.linspace(0, 180, 181)
    lut_db = 10 * np.log10(oracle.lut_build("gmf_cmod5n", gi, gw, gp) + 1e-15)
    return gi, gw, gp, lut_db


def exact_cost(slab, gw, cphi, sphi, a, b, s, dsig):
    """windspeed.py:220-225 with the reference's operation order, whole slab."""
    jw = ((gw[:, None] * cphi[None, :] - a) / 2) ** 2 + ((gw[:, None] * sphi[None, :] - b) / 2) ** 2
    return jw + ((slab - s) / dsig) ** 2


def chunk_ranges(slab, gw, phis=slice(None)):
    """Value range of every chunk over the phi nodes `phis` (all of them, or one phi group = 64 consecutive nodes: what
    one float2 slot of the scan lanes covers) and the |wspd| range of every chunk."""
    n_chunks = (len(gw) + CHUNK - 1) // CHUNK
    lo, hi, wlo, whi = (np.empty(n_chunks) for _ in range(4))
    for c in range(n_chunks):
        rows = slice(c * CHUNK, min((c + 1) * CHUNK, len(gw)))
        v = slab[rows][:, phis]
        fin = np.isfinite(v).all()
        lo[c], hi[c] = (v.min(), v.max()) if fin else (-np.inf, np.inf)
        wlo[c], whi[c] = np.abs(gw[rows]).min(), np.abs(gw[rows]).max()
    return lo, hi, wlo, whi


def phi_groups(n_phi):
    return [slice(64 * g, min(64 * (g + 1), n_phi)) for g in range((n_phi + 63) // 64)]


def kept_chunks(slab, gw, cphi, sphi, ranges, a, b, s, s_mid, dsig, s_rng=None):
    """Transcription of k_tile_plan for one pixel of a tile whose median sigma0 is s_mid.  s_rng = (s_lo, s_hi): the kernel
    evaluates the sigma0 part of the bound once per tile, for an interval that encloses the sigma0 of all its pixels (a
    weaker bound than the pixel's own, which is what s_rng=None gives)."""
    lo, hi, wlo, whi = ranges
    n_w, n_p = slab.shape
    stride = (n_p + SEED_MAX - 1) // SEED_MAX
    best, bflat = np.inf, None
    inv_d = 1.0 / abs(dsig)
    for ip in range(0, n_p, stride):
        col = slab[:, ip]
        l, h = 0, n_w
        while l < h:  # bisection as in the kernel (any row is a valid seed)
            mid = (l + h) >> 1
            if col[mid] < s_mid:
                l = mid + 1
            else:
                h = mid
        r = min(l, n_w - 1)
        if 0 < l < n_w and abs(col[l - 1] - s_mid) <= abs(col[l] - s_mid):
            r = l - 1
        if l == n_w:   # sigma0 above the column's last value: the column's largest value may sit at a lower wind speed
            rm = int(np.argmax(col))
            if col[rm] > col[r]:
                r = rm
                if col[rm] >= s_mid:
                    a_, b_ = 0, rm
                    while a_ < b_:
                        mid = (a_ + b_) >> 1
                        if col[mid] < s_mid:
                            a_ = mid + 1
                        else:
                            b_ = mid
                    r = a_
        f = np.float32
        ta = f(0.5) * (f(gw[r] * cphi[ip]) - f(a))
        tz = f(0.5) * (f(gw[r] * sphi[ip]) - f(b))
        ts = f(col[r] * inv_d) - f(s * inv_d)
        J = ta * ta + tz * tz + ts * ts
        if J < best:
            best, bflat = J, (r, ip)
    r, ip = bflat
    U = ((gw[r] * cphi[ip] - a) / 2) ** 2 + ((gw[r] * sphi[ip] - b) / 2) ** 2 + ((slab[r, ip] - s) / dsig) ** 2
    A = np.hypot(a, b)
    thr = (U * (1 + 1e-6) + 1e-6 * (1 + A * A + np.abs(gw).max() ** 2)) * (1 + 2e-9)
    s_lo, s_hi = (s, s) if s_rng is None else s_rng
    ds = np.maximum(np.maximum(lo - s_hi, s_lo - hi), 0.0) * inv_d
    dw = np.maximum(np.maximum(wlo - A, A - whi), 0.0) * 0.5
    lb = ds * ds + dw * dw
    return ~(lb > thr)


def pixels(rng, slab, gw, gp, n, kind):
    iw, ip = rng.integers(0, len(gw), n), rng.integers(0, len(gp), n)
    s = slab[iw, ip].copy()
    w, phi = gw[iw].copy(), np.radians(gp[ip])
    if kind == "benchmark":      # bench.py recipe: 5 % noise, ancillary within 2 m/s / 20 deg
        s += rng.normal(0, 0.2, n)
        wa, pa = w + rng.normal(0, 2, n), phi + np.radians(rng.normal(0, 20, n))
    elif kind == "nodes":        # sigma0 exactly on LUT nodes and ancillary exactly on candidates: ties everywhere
        wa, pa = w, phi
        s[::3] = slab[np.minimum(iw[::3] + 1, len(gw) - 1), ip[::3]]
    elif kind == "hostile":      # +-15 dB off, ancillary far away
        s += rng.choice([-15.0, 15.0], n)
        wa, pa = w + 10, phi + np.pi / 2
    else:                         # wide: everything random, huge and zero ancillary winds
        s = rng.uniform(-60, 20, n)
        wa, pa = rng.choice([0.0, 1e-3, 5, 30, 200], n), rng.uniform(0, 2 * np.pi, n)
    a, b = wa * np.cos(pa), np.abs(wa * np.sin(pa))
    return s, a, b


@pytest.mark.parametrize("kind", ["benchmark", "nodes", "hostile", "wide"])
def test_argmin_chunk_is_never_pruned(slabs, kind):
    gi, gw, gp, lut_db = slabs
    cphi, sphi = np.cos(np.radians(gp)), np.sin(np.radians(gp))
    rng = np.random.default_rng({"benchmark": 1, "nodes": 2, "hostile": 3, "wide": 4}[kind])
    dsig = 0.1
    kept_frac = []
    groups = phi_groups(len(gp))
    cells_kept = []
    for b_ in range(len(gi)):
        slab = lut_db[b_]
        ranges = chunk_ranges(slab, gw)
        granges = [chunk_ranges(slab, gw, g) for g in groups]
        s, a, b = pixels(rng, slab, gw, gp, 60, kind)
        order = np.argsort(s)
        for k0 in range(0, len(order), 32):       # tiles of 32 pixels in sigma0 order share the seed rows
            tile = order[k0:k0 + 32]
            s_mid = s[tile[len(tile) // 2]]
            f_lo, f_hi = np.float32(s[tile].min()), np.float32(s[tile].max())   # FP32 images widened by one ulp, as in the kernel
            s_rng = (float(np.nextafter(f_lo, np.float32(-np.inf))), float(np.nextafter(f_hi, np.float32(np.inf))))
            assert s_rng[0] <= s[tile].min() and s[tile].max() <= s_rng[1]
            for q in tile:
                keep = kept_chunks(slab, gw, cphi, sphi, ranges, a[q], b[q], s[q], s_mid, dsig, s_rng)
                J = exact_cost(slab, gw, cphi, sphi, a[q], b[q], s[q], dsig)
                ties = np.argwhere(J == J.min())
                assert keep[ties[:, 0] // CHUNK].all(), (kind, b_, q)
                # stronger: every skipped chunk is strictly worse than the minimum
                worst_skipped = min((J[c * CHUNK:(c + 1) * CHUNK].min() for c in np.flatnonzero(~keep)), default=np.inf)
                assert worst_skipped > J.min()
                # the same per cell = (chunk, phi group), the granularity the scan warps skip at
                n_cells = 0
                for g, gr in zip(groups, granges):
                    keep_g = kept_chunks(slab, gw, cphi, sphi, gr, a[q], b[q], s[q], s_mid, dsig, s_rng)
                    assert not (keep_g & ~keep).any()          # a cell's bound is at least its chunk's
                    assert keep_g[ties[:, 0][(ties[:, 1] >= g.start) & (ties[:, 1] < g.stop)] // CHUNK].all(), (kind, b_, q)
                    worst = min((J[c * CHUNK:(c + 1) * CHUNK, g].min() for c in np.flatnonzero(~keep_g)), default=np.inf)
                    assert worst > J.min()
                    n_cells += kept_chunks(slab, gw, cphi, sphi, gr, a[q], b[q], s[q], s[q], dsig).sum()   # dense tile, as below
                cells_kept.append(n_cells / len(groups))
                # a dense tile (the full scene: >= 1e5 pixels per bin) has its median sigma0 next to every pixel's own
                kept_frac.append(kept_chunks(slab, gw, cphi, sphi, ranges, a[q], b[q], s[q], s[q], dsig).mean())
    if kind == "benchmark":
        assert np.mean(kept_frac) < 0.2, np.mean(kept_frac)
        # the phi groups prune further inside the kept chunks (in chunk equivalents)
        assert np.mean(cells_kept) < 0.8 * np.mean(kept_frac) * len(ranges[0]), (np.mean(cells_kept), np.mean(kept_frac))


def test_non_finite_chunks_give_no_sigma0_bound(slabs):
    gi, gw, gp, lut_db = slabs
    slab = lut_db[1].copy()
    slab[100, 7] = np.inf
    slab[300:310, :] = -np.inf
    lo, hi, _, _ = chunk_ranges(slab, gw)
    assert lo[100 // CHUNK] == -np.inf and hi[100 // CHUNK] == np.inf
    assert lo[300 // CHUNK] == -np.inf and hi[304 // CHUNK] == np.inf
    cphi, sphi = np.cos(np.radians(gp)), np.sin(np.radians(gp))
    keep = kept_chunks(slab, gw, cphi, sphi, chunk_ranges(slab, gw), 3.0, 4.0, slab[98, 7], slab[98, 7], 0.1)
    J = exact_cost(slab, gw, cphi, sphi, 3.0, 4.0, slab[98, 7], 0.1)
    iw = np.unravel_index(np.argmin(J), J.shape)[0]
    assert keep[iw // CHUNK]


def _next_chunk(mask, c, sh, n_chunks):
    """Transcription of next_chunk (xs_scan.cu): smallest chunk > c whose mask bit (chunk >> sh) is set, or n_chunks."""
    c1 = c + 1
    if c1 >= n_chunks:
        return n_chunks
    b = c1 >> sh
    if (mask >> b) & 1:
        return c1
    rest = 0 if b >= 31 else ((mask >> (b + 1)) << (b + 1)) & 0xFFFFFFFF
    if not rest:
        return n_chunks
    c2 = ((rest & -rest).bit_length() - 1) << sh
    return c2 if c2 < n_chunks else n_chunks


def test_chunk_iteration_over_a_mask():
    """The scan walks a tile's plan with next_chunk: it must visit exactly the chunks whose mask bit is set, in ascending
    order, for every granularity of the mask (1, 2, 4, 8 chunks per bit) and slabs whose last bit is only partly filled."""
    rng = np.random.default_rng(7)
    for sh in range(4):
        for n_chunks in (4, 5, 31, 32, 33, 69, 100, 250):
            nb = (n_chunks + (1 << sh) - 1) >> sh
            if nb > 32:
                continue
            for _ in range(40):
                mask = int(rng.integers(0, 1 << nb)) | (1 << int(rng.integers(0, nb)))
                want = [c for c in range(n_chunks) if (mask >> (c >> sh)) & 1]
                got, c = [], _next_chunk(mask, -1, sh, n_chunks)
                while c < n_chunks:
                    got.append(c)
                    c = _next_chunk(mask, c, sh, n_chunks)
                assert got == want, (sh, n_chunks, bin(mask))
