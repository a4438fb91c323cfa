"""Generate golden vectors from the REAL reference arithmetic (dev-time; needs /root/reference + numba).

Run:  python tests/golden/make_golden.py
Writes tests/golden/gmf_points.npz, inv_small.npz, inv_slabs.npz, inv_ifr2.npz, dsig_utils.npz, gradients.npz.

Everything numerical in these files comes out of the reference's own code, executed through
tests/golden/_refload.py: the numba-compiled scalar GMFs (gmfs.py:206-230 over gmfs_impl.py) and the
numba-compiled inversion kernel (windspeed.py:183-282 wrapped as at :306-323).  The dB conversion uses the
reference expression 10*np.log10(lut + 1e-15) (models.py:215).  The test suite never imports the reference:
it only reads the .npz files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refload  # noqa: E402

ws = _refload.load_reference()
from xsarsea.windspeed.models import Model  # noqa: E402

MODELS = ["gmf_cmod5", "gmf_cmod5n", "gmf_cmod5n_pr_zhangA", "gmf_cmod5n_pr_mouche1", "gmf_cmodifr2", "gmf_rs2_v2",
          "gmf_s1_v2", "gmf_rcm_noaa", "gmf_s1_v3_ew_rec", "gmf_rs2_v3", "gmf_rcm_v3", "gmf_rcm_v4", "gmf_rs2_v4"]


def linspace_grid(r, step):
    return np.linspace(r[0], r[1], num=int(np.round((r[1] - r[0]) / step) + 1))


def ref_lut(name, inc, wspd, phi):
    """Reference K2 (guvectorize outer product); phi=None for cross-pol (dummy NaN axis as gmfs.py:287-290)."""
    m = Model._available_models[name]
    f = m._gmf_function("numba_guvectorize")
    if phi is None:
        return f(inc, wspd, np.array([np.nan]))[..., 0]
    return f(inc, wspd, phi)


def gmf_points():
    rng = np.random.default_rng(1234)
    out = {}
    n = 600
    for name in MODELS:
        m = Model._available_models[name]
        f = m._gmf_function("numba_njit")
        inc = rng.uniform(16, 66, n)
        wspd = rng.uniform(m.wspd_range[0], m.wspd_range[1], n)
        phi = rng.uniform(0, 360, n)
        # exact grid nodes and branch neighbourhoods too
        inc[:8] = [16, 17, 25, 35, 40, 50, 65, 66]
        wspd[:8] = [m.wspd_range[0], m.wspd_range[1], 5, 10, 15, 20, 25, 30]
        phi[:8] = [0, 45, 90, 135, 180, 225, 270, 360]
        val = np.array([f(a, b, c) for a, b, c in zip(inc, wspd, phi)])
        out[name + "/inc"] = inc
        out[name + "/wspd"] = wspd
        out[name + "/phi"] = phi
        out[name + "/sigma0"] = val
        # vectorised (K3) flavour must agree with the scalar one
        fv = m._gmf_function("numba_vectorize")
        assert np.array_equal(fv(inc, wspd, phi), val, equal_nan=True), name
        # small outer-product LUT (K2)
        gi = np.linspace(17, 65, 7)
        gw = np.linspace(m.wspd_range[0], m.wspd_range[1], 9)
        gp = np.linspace(0, 180, 5) if m.phi_range is not None else None
        out[name + "/lut_inc"] = gi
        out[name + "/lut_wspd"] = gw
        if gp is not None:
            out[name + "/lut_phi"] = gp
        out[name + "/lut"] = ref_lut(name, gi, gw, gp)
    np.savez_compressed(os.path.join(HERE, "gmf_points.npz"), **out)
    print("gmf_points.npz", len(out))


def synth_pixels(rng, n, inc_lo, inc_hi, co_name, cr_name, wmax=25.0):
    """SURVEY.md D2 recipe (1-D)."""
    inc = rng.uniform(inc_lo, inc_hi, n)
    wspd = rng.uniform(2, wmax, n)
    phi = rng.uniform(0, 360, n)
    fco = Model._available_models[co_name]._gmf_function("numba_vectorize")
    fcr = Model._available_models[cr_name]._gmf_function("numba_vectorize")
    s_co = fco(inc, wspd, phi) * np.exp(rng.normal(0, 0.05, n))
    s_cr = fcr(inc, wspd, np.full(n, np.nan)) * np.exp(rng.normal(0, 0.05, n))
    anc = (wspd + rng.normal(0, 2, n)) * np.exp(1j * np.deg2rad(phi + rng.normal(0, 20, n)))
    nesz = 10 ** (-3.2)
    dsig = (1.25 / (s_cr / nesz)) ** 4.0
    return inc, s_co, s_cr, dsig, anc


def add_edge_cases(inc, s_co, s_cr, dsig, anc, inc_vals):
    """Edge rows of SURVEY.md D2: NaNs, negative / zero sigma0, out-of-range incidence, Im(anc)=0."""
    k = 0

    def put(**kw):
        nonlocal k
        for name, v in kw.items():
            {"inc": inc, "s_co": s_co, "s_cr": s_cr, "dsig": dsig, "anc": anc}[name][k] = v
        k += 1

    put(inc=np.nan)
    put(s_co=np.nan)
    put(s_cr=np.nan)
    put(s_co=np.nan, s_cr=np.nan)
    put(dsig=np.nan)
    put(anc=np.nan)
    put(anc=complex(np.nan, 1.0))
    put(anc=complex(3.0, np.nan))
    put(s_co=-1e-3)
    put(s_cr=-1e-3)
    put(s_co=0.0)
    put(s_cr=0.0)
    put(s_co=-1e-16)
    put(anc=complex(7.0, 0.0))
    put(anc=complex(-7.0, 0.0))
    put(anc=complex(7.0, -0.0))
    put(anc=complex(0.0, 0.0))
    put(anc=complex(0.0, -5.0))
    put(anc=complex(1e-3, -1e-3))
    put(s_co=np.nan, anc=np.nan)           # cross-only pixel with no ancillary
    put(s_co=np.nan, anc=complex(5.0, 5.0))  # cross-only pixel with ancillary
    put(s_co=1e3)
    put(s_cr=1e3)
    put(s_co=np.inf)
    put(dsig=0.0)
    put(dsig=np.inf)
    put(dsig=-0.3)
    for v in inc_vals:
        put(inc=v)
    return k


def run_ref_kernel(co_db, gi, gw, gp, cr_db, gic, gwc, inc, s_co, s_cr, dsig, anc, dsig_co=0.1):
    closure = dict(
        np_sigma0_co_lut_db=np.ascontiguousarray(np.transpose(co_db, (1, 2, 0))) if co_db is not None else np.array([[[]]]),
        np_wspd_dim=gw if co_db is not None else np.array([]),
        np_phi_dim=gp if co_db is not None else np.array([]),
        np_inc_dim=gi if co_db is not None else np.array([]),
        phi_180=bool((180 - (gp[-1] - gp[0])) < 2) if co_db is not None else False,
        np_sigma0_cr_lut_db=np.ascontiguousarray(np.transpose(cr_db, (1, 0))) if cr_db is not None else np.array([[]]),
        np_wspd_lut_cr=gwc if cr_db is not None else np.array([]),
        np_inc_cr_dim=gic if cr_db is not None else np.array([]),
        dsig_co=dsig_co,
    )
    vect, _ = _refload.reference_inversion_kernel(closure)
    with np.errstate(all="ignore"):
        s_co_db = 10 * np.log10(s_co + 1e-15)   # windspeed.py:126
        s_cr_db = 10 * np.log10(s_cr + 1e-15)   # windspeed.py:128
        # 2-D so that the parallel gufunc has a loop dimension
        shp = (-1, 50)
        o_co, o_cr = vect(inc.reshape(shp), s_co_db.reshape(shp), s_cr_db.reshape(shp), dsig.reshape(shp),
                          anc.reshape(shp))
    return s_co_db, s_cr_db, o_co.reshape(-1), o_cr.reshape(-1)


def inv_small():
    """Coarse LUT (51 x 101 x 37 co, 51 x 155 cross) over the full incidence range, 3000 px + edge rows."""
    rng = np.random.default_rng(0)
    gi = linspace_grid([16.0, 66.0], 1.0)
    gw = linspace_grid([0.2, 50.0], 0.5)
    gp = linspace_grid([0.0, 180.0], 5.0)
    gwc = linspace_grid([3.0, 80.0], 0.5)
    co_db = 10 * np.log10(ref_lut("gmf_cmod5n", gi, gw, gp) + 1e-15)
    cr_db = 10 * np.log10(ref_lut("gmf_s1_v2", gi, gwc, None) + 1e-15)
    n = 3000
    inc, s_co, s_cr, dsig, anc = synth_pixels(rng, n, 17.5, 49.5, "gmf_cmod5n", "gmf_s1_v2")
    nan_mask = rng.uniform(size=n) < 0.01
    s_co[nan_mask] = np.nan
    add_edge_cases(inc, s_co, s_cr, dsig, anc, [10.0, 16.05, 16.5, 17.5, 66.0, 70.0, 16.499999999999996])
    dsig[n // 2:] = 0.1  # second half: scalar dsig_cr variant
    s_co_db, s_cr_db, o_co, o_cr = run_ref_kernel(co_db, gi, gw, gp, cr_db, gi, gwc, inc, s_co, s_cr, dsig, anc)
    # mono-pol variants through the same kernel: absent polarisation = all-NaN raster (windspeed.py:71,170)
    nanr = np.full(n, np.nan)
    _, _, co_only, co_only_cr = run_ref_kernel(co_db, gi, gw, gp, None, None, None, inc, s_co, nanr, nanr * 0 + 0.1, anc)
    _, _, cr_only_co, cr_only = run_ref_kernel(None, None, None, None, cr_db, gi, gwc, inc, nanr, s_cr, dsig, nanr + 0j)
    np.savez_compressed(os.path.join(HERE, "inv_small.npz"), inc_grid=gi, wspd_grid=gw, phi_grid=gp, wspd_cr_grid=gwc,
                        co_lut_db=co_db, cr_lut_db=cr_db, inc=inc, s0_co=s_co, s0_cr=s_cr, dsig_cr=dsig, anc=anc,
                        s0_co_db=s_co_db, s0_cr_db=s_cr_db, out_co=o_co, out_cr=o_cr, co_only=co_only,
                        co_only_cr=co_only_cr, cr_only_co=cr_only_co, cr_only=cr_only)
    print("inv_small.npz")


def inv_slabs():
    """Full-resolution wspd x phi (499 x 181) at three incidence nodes, cross-pol 771; 1500 px."""
    rng = np.random.default_rng(7)
    gi = np.array([30.0, 30.1, 45.3])
    gw = linspace_grid([0.2, 50.0], 0.1)
    gp = linspace_grid([0.0, 180.0], 1.0)
    gwc = linspace_grid([3.0, 80.0], 0.1)
    co_db = 10 * np.log10(ref_lut("gmf_cmod5n", gi, gw, gp) + 1e-15)
    cr_db = 10 * np.log10(ref_lut("gmf_s1_v2", gi, gwc, None) + 1e-15)
    n = 1500
    inc, s_co, s_cr, dsig, anc = synth_pixels(rng, n, 29.9, 30.2, "gmf_cmod5n", "gmf_s1_v2")
    inc[n // 2:] = rng.uniform(40, 50, n - n // 2)
    add_edge_cases(inc, s_co, s_cr, dsig, anc, [30.05, 30.049999999999997, 30.050000000000004, 37.7, 37.699999999999996])
    s_co_db, s_cr_db, o_co, o_cr = run_ref_kernel(co_db, gi, gw, gp, cr_db, gi, gwc, inc, s_co, s_cr, dsig, anc)
    np.savez_compressed(os.path.join(HERE, "inv_slabs.npz"), inc_grid=gi, wspd_grid=gw, phi_grid=gp, wspd_cr_grid=gwc,
                        co_lut_db=co_db.astype(np.float64), cr_lut_db=cr_db, inc=inc, s0_co=s_co, s0_cr=s_cr,
                        dsig_cr=dsig, anc=anc, s0_co_db=s_co_db, s0_cr_db=s_cr_db, out_co=o_co, out_cr=o_cr)
    print("inv_slabs.npz")


def inv_ifr2():
    """cmodifr2 coarse LUT: negative linear values -> NaN in dB -> 'first NaN wins' argmin;
    plus a phi grid spanning < 178 deg (phi_180 False, windspeed.py:152-156)."""
    rng = np.random.default_rng(11)
    gi = linspace_grid([16.0, 66.0], 2.0)
    gw = linspace_grid([0.2, 50.0], 0.6)
    gp = linspace_grid([0.0, 180.0], 7.5)
    with np.errstate(all="ignore"):
        co_db = 10 * np.log10(ref_lut("gmf_cmodifr2", gi, gw, gp) + 1e-15)
    n = 600
    inc, s_co, s_cr, dsig, anc = synth_pixels(rng, n, 16, 66, "gmf_cmod5n", "gmf_rs2_v3")
    nanr = np.full(n, np.nan)
    s_co_db, _, o_co, o_cr = run_ref_kernel(co_db, gi, gw, gp, None, None, None, inc, s_co, nanr, nanr * 0 + 0.1, anc)
    # phi grid [0, 170]: not mirrored; same pixels
    gp2 = linspace_grid([0.0, 170.0], 10.0)
    co2_db = 10 * np.log10(ref_lut("gmf_cmod5n", gi, gw, gp2) + 1e-15)
    _, _, o2_co, _ = run_ref_kernel(co2_db, gi, gw, gp2, None, None, None, inc, s_co, nanr, nanr * 0 + 0.1, anc, dsig_co=0.25)
    np.savez_compressed(os.path.join(HERE, "inv_ifr2.npz"), inc_grid=gi, wspd_grid=gw, phi_grid=gp, co_lut_db=co_db,
                        inc=inc, s0_co=s_co, anc=anc, s0_co_db=s_co_db, out_co=o_co, out_cr=o_cr, phi_grid2=gp2,
                        co2_lut_db=co2_db, out2_co=o2_co, dsig_co2=0.25)
    print("inv_ifr2.npz")


def dsig_utils():
    """Outputs of the reference's own windspeed/utils.py: get_dsig (:47-91), get_dsig_wspd (:18-44),
    nesz_flattening (:94-163) on seeded inputs with the edge cases the formulas have (negative / zero / NaN sigma0,
    NaN holes and an all-NaN line in the noise, a non-positive noise sample, constant incidence)."""
    import warnings

    from xsarsea.windspeed import utils as ref

    rng = np.random.default_rng(7)
    n = 4000
    inc = rng.uniform(17.0, 50.0, n)
    nesz = 10 ** rng.uniform(-3.6, -2.6, n)
    s_cr = nesz * 10 ** rng.uniform(-0.5, 2.0, n)
    s_cr[:8] = [0.0, -1e-4, np.nan, 1e-12, np.inf, 1e3, 1e-3, 1e-3]
    nesz[5:8] = [1e-3, np.nan, 0.0]
    out = dict(inc=inc, sigma0_cr=s_cr, nesz_cr=nesz)
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for name in ("gmf_s1_v2", "gmf_rs2_v2", "nc_lut_cmodms1ahw", "sarwing_lut_cmodms1ahw"):
            out["dsig__" + name] = ref.get_dsig(name, inc, s_cr, nesz)
        u = rng.uniform(0.0, 80.0, n)
        snr = rng.uniform(-5.0, 25.0, n)
        u[:4] = [np.nan, 30.0, 0.0, 1e3]
        snr[4:6] = [np.nan, 1e3]
        out.update(u_crosspol=u, snr_cr=snr)
        for name in ("dsig_wspd_rs2_v3", "dsig_wspd_s1_ew_rec_v3", "dsig_wspd_rcm_v3"):
            out["wspd__" + name] = ref.get_dsig_wspd(name, u, snr)
        # nesz_flattening: an EW-like noise field (scalloped in range, slowly varying in azimuth)
        h, w = 96, 700
        incg = np.broadcast_to(np.linspace(19.0, 47.0, w), (h, w)) + rng.normal(0, 1e-3, (h, w))
        noise = 10 ** ((-28.0 - 0.12 * (incg - 19.0) + 0.6 * np.sin(incg * 1.7) + rng.normal(0, 0.15, (h, w))) / 10.0)
        noise[rng.random((h, w)) < 0.02] = np.nan
        noise[5, :] = np.nan                 # a whole line missing: filled from the column means
        noise[:, 11] = np.nan                # a whole column missing: stays NaN -> excluded from every fit
        noise[7, 20] = 0.0                   # log10(0) = -inf: excluded
        noise[8, 21] = -1e-3                 # log10(<0) = nan: excluded
        incg[9, 30] = np.nan                 # NaN incidence only enters through the column nanmean
        out.update(noise=noise, inc2d=incg, noise_flat=ref.nesz_flattening(noise, incg))
        # degenerate: every column missing -> np.polyfit TypeError -> NaN lines
        nn = np.full((3, 16), np.nan)
        out.update(noise_allnan_flat=ref.nesz_flattening(nn, incg[:3, :16].copy()))
    np.savez_compressed(os.path.join(HERE, "dsig_utils.npz"), **out)
    print("dsig_utils.npz")


def gradients():
    """local_gradients of gradients.py:588-634 evaluated with the libraries the reference delegates to (cv2.Scharr,
    scipy.signal.convolve2d) through oracle/gradients.py -- the reference function itself needs xarray (not installable
    here), so this file pins the third-party arithmetic, not an output of the reference function."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import gradients as og

    rng = np.random.default_rng(11)
    out = {}
    for tag, (h, w) in (("odd", (101, 203)), ("even", (64, 96)), ("tiny", (2, 2)), ("thin", (3, 40))):
        yy, xx = np.mgrid[0:h, 0:w]
        img = 0.05 + 0.02 * np.sin(0.21 * xx + 0.13 * yy) + 0.01 * rng.standard_normal((h, w))   # streaky sigma0
        if tag == "odd":
            img[40:44, 50:60] = np.nan       # a masked patch: NaN spreads through the stencils, partial 2x2 blocks
            img[0, 0] = np.nan               # corner: exercises both border rules
        g2, g3, c = og.local_gradients(img)
        out.update({f"{tag}/image": img, f"{tag}/G2": g2, f"{tag}/G3": g3, f"{tag}/c": c})
    np.savez_compressed(os.path.join(HERE, "gradients.npz"), **out)
    print("gradients.npz")


if __name__ == "__main__":
    if sys.argv[1:] == ["dsig_utils"]:
        dsig_utils()
        sys.exit(0)
    if sys.argv[1:] == ["gradients"]:
        gradients()
        sys.exit(0)
    gmf_points()
    inv_small()
    inv_slabs()
    inv_ifr2()
    dsig_utils()
    gradients()
