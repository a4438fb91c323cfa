"""CPU ORACLE (test infrastructure) for the dsig_cr pre-processors -- SURVEY.md section 8 row F1.

numpy restatement of the reference's xsarsea/windspeed/utils.py: `get_dsig` (:47-91), `get_dsig_wspd` (:18-44) and
`nesz_flattening` (:94-163).  Pinned by tests/golden/dsig_utils.npz, which holds outputs of the reference's own
functions (tests/golden/make_golden.py::dsig_utils, run where /root/reference is mounted): bit-exact for the
element-wise formulas and for the flattening (same numpy calls in the same order).  Only tests/, smoke() and the
CPU-baseline legs of the benches may import this module.
"""
import warnings

import numpy as np

# (b, c0, gamma, k) of the blending sigmoid per name (utils.py:26-42)
DSIG_WSPD_COEF = {
    "dsig_wspd_rs2_v3": (-0.4908643753212401, 16.763199934792965, 1.3891445172991084, 20.616914824394343),
    "dsig_wspd_s1_ew_rec_v3": (-0.5858970325653666, 16.50039320910609, 1.1032031322520397, 7.434663633997121),
    "dsig_wspd_rcm_v3": (-0.7920301376936547, 15.8288289109038, 0.24040294696606557, 0.2538177092195224),
}


def get_dsig_wspd(name, U_crosspol, SNR_cr):
    """utils.py:18-44: sigmoid in cross-pol wind speed whose centre moves with the SNR, times a drop above 30 m/s."""
    b, c0_base, gamma, k = DSIG_WSPD_COEF[name]
    core = 1 / (1 + np.exp(-b * (U_crosspol - (c0_base - gamma * SNR_cr))))
    drop = 1 / (1 + np.exp((U_crosspol - 30) * k))
    return np.clip(core * drop, 0, 1)


def get_dsig(name, inc, sigma0_cr, nesz_cr):
    """utils.py:47-91."""
    if name == "gmf_s1_v2":
        c0, c1, d0, d1 = np.array([1.57952257, 25.61843791, 1.46852088, 1.4058646])
        expo = d0 + d1 / (1 + np.exp(-c0 * (inc - c1)))
        return 1 / np.sqrt(1 * (sigma0_cr / nesz_cr) ** expo)
    if name == "gmf_rs2_v2":
        return 1 / np.sqrt(1 * (sigma0_cr / nesz_cr) ** 8)
    if name in ("sarwing_lut_cmodms1ahw", "nc_lut_cmodms1ahw"):
        return (1.25 / (sigma0_cr / nesz_cr)) ** 4.0
    raise ValueError(name)


def nesz_flattening(noise, inc):
    """utils.py:94-163: per-line order-1 np.polyfit of the NaN-filled noise in dB against the column-mean incidence."""
    noise, inc = np.asarray(noise), np.asarray(inc)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        col_mean = np.nanmean(noise, axis=0)
        inc_row = np.nanmean(inc, axis=0)
    out = np.empty(noise.shape, dtype=np.float64)
    for r in range(noise.shape[0]):
        row = noise[r].copy()
        hole = np.isnan(row)
        row[hole] = col_mean[hole]
        with np.errstate(all="ignore"):
            row_db = 10.0 * np.log10(row)
        ok = np.isfinite(row_db)
        try:
            coef = np.polyfit(inc_row[ok], row_db[ok], 1)
        except TypeError:  # no finite point
            out[r] = np.nan
            continue
        out[r] = 10.0 ** ((inc_row * coef[0] + coef[1] - 1.0) / 10.0)
    return out
