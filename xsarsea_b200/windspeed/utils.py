"""Pre-processors feeding `dsig_cr` (reference xsarsea/windspeed/utils.py): `get_dsig`, `get_dsig_wspd`,
`nesz_flattening`.  Cheap element-wise / per-row numpy work on the host side of the boundary (SURVEY.md section 8
row F1, "next"); they accept numpy or labelled arrays and keep the container."""
import logging
import warnings

import numpy as np

logger = logging.getLogger("xsarsea.windspeed.utils")
logger.setLevel(logging.INFO)

# (b, c0, gamma, k) of the sigmoid blending weight per name (utils.py:27-43)
_DSIG_WSPD = {
    "dsig_wspd_rs2_v3": (-0.4908643753212401, 16.763199934792965, 1.3891445172991084, 20.616914824394343),
    "dsig_wspd_s1_ew_rec_v3": (-0.5858970325653666, 16.50039320910609, 1.1032031322520397, 7.434663633997121),
    "dsig_wspd_rcm_v3": (-0.7920301376936547, 15.8288289109038, 0.24040294696606557, 0.2538177092195224),
}


def get_dsig_wspd(name, U_crosspol, SNR_cr):
    """Co/cross blending weight in [0, 1] (utils.py:18-44): sigmoid in cross-pol wind speed whose centre moves with
    the cross-pol SNR, times a drop-off above 30 m/s."""
    b, c0_base, gamma, k = _DSIG_WSPD[name]
    u_max = 30
    core = 1 / (1 + np.exp(-b * (U_crosspol - (c0_base - gamma * SNR_cr))))
    drop = 1 / (1 + np.exp((U_crosspol - u_max) * k))
    return np.clip(core * drop, 0, 1)


def get_dsig(name, inc, sigma0_cr, nesz_cr):
    """dsig_cr value(s) by model name (utils.py:47-91)."""
    snr = sigma0_cr / nesz_cr
    if name == "gmf_s1_v2":
        c0, c1, d0, d1 = 1.57952257, 25.61843791, 1.46852088, 1.4058646
        expo = d0 + d1 / (1 + np.exp(-c0 * (inc - c1)))
        return 1 / np.sqrt(1 * snr ** expo)
    if name == "gmf_rs2_v2":
        return 1 / np.sqrt(1 * snr ** 8)
    if name in ("sarwing_lut_cmodms1ahw", "nc_lut_cmodms1ahw"):
        return (1.25 / snr) ** 4.0
    raise ValueError(
        "dsig names different than 'gmf_s1_v2' or 'gmf_rs2_v2' or 'sarwing_lut_cmodms1ahw' or 'nc_lut_cmodms1ahw' "
        "are not handled. You can compute your own dsig_cr.")


def nesz_flattening(noise, inc):
    """Flatten the noise (nesz, linear, shape (line, sample)) by an order-1 polynomial fit in dB along each line
    (utils.py:94-163): NaNs are first filled with the column mean; the fit uses the column-mean incidence;
    the flattened value is 10**((a*inc + b - 1)/10)."""
    if noise.ndim != 2:
        raise IndexError("Only 2D noise allowed")
    noise_v = np.asarray(getattr(noise, "values", noise), dtype=np.float64)
    inc_v = np.asarray(getattr(inc, "values", inc), dtype=np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        col_mean = np.nanmean(noise_v, axis=0)
        inc_row = np.nanmean(inc_v, axis=0)
    out = np.empty_like(noise_v)
    for r in range(noise_v.shape[0]):
        row = np.where(np.isnan(noise_v[r]), col_mean, noise_v[r])
        with np.errstate(all="ignore"):
            row_db = 10.0 * np.log10(row)
        ok = np.isfinite(row_db)
        if not ok.any():
            out[r] = np.nan
            continue
        a, b = np.polyfit(inc_row[ok], row_db[ok], 1)
        out[r] = 10.0 ** ((inc_row * a + b - 1.0) / 10.0)
    return out
