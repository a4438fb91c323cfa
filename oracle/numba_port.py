"""ORACLE (test infrastructure): the reference's CPU program shape, used as the timed CPU baseline.

`make_inverter` builds a numba gufunc with the decorator arguments the reference uses
(windspeed.py:306-323: signature "(n),(n),(n),(n),(n)->(n),(n)", f64 x4 + c128 -> c128 x2,
fastmath={"nnan": False}, target="parallel") around a per-pixel loop that, like windspeed.py:190-281,
evaluates the whole wspd x phi cost surface with array expressions on a LUT stored [wspd][phi][inc]
(:145-147) and takes np.argmin.  It is therefore the same amount and kind of CPU work as the reference
kernel; tests check it bit-for-bit against oracle/c (and, through tests/golden, against the reference).
"""
from __future__ import annotations

import numpy as np


# The per-pixel program.  Its free names (co_t, u_tab, ...) are resolved as *globals* of a function object created per
# inverter (types.FunctionType over this code with a private globals dict), which is how numba sees the reference's
# kernel too once its closure is materialised; numba freezes global arrays as constants.
def _per_line(theta, zco, zcr, dcr, prior, res_co, res_x):
    for k in range(len(theta)):
        th = theta[k]
        if np.isnan(th):
            res_co[k] = np.nan
            res_x[k] = np.nan
            continue
        if not np.isnan(np.abs(zco[k])) and np.isnan(np.abs(prior[k])):
            res_co[k] = np.nan
            res_x[k] = np.nan
            continue
        if not np.isnan(zco[k]):
            j = np.argmin(np.abs(i_ax - th))
            plane = co_t[:, :, j]
            pu = np.real(prior[k])
            pv = np.imag(prior[k])
            if mirror:
                pv = np.abs(pv)
            cost = ((u_tab - pu) / half_a) ** 2 + ((v_tab - pv) / half_z) ** 2 + ((plane - zco[k]) / dsig_co) ** 2
            q = np.argmin(cost)
            sp = ww[q // cost.shape[-1], q % cost.shape[-1]]
            dr = pp[q // cost.shape[-1], q % cost.shape[-1]]
            first = sp * np.exp(1j * np.deg2rad(dr))
            if mirror:
                second = sp * np.exp(1j * (np.deg2rad(-dr)))
                e1 = np.angle(prior[k] / first)
                e2 = np.angle(prior[k] / second)
                vec = first if np.abs(e1) <= np.abs(e2) else second
            else:
                vec = first
        else:
            vec = np.nan * 1j
        if not np.isnan(zcr[k]) and not np.isnan(dcr[k]):
            j = np.argmin(np.abs(ix_ax - th))
            col = cr_t[:, j]
            cw = ((wx_ax - np.abs(vec)) / half_w) ** 2.0
            cs = ((col - zcr[k]) / dcr[k]) ** 2.0
            if not np.isnan(np.abs(vec)):
                cx = cs + cw
            else:
                cx = cs
            sx = wx_ax[np.argmin(cx)]
            if not np.isnan(np.abs(vec)):
                ax = np.angle(vec)
            else:
                ax = 0
            both = sx * np.exp(1j * ax)
        else:
            both = np.nan * 1j
        res_co[k] = vec
        res_x[k] = both



def make_inverter(co_lut, inc_grid, wspd_grid, phi_grid, cr_lut, inc_cr_grid, wspd_cr_grid, dsig_co=0.1,
                  parallel=True, python=False):
    """co_lut [inc][wspd][phi] dB or None; cr_lut [inc][wspd] dB or None.  Returns f(inc, s_co_db, s_cr_db,
    dsig_cr, anc) -> (wind_co, wind_dual)."""
    import types

    from numba import complex128, float64, guvectorize, void

    half_a = 2  # d_antenna, d_azi, dwspd_fg (windspeed.py:139-141)
    half_z = 2
    half_w = 2
    if co_lut is not None:
        co_t = np.ascontiguousarray(np.transpose(np.asarray(co_lut, dtype=np.float64), (1, 2, 0)))
        w_ax = np.asarray(wspd_grid, dtype=np.float64)
        p_ax = np.asarray(phi_grid, dtype=np.float64)
        i_ax = np.asarray(inc_grid, dtype=np.float64)
        mirror = bool((180 - (p_ax[-1] - p_ax[0])) < 2)
    else:
        co_t = np.array([[[]]], dtype=np.float64)
        w_ax = np.array([], dtype=np.float64)
        p_ax = np.array([], dtype=np.float64)
        i_ax = np.array([], dtype=np.float64)
        mirror = False
    pp, ww = np.meshgrid(p_ax, w_ax)
    u_tab = ww * np.cos(np.radians(pp))
    v_tab = ww * np.sin(np.radians(pp))
    if cr_lut is not None:
        cr_t = np.ascontiguousarray(np.transpose(np.asarray(cr_lut, dtype=np.float64), (1, 0)))
        wx_ax = np.asarray(wspd_cr_grid, dtype=np.float64)
        ix_ax = np.asarray(inc_cr_grid, dtype=np.float64)
    else:
        cr_t = np.array([[]], dtype=np.float64)
        wx_ax = np.array([], dtype=np.float64)
        ix_ax = np.array([], dtype=np.float64)

    per_line = types.FunctionType(_per_line.__code__, dict(
        np=np, i_ax=i_ax, co_t=co_t, mirror=mirror, u_tab=u_tab, v_tab=v_tab, half_a=half_a, half_z=half_z, half_w=half_w,
        dsig_co=dsig_co, ww=ww, pp=pp, ix_ax=ix_ax, cr_t=cr_t, wx_ax=wx_ax, range=range, len=len), "per_line")

    if python:
        def run(inc, s_co, s_cr, dcr, anc):
            shp = np.shape(inc)
            a = [np.ascontiguousarray(v).reshape(-1) for v in (inc, s_co, s_cr, dcr, anc)]
            o1 = np.empty(a[0].shape, np.complex128)
            o2 = np.empty(a[0].shape, np.complex128)
            with np.errstate(all="ignore"):
                per_line(*a, o1, o2)
            return o1.reshape(shp), o2.reshape(shp)

        return run
    return guvectorize(
        [void(float64[:], float64[:], float64[:], float64[:], complex128[:], complex128[:], complex128[:])],
        "(n),(n),(n),(n),(n)->(n),(n)",
        fastmath={"nnan": False},
        target="parallel" if parallel else "cpu",
    )(per_line)
