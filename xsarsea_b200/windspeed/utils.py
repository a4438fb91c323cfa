"""Pre-processors feeding `dsig_cr` -- counterpart of xsarsea/windspeed/utils.py:
`get_dsig` (:47-91), `get_dsig_wspd` (:18-44), `nesz_flattening` (:94-163).

The arithmetic runs on the GPU (`xs_dsig`, `xs_dsig_wspd`, `xs_nesz_flatten` of include/xsarsea_b200.h); this module
only marshals.  Containers mirror the reference: numpy in -> numpy out, labelled (xarray) in -> labelled out for the
element-wise functions, `nesz_flattening` always returns a plain array (np.apply_along_axis strips labels,
utils.py:163).  torch CUDA tensors are accepted too and then the result stays on the device, so that
`get_dsig(...)` can feed `xs_invert` without a round trip through host memory.
"""
import logging

import numpy as np

from .. import _device as dev
from .. import _native as nat
from .. import _xr

logger = logging.getLogger("xsarsea.windspeed.utils")
logger.setLevel(logging.INFO)

BLOCK_ELEMS = 1 << 25  # elements per device block for host inputs (0.8 GB of f64 operands)


def _is_tensor(x):
    return type(x).__module__.split(".")[0] == "torch"


def _elementwise(kernel, args):
    """Run `kernel(*device_tensors) -> device tensor` over broadcast `args` (scalars / numpy / labelled / torch CUDA).
    Host inputs are processed in blocks; the result container follows the first labelled input, else numpy (0-d
    inputs give a numpy scalar, like the reference's numpy expressions)."""
    torch = nat.torch_cuda()
    if any(_is_tensor(a) for a in args):
        ts = [a if _is_tensor(a) else torch.as_tensor(np.asarray(a)) for a in args]
        ts = torch.broadcast_tensors(*[t.cuda() for t in ts])
        return kernel(*[t.contiguous() for t in ts])
    template = next((a for a in args if _xr.is_labelled(a)), None)
    arrs = [np.asarray(a.data if _xr.is_labelled(a) else a) for a in args]
    f32 = all(a.dtype == np.float32 for a in arrs)
    arrs = np.broadcast_arrays(*[a.astype(np.float32 if f32 else np.float64, copy=False) for a in arrs])
    shape = arrs[0].shape
    flat = [np.ascontiguousarray(a).reshape(-1) for a in arrs]
    n = flat[0].size
    out = np.empty(n, dtype=np.float64)
    for lo in range(0, n, BLOCK_ELEMS):
        hi = min(n, lo + BLOCK_ELEMS)
        res = kernel(*[torch.from_numpy(f[lo:hi]).cuda() for f in flat])
        out[lo:hi] = res.cpu().numpy()
    out = out.reshape(shape)
    if template is not None and template.shape == shape:
        return _xr.like(template, out, name=getattr(template, "name", None), attrs=dict(template.attrs))
    return out if out.ndim else out[()]


def get_dsig_wspd(name, U_crosspol, SNR_cr):
    """Co/cross blending weight in [0, 1] (utils.py:18-44): sigmoid in the cross-pol wind speed whose centre moves with
    the cross-pol SNR, times a drop-off above 30 m/s.  Names: dsig_wspd_rs2_v3, dsig_wspd_s1_ew_rec_v3,
    dsig_wspd_rcm_v3."""
    try:
        wid = nat.DSIG_WSPD_IDS[name]
    except KeyError:
        # the reference falls through its if/elif chain and fails on an unbound local (utils.py:44)
        raise UnboundLocalError(f"unknown dsig_wspd name {name!r}; known: {sorted(nat.DSIG_WSPD_IDS)}") from None
    return _elementwise(lambda u, s: dev.dsig_wspd(wid, u, s), [U_crosspol, SNR_cr])


def get_dsig(name, inc, sigma0_cr, nesz_cr):
    """dsig_cr value(s) by model name (utils.py:47-91): 'gmf_s1_v2', 'gmf_rs2_v2', 'sarwing_lut_cmodms1ahw',
    'nc_lut_cmodms1ahw'."""
    if name not in nat.DSIG_IDS:
        raise ValueError(
            "dsig names different than 'gmf_s1_v2' or 'gmf_rs2_v2' or 'sarwing_lut_cmodms1ahw' or 'nc_lut_cmodms1ahw' "
            "are not handled. You can compute your own dsig_cr.")
    did = nat.DSIG_IDS[name]
    if did == 0:
        return _elementwise(lambda s, z, i: dev.dsig(did, i, s, z), [sigma0_cr, nesz_cr, inc])
    # `inc` is unused by these formulas (and does not take part in broadcasting, as in the reference)
    return _elementwise(lambda s, z: dev.dsig(did, None, s, z), [sigma0_cr, nesz_cr])


def nesz_flattening(noise, inc):
    """Flatten the noise (nesz, linear, shape (line, sample)) by an order-1 polynomial fit in dB along each line
    (utils.py:94-163): NaNs are first filled with the column mean; the fit uses the column-mean incidence; the
    flattened value is 10**((a*inc + b - 1)/10).  Returns a float64 array (a CUDA tensor for CUDA tensor inputs)."""
    if noise.ndim != 2:
        raise IndexError("Only 2D noise allowed")
    torch = nat.torch_cuda()
    if _is_tensor(noise) or _is_tensor(inc):
        to_t = lambda a: a.cuda() if _is_tensor(a) else torch.from_numpy(np.ascontiguousarray(np.asarray(a))).cuda()
        return dev.nesz_flatten(to_t(noise), to_t(inc).expand(noise.shape))
    noise_v = np.asarray(noise.data if _xr.is_labelled(noise) else noise)
    inc_v = np.broadcast_to(np.asarray(inc.data if _xr.is_labelled(inc) else inc), noise_v.shape)
    f32 = noise_v.dtype == np.float32 and inc_v.dtype == np.float32
    dt = np.float32 if f32 else np.float64
    res = dev.nesz_flatten(torch.from_numpy(np.ascontiguousarray(noise_v, dtype=dt)).cuda(),
                           torch.from_numpy(np.ascontiguousarray(inc_v, dtype=dt)).cuda())
    return res.cpu().numpy()
