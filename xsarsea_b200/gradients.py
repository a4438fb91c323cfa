"""local_gradients -- counterpart of the local-gradient stage of xsarsea/gradients.py (:588-634, with R2 :676-722),
the consumer of `sigma0_detrend`'s output in the wind-streak analysis.

Only this stage is provided (the histogram / direction-selection classes of the reference's gradients module are out
of scope).  The arithmetic runs on the GPU (`xs_local_gradients`): Scharr gradient as a
complex number, squared, reduced by a factor 2 without moire, plus the gradient quality `c`.
"""
from __future__ import annotations

import numpy as np

from . import _device as dev
from . import _native as nat
from . import _xr


def _is_tensor(x):
    return type(x).__module__.split(".")[0] == "torch"


def local_gradients(image):
    """compute local multi_gradients (gradients.py:588-634).

    Parameters
    ----------
    image : 2-D array with dims ('line', 'sample') -- labelled (xarray), numpy, or a torch CUDA tensor.

    Returns
    -------
    dataset with variables
      G2 : complex gradients, half size of `image`; they are square roots of squared gradients, so angles are in
           [-pi/2, pi/2] and a gradient and its negative yield the same value
      G3 : smoothed modulus of the squared gradient
      c  : G2 quality in [0, 1]
    (an xarray.Dataset when xarray is installed and the input is labelled, otherwise a dict-like with attribute access;
    CUDA tensors in -> CUDA tensors out.)
    """
    if _is_tensor(image):
        if image.dim() != 2:
            raise ValueError("local_gradients needs a 2D image with dims ['line', 'sample']")
        nat.torch_cuda()
        g2, g3, c = dev.local_gradients(image.cuda())
        return _xr.DatasetLite(G2=g2, G3=g3, c=c)
    vals = np.asarray(image.data if _xr.is_labelled(image) else image)
    if vals.ndim != 2:
        raise ValueError("local_gradients needs a 2D image with dims ['line', 'sample']")
    torch = nat.torch_cuda()
    dt = np.float32 if vals.dtype == np.float32 else np.float64
    g2, g3, c = (t.cpu().numpy() for t in dev.local_gradients(torch.from_numpy(np.ascontiguousarray(vals, dtype=dt)).cuda()))
    if not _xr.is_labelled(image):
        return _xr.DatasetLite(G2=g2, G3=g3, c=c)
    # coarsen(...).mean() also averages the coordinates of each 2x2 block (trailing odd element trimmed)
    dims = tuple(image.dims)
    coords = {}
    for d, n in zip(dims, g2.shape):
        cv = getattr(image, "coords", {}).get(d) if hasattr(image.coords, "get") else None
        if cv is not None:
            cv = np.asarray(getattr(cv, "values", cv), dtype=np.float64)[:2 * n]
            coords[d] = cv.reshape(n, 2).mean(axis=1)
    out = {name: _xr.make_dataarray(arr, dims, coords=coords, name=name) for name, arr in (("G2", g2), ("G3", g3), ("c", c))}
    return _xr.make_dataset(out)
