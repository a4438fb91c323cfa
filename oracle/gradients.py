"""CPU ORACLE (test infrastructure) for the local-gradient stage -- SURVEY.md section 8 row F4.

Restates `local_gradients` and `R2` of the reference's xsarsea/gradients.py (:588-634, :688-722) with the same
third-party calls the reference makes (cv2.Scharr, scipy.signal.convolve2d / convolve) on plain numpy arrays; the
xarray operations of the reference are replaced by their numpy meaning:
  * `image.coarsen({"line": 2, "sample": 2}, boundary="trim").mean()` -> trim to even sizes, 2x2 blocks, mean that
    skips NaN (xarray reductions default to skipna=True for floating and complex data; an all-NaN block is NaN);
  * `c.where(c <= 1).fillna(0)` -> values that are > 1 or NaN become 0;
  * `xr.merge([np.sqrt(grad2), grad3, c])` -> the three arrays (the first keeps the name "G2").
PARITY UNPINNED BY THE REFERENCE ITSELF for this row: `local_gradients` needs xarray, which cannot be installed here,
so no output of the reference function exists; the arithmetic is pinned to cv2 4.13 / scipy 1.18, i.e. to the libraries
the reference delegates to (tests/golden/gradients.npz, made by tests/golden/make_golden.py::gradients).
"""
import warnings

import numpy as np

B2 = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], float) * 1 / 16   # gradients.py:693


def _smooth_norm(a, kernel):
    """convolve2d(a, k, 'same', 'symm') / convolve2d(ones_like(a), k, 'same', 'symm') (gradients.py:699-701)."""
    from scipy import signal

    num = signal.convolve2d(np.ones_like(a), kernel, mode="same", boundary="symm")
    return signal.convolve2d(a, kernel, mode="same", boundary="symm") / num


def coarsen2(a):
    """2x2 block mean, trailing odd line/sample trimmed, NaN skipped (xarray coarsen(boundary='trim').mean())."""
    h2, w2 = a.shape[0] // 2, a.shape[1] // 2
    blocks = a[:2 * h2, :2 * w2].reshape(h2, 2, w2, 2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmean(blocks, axis=(1, 3))


def R2(a):
    """Reduce by a factor 2 without moire: 5x5 pre-smoothing, 2x2 mean, 3x3 post-smoothing (gradients.py:676-722)."""
    from scipy import signal

    B4 = signal.convolve(B2, B2)
    return _smooth_norm(coarsen2(_smooth_norm(a, B4)), B2)


def local_gradients(image):
    """gradients.py:588-634.  Returns (G2, G3, c): G2 = sqrt of the smoothed squared complex gradient (half size),
    G3 = smoothed |squared gradient|, c = |grad2| / (G3 + 1e-5) with values > 1 or NaN set to 0."""
    import cv2

    image = np.ascontiguousarray(image, dtype=np.float64)
    grad_r = cv2.Scharr(image, cv2.CV_64F, 1, 0)
    grad_i = cv2.Scharr(image, cv2.CV_64F, 0, 1)
    grad12 = (grad_r + 1j * grad_i) ** 2
    grad2 = R2(grad12)
    grad3 = R2(abs(grad12))
    with np.errstate(all="ignore"):
        c = abs(grad2) / (grad3 + 0.00001)
        c = np.where(c <= 1, c, 0.0)
        return np.sqrt(grad2), grad3, c
