"""Warping + spatio-temporal derivatives (reference: utils/derivatives.py:148-296)."""
import ctypes as C

import numpy as np

from optical_flow import _lib

INTERP_CODES = {'bi-cubic': 0, 'cubic': 1, 'bi-linear': 2}


def partial_deriv(images, uv, interp_method='cubic', deriv_filter=None, blend=0.5):
    """It, Ix, Iy of a frame pair warped by uv.  'bi-cubic' = Hermite bicubic with analytic derivatives, 'cubic' =
    scipy-compatible cubic B-spline, 'bi-linear'.  images (H,W,2) -> (H,W) planes; images (H,W,2C) with C > 1 (the C
    channels of frame 1 followed by those of frame 2, derivatives.py:171-173) -> (H,W,C) planes."""
    if interp_method not in INTERP_CODES:
        raise ValueError(f"Unknown interpolation method: {interp_method}")
    images = _lib.f64(images)
    uv = _lib.f64(uv)
    if images.ndim != 3 or images.shape[2] < 2 or images.shape[2] % 2:
        raise ValueError("images must be (H, W, 2C): C channels of frame 1 followed by C channels of frame 2")
    nc = images.shape[2] // 2
    if deriv_filter is None:
        deriv_filter = np.array([1, -8, 0, 8, -1]) / 12.0
    h = _lib.f64(deriv_filter).reshape(-1)
    if h.size != 5:
        raise ValueError("deriv_filter must have 5 taps")
    H, W = images.shape[:2]
    shp = (H, W) if nc == 1 else (H, W, nc)
    It, Ix, Iy = np.empty(shp), np.empty(shp), np.empty(shp)
    _lib.default_context().call("b200flow_partial_deriv_mc", _lib.ptr(images), _lib.ptr(uv), H, W, nc,
                                INTERP_CODES[interp_method], h.ctypes.data_as(C.POINTER(C.c_double)), float(blend),
                                _lib.ptr(It), _lib.ptr(Ix), _lib.ptr(Iy))
    return It, Ix, Iy
