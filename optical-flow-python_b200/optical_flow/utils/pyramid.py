"""Gaussian image pyramid (reference: utils/pyramid.py:44-73)."""
import ctypes as C

import numpy as np

from optical_flow import _lib


def compute_image_pyramid(img, f, n_levels, ratio):
    """pyramid[0] = exact copy of img; each further level = correlate(prev, f, 'reflect') then MATLAB-convention
    bilinear resize to max(1, floor(n*ratio + 0.5))."""
    img = _lib.f64(img)
    f = _lib.f64(np.atleast_2d(f))
    if f.shape[0] != f.shape[1]:
        raise ValueError("pyramid smoothing kernel must be square, got %r" % (f.shape,))
    H, W = img.shape[:2]
    Cn = 1 if img.ndim == 2 else img.shape[2]
    n_levels = max(1, int(n_levels))
    Hs = (C.c_int * n_levels)()
    Ws = (C.c_int * n_levels)()
    ctx = _lib.default_context()
    ctx.call("b200flow_pyramid", _lib.ptr(img), H, W, Cn, n_levels, _lib.ptr(f), f.shape[0], float(ratio), None, Hs, Ws)
    outs = [np.empty((Hs[l], Ws[l]) + img.shape[2:]) for l in range(n_levels)]
    arr = (C.c_void_p * n_levels)(*[o.ctypes.data for o in outs])
    ctx.call("b200flow_pyramid", _lib.ptr(img), H, W, Cn, n_levels, _lib.ptr(f), f.shape[0], float(ratio), arr, Hs, Ws)
    return outs
