"""scale_image / fspecial_gaussian / ROF structure-texture decomposition (reference: utils/image_processing.py)."""
import numpy as np

from optical_flow import _lib


def scale_image(im, vlow, vhigh, ilow=None, ihigh=None):
    """Linear rescale; global min/max over all channels jointly, constant image -> mid value
    (image_processing.py:6-26).  Explicit ilow/ihigh is plain host arithmetic on two scalars."""
    im = _lib.f64(im)
    if ilow is not None or ihigh is not None:
        lo = im.min() if ilow is None else ilow
        hi = im.max() if ihigh is None else ihigh
        if hi == lo:
            return np.full_like(im, (vlow + vhigh) / 2.0)
        return (im - lo) / (hi - lo) * (vhigh - vlow) + vlow
    out = np.empty_like(im)
    _lib.default_context().call("b200flow_scale_image", _lib.ptr(im), im.size, float(vlow), float(vhigh), _lib.ptr(out))
    return out


def fspecial_gaussian(size, sigma):
    """MATLAB fspecial('gaussian') (image_processing.py:29-49): a <= 9x9 host-side constant table."""
    if isinstance(size, (int, np.integer)):
        size = (int(size), int(size))
    m, n = [(s - 1) / 2.0 for s in size]
    y, x = np.ogrid[-m:m + 1, -n:n + 1]
    h = np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    s = h.sum()
    if s != 0:
        h /= s
    return h


def structure_texture_decomposition_rof(im, theta=1.0 / 8, n_iters=100, alp=0.95):
    """ROF structure-texture decomposition, texture scaled to [0,255] (image_processing.py:52-136)."""
    im = _lib.f64(im)
    shp = im.shape
    H, W = shp[:2]
    Cn = 1 if im.ndim == 2 else shp[2]
    out = np.empty_like(im)
    _lib.default_context().call("b200flow_rof_texture", _lib.ptr(im), H, W, Cn, float(theta), int(n_iters), float(alp),
                                _lib.ptr(out))
    return out
