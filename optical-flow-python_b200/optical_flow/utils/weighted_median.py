"""Colour / occlusion weighted median of the non-local term (reference: utils/weighted_median.py:24-112)."""
import numpy as np

from optical_flow import _lib


def median_filter_uv(uv, size):
    """scipy.ndimage.median_filter(size, mode='reflect') applied to u and v (hs.py:96-97, ba.py:198-199)."""
    uv = _lib.f64(uv)
    kh, kw = (int(size[0]), int(size[1])) if hasattr(size, '__len__') else (int(size), int(size))
    H, W = uv.shape[:2]
    out = np.empty_like(uv)
    _lib.default_context().call("b200flow_median_filter", _lib.ptr(uv), H, W, kh, kw, _lib.ptr(out))
    return out


def denoise_color_weighted_medfilt2(uv, color_images, occ, area_hsz, mfsz, sigma_i, full_version=False):
    """Window (2*area_hsz+1)^2, weight = max(exp(-|dLab|^2/2 sigma_i^2) * occ, 1e-10), NumPy 'reflect' border.
    `mfsz` is only used by the no-colour fallback (plain median); `full_version` is accepted and ignored, as in
    the reference (weighted_median.py:24,62,67)."""
    uv = _lib.f64(uv)
    H, W = uv.shape[:2]
    if color_images is None or np.size(color_images) < H * W:
        sz = int(mfsz[0]) if hasattr(mfsz, '__len__') else int(mfsz)
        return median_filter_uv(uv, (sz, sz))
    color = _lib.f64(color_images)
    if color.shape[0] != H or color.shape[1] != W:
        raise ValueError("color_images must have the flow's height and width")
    Cn = 1 if color.ndim == 2 else color.shape[2]
    occ = _lib.f64(occ)
    out = np.empty_like(uv)
    _lib.default_context().call("b200flow_weighted_median", _lib.ptr(uv), _lib.ptr(color), _lib.ptr(occ), H, W, Cn,
                                int(area_hsz), float(sigma_i), _lib.ptr(out))
    return out
