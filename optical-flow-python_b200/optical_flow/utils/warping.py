"""Flow resampling between pyramid levels (reference: utils/warping.py:6-45)."""
import numpy as np

from optical_flow import _lib


def resample_flow(uv, target_sz, method='bilinear'):
    """Bilinear resize of (H,W,2) flow to target_sz; both components scaled by the HEIGHT ratio."""
    uv = _lib.f64(uv)
    H, W = uv.shape[:2]
    nh, nw = int(target_sz[0]), int(target_sz[1])
    if (H, W) == (nh, nw):
        return uv.copy()
    out = np.empty((nh, nw, 2))
    _lib.default_context().call("b200flow_resample_flow", _lib.ptr(uv), H, W, nh, nw, _lib.ptr(out))
    return out
