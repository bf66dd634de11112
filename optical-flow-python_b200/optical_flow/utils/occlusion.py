"""Occlusion confidence from flow divergence and brightness constancy (reference: utils/occlusion.py:6-56)."""
import numpy as np

from optical_flow import _lib


def detect_occlusion(uv, images, sigma_d=0.3, sigma_i=20.0):
    uv = _lib.f64(uv)
    images = _lib.f64(images)
    if images.ndim != 3 or images.shape[2] < 2 or images.shape[2] % 2:
        raise ValueError("images must be (H, W, 2C): C channels of frame 1 followed by C channels of frame 2")
    H, W = uv.shape[:2]
    occ = np.empty((H, W))
    _lib.default_context().call("b200flow_detect_occlusion_mc", _lib.ptr(uv), _lib.ptr(images), H, W,
                                images.shape[2] // 2, float(sigma_d), float(sigma_i), _lib.ptr(occ))
    return occ
