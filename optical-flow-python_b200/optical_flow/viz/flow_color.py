"""Middlebury colour-wheel coding of a flow field (reference: viz/flow_color.py:5-107), computed on the device."""
import numpy as np

from optical_flow import _lib


def flow_to_color(flow, max_flow=None):
    """(H, W, 2) or (B, H, W, 2) flow -> uint8 RGB of the same leading shape.  max_flow=None normalises each field by its
    own largest known magnitude; pixels with |u| or |v| > 1e9 (unknown flow) are black."""
    flow = _lib.f64(flow)
    single = flow.ndim == 3
    if single:
        flow = flow[None]
    if flow.ndim != 4 or flow.shape[3] != 2:
        raise ValueError(f"Flow must be (H, W, 2) or (B, H, W, 2), got shape {flow.shape}")
    B, H, W = flow.shape[:3]
    img = np.empty((B, H, W, 3), dtype=np.uint8)
    _lib.default_context().call("b200flow_flow_to_color", _lib.ptr(flow), B, H, W,
                                -1.0 if max_flow is None else float(max_flow), _lib.ptr(img))
    return img[0] if single else img
