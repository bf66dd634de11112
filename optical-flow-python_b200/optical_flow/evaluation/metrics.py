"""Average angular error / end-point error (parity metric; reference: evaluation/metrics.py:5-53)."""
import numpy as np


def flow_angular_error(tu, tv, u, v, border=0):
    """Returns (AAE degrees, std of AE, AEPE) over pixels whose ground truth is known (|gt| < 1e9)."""
    arrs = [np.asarray(a, dtype=float) for a in (tu, tv, u, v)]
    if border > 0:
        arrs = [a[border:-border, border:-border] for a in arrs]
    tu, tv, u, v = arrs
    known = (np.abs(tu) < 1e9) & (np.abs(tv) < 1e9)
    if not known.all():
        tu, tv, u, v = tu[known], tv[known], u[known], v[known]
    c = (u * tu + v * tv + 1.0) / np.sqrt(u ** 2 + v ** 2 + 1.0) / np.sqrt(tu ** 2 + tv ** 2 + 1.0)
    ae = np.degrees(np.arccos(np.clip(c, -1.0, 1.0)))
    epe = np.hypot(tu - u, tv - v)
    return ae.mean(), ae.std(), epe.mean()


def flow_error_batch(uv, gt, border=0):
    """The same three numbers for B flow fields at once, reduced ON THE DEVICE (b200flow_flow_error): uv, gt are
    (B, H, W, 2) (or (H, W, 2)); returns (B, 4) = AAE degrees, std(AE), AEPE, number of known pixels."""
    from optical_flow import _lib
    uv, gt = _lib.f64(uv), _lib.f64(gt)
    if uv.ndim == 3:
        uv, gt = uv[None], gt[None]
    if uv.shape != gt.shape or uv.ndim != 4 or uv.shape[3] != 2:
        raise ValueError("uv and gt must both be (B, H, W, 2)")
    B, H, W = uv.shape[:3]
    out = np.empty((B, 4))
    _lib.default_context().call("b200flow_flow_error", _lib.ptr(uv), _lib.ptr(gt), B, H, W, int(border), _lib.ptr(out))
    return out
