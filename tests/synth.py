"""Synthetic inputs of SURVEY.md section 8d: a smooth random texture warped by a known affine flow (max |flow| ~2.4 px
at any size), optionally with a foreground disc translating over it (config 3).  im2(p + uv(p)) = im1(p)."""
import numpy as np
from scipy.ndimage import gaussian_filter, map_coordinates


def affine_flow(h, w):
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    a = 0.004 * 256 / max(h, w)
    b = 0.003 * 256 / max(h, w)
    yy, xx = np.mgrid[0:h, 0:w].astype(float)
    u = 1.5 + a * (xx - cx) - b * (yy - cy)
    v = -0.8 + b * (xx - cx) + a * (yy - cy)
    return np.stack([u, v], axis=2), (a, b, cx, cy)


def gray_pair(h, w, seed, disc=False):
    """float64 gray pair in [0, 255] and its ground-truth flow (H, W, 2)."""
    rng = np.random.default_rng(seed)
    pad = 64
    base = gaussian_filter(rng.random((h + 2 * pad, w + 2 * pad)), 1.5)
    base = (base - base.min()) / (base.max() - base.min()) * 255.0
    flow, (a, b, cx, cy) = affine_flow(h, w)
    M = np.array([[1 + a, -b], [b, 1 + a]])
    Mi = np.linalg.inv(M)
    yy, xx = np.mgrid[0:h, 0:w].astype(float)
    qx, qy = xx - cx - 1.5, yy - cy + 0.8
    px = cx + Mi[0, 0] * qx + Mi[0, 1] * qy
    py = cy + Mi[1, 0] * qx + Mi[1, 1] * qy
    im1 = base[pad:pad + h, pad:pad + w].copy()
    im2 = map_coordinates(base, [py + pad, px + pad], order=3, mode="nearest")
    if disc:
        tex = gaussian_filter(np.random.default_rng(seed + 10).random((h, w)), 1.5)
        tex = (tex - tex.min()) / (tex.max() - tex.min()) * 255.0
        r = h / 6.0
        m1 = (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
        m2 = (xx - 3 - cx) ** 2 + (yy + 2 - cy) ** 2 <= r * r
        im1[m1] = tex[m1]
        sh = map_coordinates(tex, [yy + 2, xx - 3], order=3, mode="nearest")
        im2[m2] = sh[m2]
        flow[m1] = (3.0, -2.0)
    return im1, im2, flow


def interior_epe(uv, flow, margin=8):
    return float(np.sqrt(((uv - flow) ** 2).sum(-1))[margin:-margin, margin:-margin].mean())
