"""CPU ORACLE -- test infrastructure, NOT the product.

A NumPy/SciPy restatement of the coarse-to-fine HS / BA / Classic+NL hot path of
jordanshivers/optical-flow-python, written from the algorithm (not copied), every function citing
the reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the product
(`optical-flow-python_b200/`) never does and has no CPU fallback.

Parity PINNED: `tests/test_oracle_vs_golden.py` checks every function here against fixtures that
`tests/golden/gen_golden.py` produced by running the unmodified reference in the build container
(numpy 2.3.5 / scipy 1.18.1).  Third-party arithmetic the reference leans on and that is not under
/root/reference: scipy (unpinned `scipy>=1.7`, installed 1.18.1) -- `sparse.linalg.spsolve` is used
here as-is (SuperLU); `ndimage.correlate`, `ndimage.map_coordinates` (cubic B-spline with mirror
prefilter) and `ndimage.median_filter` are RESTATED below from their published algorithms so that
the oracle spells out exactly what the CUDA kernels implement; the tests cross-check those
restatements against scipy itself.

Layout conventions used throughout (same as the CUDA side): images are row-major (H, W) planes,
uv is (H, W, 2); a linear system is kept matrix-free as coefficient planes
    a11, a12, a22          data term  d*Ix^2, d*Ix*Iy, d*Iy^2
    wuh, wuv, wvh, wvv     lambda-scaled IRLS edge weights; w?h[i,j] joins (i,j)-(i,j+1) and is 0 in
                           the last column, w?v[i,j] joins (i,j)-(i+1,j) and is 0 in the last row
    bu, bv                 right-hand side
which is SURVEY.md section 3.5's reading of flow_operator (classic_nl.py:279-378, ba.py:208-302,
hs.py:144-203).
"""
import math

import numpy as np
from scipy import sparse
from scipy.sparse.linalg import spsolve, cg, LinearOperator

DERIV5 = np.array([1.0, -8.0, 0.0, 8.0, -1.0]) / 12.0      # base.py:30

# ----------------------------------------------------------------------------------------------
# boundary helpers
# ----------------------------------------------------------------------------------------------


def reflect_index(i, n):
    """scipy 'reflect' (half-sample symmetric, d c b a | a b c d | d c b a) -- SURVEY App. A.1."""
    i = np.asarray(i)
    p = 2 * n
    i = np.mod(i, p)
    return np.where(i >= n, p - 1 - i, i)


def mirror_index(i, n):
    """whole-sample mirror (d c b | a b c d | c b a): numpy.pad 'reflect' and scipy 'mirror'."""
    i = np.asarray(i)
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def correlate_reflect(img, kern):
    """scipy.ndimage.correlate(img, kern, mode='reflect') for odd kernels: out[y,x] = sum k[a,b] *
    img[y+a-cy, x+b-cx] (correlation, no flip) -- derivatives.py:84-86, pyramid.py:64-68."""
    kern = np.atleast_2d(np.asarray(kern, dtype=float))
    H, W = img.shape
    kh, kw = kern.shape
    cy, cx = kh // 2, kw // 2
    rows = [reflect_index(np.arange(H) + a - cy, H) for a in range(kh)]
    cols = [reflect_index(np.arange(W) + b - cx, W) for b in range(kw)]
    out = np.zeros((H, W))
    for a in range(kh):
        ra = img[rows[a], :]
        for b in range(kw):
            if kern[a, b] != 0.0:
                out += kern[a, b] * ra[:, cols[b]]
    return out


# ----------------------------------------------------------------------------------------------
# a1: colour conversion (interface.py:74-141)
# ----------------------------------------------------------------------------------------------


def rgb2gray(im):
    """double(rgb2gray(uint8(im))) with round-half-up at both quantisation steps (interface.py:74-88)."""
    if im.ndim == 2:
        return im
    q = np.clip(np.floor(im + 0.5), 0, 255).astype(np.uint8).astype(float)
    g = 0.2989 * q[:, :, 0] + 0.5870 * q[:, :, 1] + 0.1140 * q[:, :, 2]
    return np.floor(g + 0.5)


def rgb2lab(im):
    """BT.709 / D65 Lab, threshold 0.008856 (interface.py:91-141)."""
    im = np.asarray(im, dtype=float)
    r, g, b = im[:, :, 0], im[:, :, 1], im[:, :, 2]
    if r.max() > 1.0 or g.max() > 1.0 or b.max() > 1.0:
        r, g, b = r / 255.0, g / 255.0, b / 255.0
    m = np.array([[0.412453, 0.357580, 0.180423],
                  [0.212671, 0.715160, 0.072169],
                  [0.019334, 0.119193, 0.950227]])
    shp = r.shape
    xyz = m @ np.array([r.ravel(), g.ravel(), b.ravel()])
    x, y, z = xyz[0] / 0.950456, xyz[1], xyz[2] / 1.088754
    t = 0.008856

    def f(c):
        big = c > t
        return big * c ** (1.0 / 3.0) + (~big) * (7.787 * c + 16.0 / 116.0)

    y3 = y ** (1.0 / 3.0)
    yt = y > t
    fx, fy, fz = f(x), yt * y3 + (~yt) * (7.787 * y + 16.0 / 116.0), f(z)
    L = yt * (116.0 * y3 - 16.0) + (~yt) * (903.3 * y)
    return np.stack([L.reshape(shp), (500.0 * (fx - fy)).reshape(shp), (200.0 * (fy - fz)).reshape(shp)], axis=2)


# ----------------------------------------------------------------------------------------------
# a3/a4/a5/a6: pre-processing
# ----------------------------------------------------------------------------------------------


def scale_image(im, lo, hi):
    """global (all channels jointly) min/max affine map; constant image -> mid value (image_processing.py:6-26)."""
    im = np.asarray(im, dtype=float)
    a, b = im.min(), im.max()
    if a == b:
        return np.full_like(im, (lo + hi) / 2.0)
    return (im - a) / (b - a) * (hi - lo) + lo


def gaussian_kernel(size, sigma):
    """fspecial('gaussian') (image_processing.py:29-49)."""
    r = (size - 1) / 2.0
    ax = np.arange(size) - r
    k = np.exp(-(ax[None, :] ** 2 + ax[:, None] ** 2) / (2 * sigma ** 2))
    k[k < np.finfo(float).eps * k.max()] = 0
    return k / k.sum()


def _rof_div(px, py):
    d = np.zeros_like(px)
    d[:, 1:] += px[:, 1:] - px[:, :-1]
    d[:, 0] += px[:, 0]
    d[1:, :] += py[1:, :] - py[:-1, :]
    d[0, :] += py[0, :]
    return d


def rof_structure(im, theta, iters):
    """Chambolle dual iterations (image_processing.py:86-136)."""
    px = np.zeros_like(im)
    py = np.zeros_like(im)
    delta = 1.0 / (4.0 * theta)
    for _ in range(iters):
        u = im + theta * _rof_div(px, py)
        gx = np.zeros_like(im)
        gy = np.zeros_like(im)
        gx[:, :-1] = u[:, 1:] - u[:, :-1]
        gy[:-1, :] = u[1:, :] - u[:-1, :]
        px = px + delta * gx
        py = py + delta * gy
        nrm = np.maximum(np.sqrt(px ** 2 + py ** 2), 1.0)
        px = px / nrm
        py = py / nrm
    return im + theta * _rof_div(px, py)


def rof_texture(im, theta=1.0 / 8, iters=100, alp=0.95):
    """structure_texture_decomposition_rof (image_processing.py:52-83): joint [-1,1] normalisation,
    per-channel ROF, texture = norm - alp*structure rescaled jointly to [0,255]."""
    n = scale_image(im, -1, 1)
    if n.ndim == 2:
        s = rof_structure(n, theta, iters)
    else:
        s = np.stack([rof_structure(n[:, :, c], theta, iters) for c in range(n.shape[2])], axis=2)
    return scale_image(n - alp * s, 0, 255)


def matlab_round(x):
    return int(math.floor(x + 0.5))


def resize_coords(n_out, n_in):
    """(o + 0.5)/scale - 0.5 clipped to [0, n_in-1], scale = n_out/n_in (pyramid.py:21-31)."""
    scale = n_out / n_in
    c = (np.arange(n_out) + 0.5) / scale - 0.5
    return np.clip(c, 0, n_in - 1)


def bilinear_resize(img, new_h, new_w):
    """separable bilinear gather at resize_coords -- map_coordinates(order=1, mode='nearest')
    semantics: value = (1-ty)(1-tx) z00 + ... with clamped +1 neighbours (pyramid.py:33-40)."""
    H, W = img.shape[:2]
    r = resize_coords(new_h, H)
    c = resize_coords(new_w, W)
    r0 = np.floor(r).astype(int)
    c0 = np.floor(c).astype(int)
    tr = (r - r0)[:, None]
    tc = (c - c0)[None, :]
    r1 = np.minimum(r0 + 1, H - 1)
    c1 = np.minimum(c0 + 1, W - 1)
    if img.ndim == 3:
        tr = tr[:, :, None]
        tc = tc[:, :, None]
    z00 = img[r0][:, c0]
    z01 = img[r0][:, c1]
    z10 = img[r1][:, c0]
    z11 = img[r1][:, c1]
    return (1 - tr) * ((1 - tc) * z00 + tc * z01) + tr * ((1 - tc) * z10 + tc * z11)


def pyramid_kernel(spacing):
    """base.py:185-188: sigma = sqrt(spacing)/sqrt(2), size = 2*round(1.5 sigma)+1 (banker's round)."""
    sigma = math.sqrt(spacing) / math.sqrt(2)
    return gaussian_kernel(int(2 * round(1.5 * sigma) + 1), sigma)


def level_size(n, ratio):
    return max(1, matlab_round(n * ratio))


def build_pyramid(img, levels, spacing):
    """base.py:174-190 + pyramid.py:44-73: level 0 exact copy; each next level = correlate(prev, G,
    'reflect') then bilinear resample to floor(n/spacing + .5)."""
    g = pyramid_kernel(spacing)
    ratio = 1.0 / spacing
    out = [img.copy()]
    cur = img
    for _ in range(1, levels):
        if cur.ndim == 2:
            sm = correlate_reflect(cur, g)
        else:
            sm = np.stack([correlate_reflect(cur[:, :, c], g) for c in range(cur.shape[2])], axis=2)
        cur = bilinear_resize(sm, level_size(cur.shape[0], ratio), level_size(cur.shape[1], ratio))
        out.append(cur)
    return out


def auto_pyramid_levels(H, W, spacing):
    """base.py:192-195."""
    return 1 + int(math.floor(math.log(min(H, W) / 16.0) / math.log(spacing)))


def resample_flow(uv, size):
    """warping.py:6-45: bilinear resize, BOTH components scaled by the height ratio."""
    H, W = uv.shape[:2]
    nh, nw = size
    if (H, W) == (nh, nw):
        return uv.copy()
    return bilinear_resize(uv, nh, nw) * (nh / H)


# ----------------------------------------------------------------------------------------------
# a7/a8: warping + derivatives (derivatives.py)
# ----------------------------------------------------------------------------------------------


def deriv_x(im, h=DERIV5):
    return correlate_reflect(im, np.asarray(h).reshape(1, -1))


def deriv_y(im, h=DERIV5):
    return correlate_reflect(im, np.asarray(h).reshape(-1, 1))


def hermite_warp(Z, x1, y1, h=DERIV5):
    """interp2_bicubic (derivatives.py:27-145) in closed form: tensor-product
    cubic Hermite on the unit cell from Z, DX, DY, DXY at the 4 (clamped) corners.  x1,y1 are
    1-based.  Returns value (NaN when out of bounds), d/dx, d/dy."""
    H, W = Z.shape
    DX = deriv_x(Z, h)
    DY = deriv_y(Z, h)
    DXY = correlate_reflect(Z, np.outer(h, h))
    fx = np.floor(x1).astype(int)
    fy = np.floor(y1).astype(int)
    oob = (fx < 1) | (fx + 1 > W) | (fy < 1) | (fy + 1 > H)
    x0 = np.clip(fx, 1, W) - 1
    x1i = np.clip(fx + 1, 1, W) - 1
    y0 = np.clip(fy, 1, H) - 1
    y1i = np.clip(fy + 1, 1, H) - 1
    ax = np.where(oob, 0.0, x1 - np.floor(x1))
    ay = np.where(oob, 0.0, y1 - np.floor(y1))

    def basis(t):
        t2, t3 = t * t, t * t * t
        Hf = (2 * t3 - 3 * t2 + 1, -2 * t3 + 3 * t2)
        Gf = (t3 - 2 * t2 + t, t3 - t2)
        Hd = (6 * t2 - 6 * t, -6 * t2 + 6 * t)
        Gd = (3 * t2 - 4 * t + 1, 3 * t2 - 2 * t)
        return Hf, Gf, Hd, Gd

    Hx, Gx, Hxd, Gxd = basis(ax)
    Hy, Gy, Hyd, Gyd = basis(ay)
    xs = (x0, x1i)
    ys = (y0, y1i)
    val = np.zeros(x1.shape)
    ddx = np.zeros(x1.shape)
    ddy = np.zeros(x1.shape)
    for a in (0, 1):
        for b in (0, 1):
            z, dx, dy, dxy = Z[ys[b], xs[a]], DX[ys[b], xs[a]], DY[ys[b], xs[a]], DXY[ys[b], xs[a]]
            val += z * Hx[a] * Hy[b] + dx * Gx[a] * Hy[b] + dy * Hx[a] * Gy[b] + dxy * Gx[a] * Gy[b]
            ddx += z * Hxd[a] * Hy[b] + dx * Gxd[a] * Hy[b] + dy * Hxd[a] * Gy[b] + dxy * Gxd[a] * Gy[b]
            ddy += z * Hx[a] * Hyd[b] + dx * Gx[a] * Hyd[b] + dy * Hx[a] * Gyd[b] + dxy * Gx[a] * Gyd[b]
    val[oob] = np.nan
    return val, ddx, ddy


SPLINE_POLE = math.sqrt(3.0) - 2.0


def bspline_prefilter_axis(c, axis):
    """scipy.ndimage.spline_filter1d(order=3, mode='mirror') restated (scipy 1.18.1 ni_splines.c
    apply_filter/_init_causal_mirror/_init_anticausal_mirror).  map_coordinates
    with mode='constant' uses exactly this prefilter (no pre-padding)."""
    c = np.moveaxis(np.array(c, dtype=float), axis, 0)
    n = c.shape[0]
    if n == 1:
        return np.moveaxis(c, 0, axis)
    z = SPLINE_POLE
    c *= (1.0 - z) * (1.0 - 1.0 / z)
    zn1 = z ** (n - 1)
    c0 = c[0] + zn1 * c[n - 1]
    zi = z
    for i in range(1, n - 1):
        c0 = c0 + zi * (c[i] + zn1 * c[n - 1 - i])
        zi *= z
    c[0] = c0 / (1.0 - zn1 * zn1)
    for i in range(1, n):
        c[i] += z * c[i - 1]
    c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1.0)
    for i in range(n - 2, -1, -1):
        c[i] = z * (c[i + 1] - c[i])
    return np.moveaxis(c, 0, axis)


def bspline_prefilter(img):
    """spline_filter over both axes, axis 0 first (scipy _interpolation.py spline_filter)."""
    return bspline_prefilter_axis(bspline_prefilter_axis(img, 0), 1)


def bspline_weights(t):
    """cubic B-spline tap weights at floor-1 .. floor+2 for fractional offset t in [0,1)
    (ni_splines.c get_spline_interpolation_weights, order 3; last weight = 1 - sum of the others)."""
    z = 1.0 - t
    w1 = (t * t * (t - 2.0) * 3.0 + 4.0) / 6.0
    w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0
    w0 = z * z * z / 6.0
    w3 = 1.0 - w0 - w1 - w2
    return w0, w1, w2, w3


def bspline_eval(coef, y0, x0):
    """map_coordinates(order=3, mode='constant', cval=nan) evaluation on prefiltered coefficients:
    NaN iff a coordinate is < 0 or > n-1 (0-based, strict); out-of-range taps index with mirror."""
    H, W = coef.shape
    oob = (x0 < 0) | (x0 > W - 1) | (y0 < 0) | (y0 > H - 1)
    xs = np.where(oob, 0.0, x0)
    ys = np.where(oob, 0.0, y0)
    fx = np.floor(xs)
    fy = np.floor(ys)
    wx = bspline_weights(xs - fx)
    wy = bspline_weights(ys - fy)
    fx = fx.astype(int)
    fy = fy.astype(int)
    out = np.zeros(x0.shape)
    for a in range(4):
        yi = mirror_index(fy + a - 1, H)
        for b in range(4):
            xi = mirror_index(fx + b - 1, W)
            out += coef[yi, xi] * wy[a] * wx[b]
    out[oob] = np.nan
    return out


def bilinear_eval_nan(img, y0, x0):
    """map_coordinates(order=1, mode='constant', cval=nan)."""
    H, W = img.shape
    oob = (x0 < 0) | (x0 > W - 1) | (y0 < 0) | (y0 > H - 1)
    xs = np.where(oob, 0.0, x0)
    ys = np.where(oob, 0.0, y0)
    fx = np.floor(xs).astype(int)
    fy = np.floor(ys).astype(int)
    tx = xs - fx
    ty = ys - fy
    x1 = mirror_index(fx + 1, W)
    y1 = mirror_index(fy + 1, H)
    out = (1 - ty) * ((1 - tx) * img[fy, fx] + tx * img[fy, x1]) + ty * ((1 - tx) * img[y1, fx] + tx * img[y1, x1])
    out[oob] = np.nan
    return out


def partial_deriv(images, uv, interp="cubic", h=DERIV5, blend=0.5):
    """derivatives.py:148-296: warp frame 2 by uv, It = warp - frame 1,
    Ix/Iy = blend*warped-derivative + (1-blend)*frame-1 derivative, zero where out of bounds.
    'bi-cubic' = Hermite (OOB rule floor(x)+1 > W, App. A.4); 'cubic' = B-spline, 'bi-linear'.
    images (H,W,2) -> (H,W) planes; images (H,W,2C), C > 1 (frame-1 channels then frame-2 channels,
    derivatives.py:171-173) -> (H,W,C) planes, every channel treated independently (:208-233, :265-292)."""
    nc = images.shape[2] // 2
    if nc > 1:
        outs = [partial_deriv(np.stack([images[:, :, c], images[:, :, nc + c]], axis=2), uv, interp, h, blend)
                for c in range(nc)]
        return tuple(np.stack([o[k] for o in outs], axis=2) for k in range(3))
    im1, im2 = images[:, :, 0], images[:, :, 1]
    H, W = im1.shape
    xg, yg = np.meshgrid(np.arange(1, W + 1, dtype=float), np.arange(1, H + 1, dtype=float))
    x2 = xg + uv[:, :, 0]
    y2 = yg + uv[:, :, 1]
    i1x, i1y = deriv_x(im1, h), deriv_y(im1, h)
    if interp == "bi-cubic":
        w, wx, wy = hermite_warp(im2, x2, y2, h)
        bad = np.isnan(w)
    elif interp in ("cubic", "bi-linear"):
        bad = (x2 > W) | (x2 < 1) | (y2 > H) | (y2 < 1)
        i2x, i2y = deriv_x(im2, h), deriv_y(im2, h)
        if interp == "cubic":
            ev = lambda p: bspline_eval(bspline_prefilter(p), y2 - 1.0, x2 - 1.0)  # noqa: E731
        else:
            ev = lambda p: bilinear_eval_nan(p, y2 - 1.0, x2 - 1.0)  # noqa: E731
        w, wx, wy = ev(im2), ev(i2x), ev(i2y)
    else:
        raise ValueError("Unknown interpolation method: %s" % interp)
    It = w - im1
    Ix = blend * wx + (1 - blend) * i1x
    Iy = blend * wy + (1 - blend) * i1y
    It[bad] = 0.0
    Ix[bad] = 0.0
    Iy[bad] = 0.0
    return It, Ix, Iy


# ----------------------------------------------------------------------------------------------
# a9: penalties (penalties.py:18-345; kinds in PENALTY_MAP order robust_function.py:16-27)
# ----------------------------------------------------------------------------------------------

PENALTY_KINDS = ["quadratic", "lorentzian", "charbonnier", "generalized_charbonnier", "geman_mcclure",
                 "huber", "tukey", "gaussian", "tdist", "tdist_unnorm"]


def penalty(kind, p, d_type, x):
    x = np.asarray(x, dtype=float)
    p = np.atleast_1d(np.asarray(p, dtype=float))
    s2 = p[0] ** 2
    if d_type not in (0, 1, 2):
        raise ValueError("Unknown d_type: %r" % (d_type,))
    if kind == "quadratic":
        return [x ** 2 / s2, 2 * x / s2, np.full_like(x, 2 / s2)][d_type]
    if kind == "lorentzian":
        return [np.log(1 + x ** 2 / (2 * s2)), 2 * x / (2 * s2 + x ** 2), 2 / (2 * s2 + x ** 2)][d_type]
    if kind == "charbonnier":
        r = np.sqrt(1 + (x / s2) ** 2)
        return [s2 * r, x / (s2 * r), 1 / (s2 * r)][d_type]
    if kind == "generalized_charbonnier":
        a = p[1]
        base = s2 + x ** 2
        return [base ** a, 2 * a * x * base ** (a - 1), 2 * a * base ** (a - 1)][d_type]
    if kind == "geman_mcclure":
        den = s2 + x ** 2
        return [x ** 2 / den, 2 * s2 * x / den ** 2, 2 * s2 / den ** 2][d_type]
    if kind == "huber":
        ax = np.abs(x)
        m = ax <= s2
        return [np.where(m, x ** 2, 2 * s2 * ax - s2 ** 2), np.where(m, 2 * x, 2 * s2 * np.sign(x)),
                np.where(m, 2.0, 2 * s2 / np.maximum(ax, 1e-30))][d_type]
    if kind == "tukey":
        m = np.abs(x) <= p[0]
        om = 1 - x ** 2 / s2
        return [np.where(m, (1 - om ** 3) / 3.0, 1.0 / 3.0), np.where(m, 2 * x * om ** 2 / s2, 0.0),
                np.where(m, 2 * om ** 2 / s2, 0.0)][d_type]
    if kind == "gaussian":
        return [0.5 * math.log(2 * math.pi) + math.log(p[0]) + 0.5 * (x / p[0]) ** 2, x / s2,
                np.full_like(x, 1 / s2)][d_type]
    if kind in ("tdist", "tdist_unnorm"):
        r, s = p[0], p[1]
        s2r = s * s * r
        if d_type == 0:
            c = 0.0
            if kind == "tdist":
                c = math.lgamma(r / 2) - math.lgamma((r + 1) / 2) + 0.5 * math.log(r * math.pi) + math.log(s)
            return (r + 1) / 2 * np.log(1 + x ** 2 / s2r) + c
        return [(r + 1) * x / (s2r + x ** 2), (r + 1) / (s2r + x ** 2)][d_type - 1]
    raise ValueError("Unknown penalty method '%s'" % kind)


# ----------------------------------------------------------------------------------------------
# a10/a11: matrix-free linear system
# ----------------------------------------------------------------------------------------------


def edge_weights(f, pen_h, pen_v, lam):
    """lam * rho'(delta)/delta on forward differences; last column / row carry no edge."""
    wh = np.zeros_like(f)
    wv = np.zeros_like(f)
    wh[:, :-1] = lam * penalty(pen_h[0], pen_h[1], 2, f[:, 1:] - f[:, :-1])
    wv[:-1, :] = lam * penalty(pen_v[0], pen_v[1], 2, f[1:, :] - f[:-1, :])
    return wh, wv


def graph_laplacian(wh, wv, f):
    """sum_q w_pq (f[p] - f[q]) over the 4-neighbourhood (natural Neumann boundary)."""
    out = np.zeros_like(f)
    dh = wh[:, :-1] * (f[:, :-1] - f[:, 1:])
    out[:, :-1] += dh
    out[:, 1:] -= dh
    dv = wv[:-1, :] * (f[:-1, :] - f[1:, :])
    out[:-1, :] += dv
    out[1:, :] -= dv
    return out


def assemble(uv, duv, It, Ix, Iy, spec, alpha):
    """flow_operator + the GNC blend of compute_flow_base (classic_nl.py:238-246,279-378;
    ba.py:172-182,208-302).  spec: dict(rho_su=[(kind,p)]*2, rho_sv=[...]*2, rho_d=(kind,p),
    lam, lam_q, qua_su, qua_sv, qua_d) where qua_* are the quadratic stand-ins."""
    u = uv[:, :, 0] + duv[:, :, 0]
    v = uv[:, :, 1] + duv[:, :, 1]
    multi = It.ndim == 3
    if multi:    # classic_nl.py:330-343 / ba.py:254-267: weights and products are averaged over the channels
        itl = It + Ix * duv[:, :, 0:1] + Iy * duv[:, :, 1:2]
    else:
        itl = It + Ix * duv[:, :, 0] + Iy * duv[:, :, 1]
    parts = []
    if alpha > 0:
        parts.append((alpha, spec["qua_su"], spec["qua_sv"], spec["qua_d"], spec["lam_q"]))
    if alpha < 1:
        parts.append((1 - alpha, spec["rho_su"], spec["rho_sv"], spec["rho_d"], spec["lam"]))
    wuh = wuv = wvh = wvv = d = 0.0
    for wgt, su, sv, pd, lam in parts:
        h_, v_ = edge_weights(u, su[0], su[1], lam)
        wuh, wuv = wuh + wgt * h_, wuv + wgt * v_
        h_, v_ = edge_weights(v, sv[0], sv[1], lam)
        wvh, wvv = wvh + wgt * h_, wvv + wgt * v_
        pw = penalty(pd[0], pd[1], 2, itl)
        d = d + wgt * (pw.mean(axis=2) if multi else pw)
    if multi:
        ix2, ixy, iy2 = (Ix * Ix).mean(axis=2), (Ix * Iy).mean(axis=2), (Iy * Iy).mean(axis=2)
        itx, ity = (itl * Ix).mean(axis=2), (itl * Iy).mean(axis=2)
    else:
        ix2, ixy, iy2, itx, ity = Ix * Ix, Ix * Iy, Iy * Iy, itl * Ix, itl * Iy
    sys = dict(a11=d * ix2, a12=d * ixy, a22=d * iy2, wuh=wuh, wuv=wuv, wvh=wvh, wvv=wvv)
    sys["bu"] = -graph_laplacian(wuh, wuv, uv[:, :, 0]) - d * itx
    sys["bv"] = -graph_laplacian(wvh, wvv, uv[:, :, 1]) - d * ity
    return sys


def assemble_hs(uv, It, Ix, Iy, lam, sigmaD2=1.0, sigmaS2=1.0):
    """hs.py:144-203: d = 1/sigmaD2, unit edge weights scaled by lambda/sigmaS2 (replicate-boundary
    5-point Laplacian == Neumann graph Laplacian)."""
    H, W = It.shape[:2]
    w = lam / sigmaS2
    wh = np.full((H, W), w)
    wh[:, -1] = 0
    wv = np.full((H, W), w)
    wv[-1, :] = 0
    d = 1.0 / sigmaD2
    if It.ndim == 3:     # hs.py:176-181: channel means of the products
        ix2, ixy, iy2 = (Ix * Ix).mean(axis=2), (Ix * Iy).mean(axis=2), (Iy * Iy).mean(axis=2)
        itx, ity = (It * Ix).mean(axis=2), (It * Iy).mean(axis=2)
    else:
        ix2, ixy, iy2, itx, ity = Ix * Ix, Ix * Iy, Iy * Iy, It * Ix, It * Iy
    sys = dict(a11=d * ix2, a12=d * ixy, a22=d * iy2, wuh=wh, wuv=wv, wvh=wh, wvv=wv)
    sys["bu"] = -graph_laplacian(wh, wv, uv[:, :, 0]) - d * itx
    sys["bv"] = -graph_laplacian(wh, wv, uv[:, :, 1]) - d * ity
    return sys


def apply_operator(sys, x):
    """A @ x for x of shape (H, W, 2)."""
    xu, xv = x[:, :, 0], x[:, :, 1]
    au = sys["a11"] * xu + sys["a12"] * xv + graph_laplacian(sys["wuh"], sys["wuv"], xu)
    av = sys["a12"] * xu + sys["a22"] * xv + graph_laplacian(sys["wvh"], sys["wvv"], xv)
    return np.stack([au, av], axis=2)


def operator_diag(sys):
    def deg(wh, wv):
        dg = wh + wv
        dg[:, 1:] += wh[:, :-1]
        dg[1:, :] += wv[:-1, :]
        return dg
    return np.stack([sys["a11"] + deg(sys["wuh"], sys["wuv"]), sys["a22"] + deg(sys["wvh"], sys["wvv"])], axis=2)


def to_sparse(sys):
    """2N x 2N CSC matrix, unknown order [u row-major ; v row-major]."""
    H, W = sys["a11"].shape
    N = H * W

    def block(a, wh, wv):
        dg = a + wh + wv
        dg[:, 1:] += wh[:, :-1]
        dg[1:, :] += wv[:-1, :]
        offh = -wh.ravel()[:-1]
        offv = -wv.ravel()[:-W] if H > 1 else np.zeros(0)
        return sparse.diags([dg.ravel(), offh, offh, offv, offv], [0, 1, -1, W, -W], shape=(N, N))

    c = sparse.diags(sys["a12"].ravel(), 0, shape=(N, N))
    return sparse.bmat([[block(sys["a11"], sys["wuh"], sys["wuv"]), c],
                        [c, block(sys["a22"], sys["wvh"], sys["wvv"])]]).tocsc()


def sor_solve(sys, omega=1.9, max_iters=10000, tol=1e-2):
    """_sor_solve (base.py:138-172): Gauss-Seidel SOR in the reference's unknown order -- all u column-major, then
    all v column-major -- from x = 0, stopping when ||x - x_old|| < tol ||x||.  A cell (y, x) only depends on the
    already-updated cells (y-1, x) and (y, x-1) of its own component, so the cells of one anti-diagonal x + y = d are
    independent: sweeping the anti-diagonals in order reproduces the lexicographic iterates (up to the rounding of
    the row dot product) while staying vectorised."""
    H, W = sys["a11"].shape
    diag = operator_diag(sys)
    xs = np.zeros((H, W, 2))
    comps = ((0, sys["wuh"], sys["wuv"], sys["bu"]), (1, sys["wvh"], sys["wvv"], sys["bv"]))
    waves = []
    for d in range(H + W - 1):
        ys = np.arange(max(0, d - W + 1), min(H - 1, d) + 1)
        waves.append((ys, d - ys))
    for it in range(max_iters):
        old = xs.copy()
        for k, wh, wv, b in comps:
            f = xs[:, :, k]
            other = xs[:, :, 1 - k]
            dg = diag[:, :, k]
            for ys, xx in waves:
                sig = sys["a12"][ys, xx] * other[ys, xx]
                m = xx > 0
                sig[m] -= wh[ys[m], xx[m] - 1] * f[ys[m], xx[m] - 1]
                m = ys > 0
                sig[m] -= wv[ys[m] - 1, xx[m]] * f[ys[m] - 1, xx[m]]
                m = ys + 1 < H
                sig[m] -= wv[ys[m], xx[m]] * f[ys[m] + 1, xx[m]]
                m = xx + 1 < W
                sig[m] -= wh[ys[m], xx[m]] * f[ys[m], xx[m] + 1]
                dd = dg[ys, xx]
                ok = np.abs(dd) >= 1e-15
                new = (1 - omega) * f[ys, xx] + omega * (b[ys, xx] - sig) / np.where(ok, dd, 1.0)
                f[ys, xx] = np.where(ok, new, f[ys, xx])
        if np.linalg.norm(xs - old) < tol * np.linalg.norm(xs):
            break
    return xs


def solve_system(sys, solver="backslash", rtol=1e-3, maxiter=200, sor_max_iters=10000):
    """_solve_linear_system (base.py:87-114): 'backslash' = SuperLU direct solve (reference default);
    'pcg' = scipy cg with Jacobi preconditioner (base.py:116-136); 'sor' = sor_solve above."""
    if solver == "sor":
        return sor_solve(sys, 1.9, sor_max_iters, 1e-2)
    H, W = sys["a11"].shape
    A = to_sparse(sys)
    b = np.concatenate([sys["bu"].ravel(), sys["bv"].ravel()])
    if solver == "backslash":
        x = spsolve(A, b)
    elif solver == "pcg":
        dg = A.diagonal()
        dinv = np.where(np.abs(dg) > 1e-12, 1.0 / dg, 0.0)
        x, _ = cg(A, b, M=LinearOperator(A.shape, matvec=lambda v: dinv * v), maxiter=maxiter, rtol=rtol)
    else:
        raise ValueError("Unknown solver: %s" % solver)
    return np.stack([x[:H * W].reshape(H, W), x[H * W:].reshape(H, W)], axis=2)


# ----------------------------------------------------------------------------------------------
# a13/a14/a15: median, occlusion, weighted median
# ----------------------------------------------------------------------------------------------


def median_filter_reflect(f, kh=5, kw=5):
    """scipy.ndimage.median_filter(size=[kh,kw], mode='reflect'), odd sizes: rank (kh*kw)//2 of the window."""
    H, W = f.shape
    rows = [reflect_index(np.arange(H) + a - kh // 2, H) for a in range(kh)]
    cols = [reflect_index(np.arange(W) + b - kw // 2, W) for b in range(kw)]
    stack = np.stack([f[rows[a]][:, cols[b]] for a in range(kh) for b in range(kw)], axis=0)
    return np.partition(stack, (kh * kw) // 2, axis=0)[(kh * kw) // 2]


def median_uv(uv, size=(5, 5)):
    return np.stack([median_filter_reflect(uv[:, :, 0], size[0], size[1]),
                     median_filter_reflect(uv[:, :, 1], size[0], size[1])], axis=2)


def detect_occlusion(uv, images, sigma_d=0.3, sigma_i=20.0):
    """occlusion.py:6-56: exp(-div^2/2sd^2) * exp(-|I2(p+uv) - I1|^2/2si^2), backward-difference
    divergence (0 in first column/row), bilinear clamp sampling."""
    u, v = uv[:, :, 0], uv[:, :, 1]
    H, W = u.shape
    div = np.zeros_like(u)
    div[:, 1:] += u[:, 1:] - u[:, :-1]
    div[1:, :] += v[1:, :] - v[:-1, :]
    nc = images.shape[2] // 2
    yy, xx = np.mgrid[0:H, 0:W].astype(float)
    x2 = np.clip(xx + u, 0, W - 1)      # map_coordinates mode='nearest' == clamp the coordinate
    y2 = np.clip(yy + v, 0, H - 1)
    fx = np.floor(x2).astype(int)
    fy = np.floor(y2).astype(int)
    tx, ty = x2 - fx, y2 - fy
    x1 = np.minimum(fx + 1, W - 1)
    y1 = np.minimum(fy + 1, H - 1)
    it = np.zeros((H, W))
    for c in range(nc):                 # occlusion.py:47-54: mean over the channels of |warp - frame 1|
        im1, im2 = images[:, :, c], images[:, :, nc + c]
        w2 = (1 - ty) * ((1 - tx) * im2[fy, fx] + tx * im2[fy, x1]) + ty * ((1 - tx) * im2[y1, fx] + tx * im2[y1, x1])
        it += np.abs(w2 - im1)
    if nc > 1:
        it /= nc
    return np.exp(-div ** 2 / (2 * sigma_d ** 2)) * np.exp(-it ** 2 / (2 * sigma_i ** 2))


def weighted_median_filter(uv, color, occ, hsz, sigma_i, rows_per_chunk=8, flip=False):
    """denoise_color_weighted_medfilt2 / _wmedfilt_vectorized (weighted_median.py:24-112): window
    (2hsz+1)^2 with numpy 'reflect' padding (mirror, no edge repeat), weight = max(exp(-|dLab|^2 /
    2 sigma_i^2) * occ_q, 1e-10); per component sort, sequential cumsum, first k with cum >= total/2.
    The reference sorts with np.argsort's default (unstable) kind, so the order of EQUAL flow values -- and with it the
    rounding of the sequential cumsum -- is an implementation detail of NumPy on the host it ran on.  This restatement
    sorts stably; flip=True presents the window in reverse order, i.e. the opposite tie order.  Where the two disagree
    the reference's own answer is summation-order dependent (tests/test_oracle_vs_golden.py::test_weighted_median_tie_fuzz)."""
    H, W = uv.shape[:2]
    if color.ndim == 2:
        color = color[:, :, None]
    k = 2 * hsz + 1
    pad2 = ((hsz, hsz), (hsz, hsz))
    up = np.pad(uv[:, :, 0], pad2, mode="reflect")
    vp = np.pad(uv[:, :, 1], pad2, mode="reflect")
    op = np.pad(occ, pad2, mode="reflect")
    cp = np.pad(color, pad2 + ((0, 0),), mode="reflect")
    swv = np.lib.stride_tricks.sliding_window_view
    inv = 1.0 / (2.0 * sigma_i ** 2)
    out = np.empty((H, W, 2))
    for r0 in range(0, H, rows_per_chunk):
        r1 = min(H, r0 + rows_per_chunk)
        sl = slice(r0, r1 + 2 * hsz)
        uw = swv(up[sl], (k, k)).reshape(r1 - r0, W, k * k)
        vw = swv(vp[sl], (k, k)).reshape(r1 - r0, W, k * k)
        ow = swv(op[sl], (k, k)).reshape(r1 - r0, W, k * k)
        cd = np.zeros((r1 - r0, W, k * k))
        for c in range(color.shape[2]):
            cw = swv(cp[sl, :, c], (k, k)).reshape(r1 - r0, W, k * k)
            cd += (cw - color[r0:r1, :, c][:, :, None]) ** 2
        wgt = np.maximum(np.exp(-cd * inv) * ow, 1e-10)
        if flip:
            uw, vw, wgt = uw[:, :, ::-1], vw[:, :, ::-1], wgt[:, :, ::-1]
        for comp, vals in ((0, uw), (1, vw)):
            order = np.argsort(vals, axis=2, kind="stable")
            vs = np.take_along_axis(vals, order, axis=2)
            cw = np.cumsum(np.take_along_axis(wgt, order, axis=2), axis=2)
            half = cw[:, :, -1:] / 2.0
            idx = np.minimum((cw >= half).argmax(axis=2), k * k - 1)
            out[r0:r1, :, comp] = np.take_along_axis(vs, idx[:, :, None], axis=2)[:, :, 0]
    return out


def weighted_median_admissible(uv, color, occ, hsz, sigma_i, rows_per_chunk=8):
    """Test helper for the decision boundary of weighted_median_1d (weighted_median.py:15-21).  With T the window's total
    weight and S(x) the weight of the samples <= x, the reference returns min{x : S(x) >= T/2} -- evaluated on a
    SEQUENTIAL fp64 cumsum in a sort order that is arbitrary among equal values.  Where some S(x) lies within the
    rounding error of that cumsum (delta = n 2^-52 T) of T/2, the reference's own answer depends on summation order.
    Returns (lo, hi): lo = min{x : S(x) >= T/2 - delta}, hi = min{x : S(x) >= T/2 + delta}, sums in extended precision.
    lo == hi: the answer is order independent and must be reproduced bit-exactly; otherwise any window value in
    [lo, hi] is an answer the reference itself may give."""
    H, W = uv.shape[:2]
    if color.ndim == 2:
        color = color[:, :, None]
    k = 2 * hsz + 1
    pad2 = ((hsz, hsz), (hsz, hsz))
    planes = [np.pad(uv[:, :, 0], pad2, mode="reflect"), np.pad(uv[:, :, 1], pad2, mode="reflect")]
    op = np.pad(occ, pad2, mode="reflect")
    cp = np.pad(color, pad2 + ((0, 0),), mode="reflect")
    swv = np.lib.stride_tricks.sliding_window_view
    inv = 1.0 / (2.0 * sigma_i ** 2)
    lo, hi = np.empty((H, W, 2)), np.empty((H, W, 2))
    for r0 in range(0, H, rows_per_chunk):
        r1 = min(H, r0 + rows_per_chunk)
        sl = slice(r0, r1 + 2 * hsz)
        ow = swv(op[sl], (k, k)).reshape(r1 - r0, W, k * k)
        cd = np.zeros((r1 - r0, W, k * k))
        for c in range(color.shape[2]):
            cw = swv(cp[sl, :, c], (k, k)).reshape(r1 - r0, W, k * k)
            cd += (cw - color[r0:r1, :, c][:, :, None]) ** 2
        wgt = np.maximum(np.exp(-cd * inv) * ow, 1e-10)
        for comp in (0, 1):
            vals = swv(planes[comp][sl], (k, k)).reshape(r1 - r0, W, k * k)
            order = np.argsort(vals, axis=2, kind="stable")
            vs = np.take_along_axis(vals, order, axis=2)
            cwl = np.cumsum(np.take_along_axis(wgt, order, axis=2).astype(np.longdouble), axis=2)
            T = cwl[:, :, -1:]
            delta = (k * k) * np.longdouble(2.0) ** -52 * T
            for dst, thr in ((lo, T / 2 - delta), (hi, T / 2 + delta)):
                idx = np.minimum((cwl >= thr).argmax(axis=2), k * k - 1)
                dst[r0:r1, :, comp] = np.take_along_axis(vs, idx[:, :, None], axis=2)[:, :, 0]
    return lo, hi


def nonlocal_filter(uv, color, occ, area_hsz, mfsz, sigma_i):
    """denoise_color_weighted_medfilt2's dispatch: plain median when no usable colour image."""
    H, W = uv.shape[:2]
    if color is None or color.size < H * W:
        return median_uv(uv, (int(mfsz[0]), int(mfsz[0])))
    return weighted_median_filter(uv, color, occ, area_hsz, sigma_i)


# ----------------------------------------------------------------------------------------------
# a2/a16: presets and drivers
# ----------------------------------------------------------------------------------------------


def _pen(kind, *p):
    return (kind, tuple(float(q) for q in p))


def preset(name):
    """load_of_method (config.py:10-176) + class defaults as a flat dict."""
    base = dict(solver="backslash", pcg_rtol=1e-3, pcg_maxiter=200, deriv_filter=DERIV5, blend=0.5,
                texture=False, median_filter_size=[5, 5], limit_update=True, alp=0.95, pyramid_spacing=2.0,
                gnc_pyramid_levels=2, gnc_pyramid_spacing=1.25, max_linear=1, alpha=1.0, color=False)
    gc = _pen("generalized_charbonnier", 1e-3, 0.45)
    if name in ("classic+nl", "classic+nl-full", "classic+nl-fast"):
        p = dict(base, cls="cnl", texture=True, interp="bi-cubic", gnc_iters=3, max_iters=10,
                 rho_s=gc, rho_d=gc, lam=3.0, lam_q=3.0, area_hsz=7, sigma_i=7.0, color=True)
        if name == "classic+nl-fast":
            p.update(max_iters=3, gnc_iters=2)
        return p
    if name in ("hs-brightness", "hs"):
        lam = 10.0 if name == "hs-brightness" else 40.0
        return dict(base, cls="hs", interp="cubic", texture=(name == "hs"), lam=lam, lam_q=lam,
                    max_warping_iters=10, mf_iter=1, sigmaD2=1.0, sigmaS2=1.0)
    ba = dict(base, cls="ba", interp="cubic", gnc_iters=3, max_iters=10)
    if name == "ba-brightness":
        return dict(ba, rho_s=_pen("lorentzian", 0.1), rho_d=_pen("lorentzian", 3.5), lam=0.045, lam_q=0.045)
    if name in ("ba", "classic-l"):
        return dict(ba, texture=True, rho_s=_pen("lorentzian", 0.03), rho_d=_pen("lorentzian", 1.5), lam=0.06, lam_q=0.06)
    if name in ("classic-c-brightness", "classic-c"):
        lam = 3.0 if name.endswith("brightness") else 5.0
        ch = _pen("charbonnier", 1e-3)
        return dict(ba, texture=(name == "classic-c"), rho_s=ch, rho_d=ch, lam=lam, lam_q=lam)
    if name == "classic++":
        return dict(ba, texture=True, interp="bi-cubic", rho_s=gc, rho_d=gc, lam=3.0, lam_q=3.0)
    raise ValueError("Unknown optical flow method: '%s'" % name)


def _spec(p):
    rs, rd = p["rho_s"], p["rho_d"]
    if p["cls"] == "cnl":       # classic_nl.py:212-226: quadratic with the robust sigma
        qs, qd = _pen("quadratic", rs[1][0]), _pen("quadratic", rd[1][0])
    else:                       # ba.py:150-160: quadratic(1) spatial, quadratic(sigma_d/sigma_s) data
        qs, qd = _pen("quadratic", 1.0), _pen("quadratic", rd[1][0] / rs[1][0])
    return dict(rho_su=[rs, rs], rho_sv=[rs, rs], rho_d=rd, qua_su=[qs, qs], qua_sv=[qs, qs], qua_d=qd,
                lam=p["lam"], lam_q=p["lam_q"])


def _preprocess(p, images):
    if p["texture"]:
        return rof_texture(images, 1.0 / 8, 100, p["alp"])
    return scale_image(images, 0, 255)


def _solve(p, sys):
    return solve_system(sys, p["solver"], p["pcg_rtol"], p["pcg_maxiter"], p.get("sor_max_iters", 10000))


def flow_base_gnc(p, spec, images, color, uv, alpha, trace=None):
    """compute_flow_base of BA (ba.py:140-206) and Classic+NL (classic_nl.py:200-277)."""
    for _ in range(p["max_iters"]):
        duv = np.zeros_like(uv)
        It, Ix, Iy = partial_deriv(images, uv, p["interp"], p["deriv_filter"], p["blend"])
        for _j in range(p["max_linear"]):
            if not 0.0 <= alpha <= 1.0:
                raise ValueError("Invalid GNC alpha: %r" % alpha)
            sys = assemble(uv, duv, It, Ix, Iy, spec, alpha)
            x = _solve(p, sys)
            if p["limit_update"]:
                x = np.clip(x, -1, 1)
            new = uv + x
            if p["median_filter_size"] is not None:
                if p["cls"] == "cnl":
                    occ = detect_occlusion(new, images)
                    new = nonlocal_filter(new, color, occ, p["area_hsz"], p["median_filter_size"], p["sigma_i"])
                else:
                    new = median_uv(new, p["median_filter_size"])
            duv = new - uv
            if trace is not None:
                trace.append(dict(uv_in=uv.copy(), It=It, Ix=Ix, Iy=Iy, x=x, duv=duv.copy()))
        uv = uv + duv
    return uv


def flow_gnc(p, images, color, init=None, trace=None):
    """compute_flow of BA (ba.py:57-138) / Classic+NL (classic_nl.py:89-198)."""
    H, W = images.shape[:2]
    uv = np.zeros((H, W, 2)) if init is None else init.copy()
    pre = _preprocess(p, images)
    levels = auto_pyramid_levels(H, W, p["pyramid_spacing"])
    pyr = build_pyramid(pre, levels, p["pyramid_spacing"])
    gpyr = build_pyramid(pre, p["gnc_pyramid_levels"], p["gnc_pyramid_spacing"])
    cpyr = gcpyr = None
    if p["cls"] == "cnl" and color is not None:
        cpyr = build_pyramid(color, levels, p["pyramid_spacing"])
        gcpyr = build_pyramid(color, p["gnc_pyramid_levels"], p["gnc_pyramid_spacing"])
    spec = _spec(p)
    alpha = p["alpha"]
    for ignc in range(p["gnc_iters"]):
        cur, ccur = (pyr, cpyr) if ignc == 0 else (gpyr, gcpyr)
        nl = levels if ignc == 0 else p["gnc_pyramid_levels"]
        for l in range(nl - 1, -1, -1):
            uv = resample_flow(uv, cur[l].shape[:2])
            q = p if ignc > 0 else dict(p, max_linear=1)
            uv = flow_base_gnc(q, spec, cur[l], None if ccur is None else ccur[l], uv, alpha, trace)
        if p["gnc_iters"] > 1:
            alpha = max(0.0, min(alpha, 1 - (ignc + 1) / (p["gnc_iters"] - 1)))
    return uv


def flow_hs(p, images, init=None, trace=None):
    """HSOpticalFlow.compute_flow / compute_flow_base (hs.py:49-142)."""
    H, W = images.shape[:2]
    uv = np.zeros((H, W, 2)) if init is None else init.copy()
    pre = rof_texture(images) if p["texture"] else scale_image(images, 0, 255)
    levels = auto_pyramid_levels(H, W, p["pyramid_spacing"])
    pyr = build_pyramid(pre, levels, p["pyramid_spacing"])
    for l in range(levels - 1, -1, -1):
        uv = resample_flow(uv, pyr[l].shape[:2])
        for _ in range(p["max_warping_iters"]):
            It, Ix, Iy = partial_deriv(pyr[l], uv, p["interp"], p["deriv_filter"])
            sys = assemble_hs(uv, It, Ix, Iy, p["lam"], p["sigmaD2"], p["sigmaS2"])
            x = _solve(p, sys)
            if trace is not None:
                trace.append(dict(uv_in=uv.copy(), It=It, Ix=Ix, Iy=Iy, x=x))
            if np.linalg.norm(x) < 1e-3:
                break
            if p["limit_update"]:
                x = np.clip(x, -1, 1)
            uv = uv + x
            if p["median_filter_size"] is not None:
                for _k in range(p["mf_iter"]):
                    uv = median_uv(uv, p["median_filter_size"])
    if p["median_filter_size"] is not None:
        uv = median_uv(uv, p["median_filter_size"])
    return uv


def estimate_flow(im1, im2, method="classic+nl-fast", params=None, trace=None):
    """interface.py:11-71."""
    im1 = np.asarray(im1, dtype=float)
    im2 = np.asarray(im2, dtype=float)
    p = preset(method)
    if params:
        for k, v in params.items():
            k = {"lambda": "lam", "lambda_": "lam", "lambda_q": "lam_q", "interpolation_method": "interp"}.get(k, k)
            if k in p:
                p[k] = v
    if im1.ndim == 3 and im1.shape[2] >= 3:
        images = np.stack([rgb2gray(im1), rgb2gray(im2)], axis=2)
    elif im1.ndim == 3:                 # 1-2 channel stacks are concatenated (interface.py:46-52)
        images = np.concatenate([im1, im2], axis=2)
    else:
        images = np.stack([im1, im2], axis=2)
    color = None
    if p["color"]:
        if im1.ndim == 3 and im1.shape[2] >= 3:
            color = rgb2lab(im1)
            for j in range(3):
                color[:, :, j] = scale_image(color[:, :, j], 0, 255)
        else:
            color = im1.copy()
    if p["cls"] == "hs":
        return flow_hs(p, images, trace=trace)
    return flow_gnc(p, images, color, trace=trace)
