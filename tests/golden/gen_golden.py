#!/usr/bin/env python
"""Generate golden vectors by RUNNING THE UNMODIFIED REFERENCE in this container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden.py [--full] [--only NAME]

The reference (/root/reference, read-only) is imported with a 2-file matplotlib stub
(tests/golden/_stubs) ahead of it on sys.path -- the only thing the stub does is let
`optical_flow/__init__.py` import its viz module.  Nothing of the reference is copied:
this script only calls its public functions / classes and stores inputs + outputs as
compressed .npz fixtures that travel to the GPU box (where /root/reference does not exist).

Files written (tests/golden/):
  stages.npz          per-operator goldens
  systems_<cls>.npz   flow_operator A@probe, b, diag, spsolve(x) for HS / BA / Classic+NL at alpha 1, .5, 0
  e2e_<preset>.npz    estimate_flow() final uv on a 64x80 RubberWhale crop, every in-scope preset
  tape_<preset>.npz   teacher-forcing tape (uv at the start of selected warp iterations + outputs)
  multichannel.npz    two-channel frames: channel-mean data term, occlusion, ROF, end-to-end flows
  eval.npz            flow_angular_error / flow_to_color / write_flo outputs of the reference
  sor.npz             legacy solver='sor' results (tiny windows: the reference's SOR is a per-row Python loop)
  rubberwhale_full.npz (--full) estimate_flow(RubberWhale, 'classic+nl-fast') final uv + AAE/AEPE
  rubberwhale_10_11.npz the two RGB frames + .flo ground truth as uint8 / float32 arrays (fixture data)
"""
import argparse
import contextlib
import io
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(1, "/root/reference")
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402

import optical_flow  # noqa: E402  (the REFERENCE package)
from optical_flow import estimate_flow, load_of_method  # noqa: E402
from optical_flow.interface import _rgb2gray, _rgb2lab  # noqa: E402
from optical_flow.io.flo_io import read_flow_file  # noqa: E402
from optical_flow.evaluation.metrics import flow_angular_error  # noqa: E402
from optical_flow.utils.image_processing import (  # noqa: E402
    scale_image, fspecial_gaussian, structure_texture_decomposition_rof)
from optical_flow.utils.pyramid import compute_image_pyramid  # noqa: E402
from optical_flow.utils.warping import resample_flow  # noqa: E402
from optical_flow.utils.derivatives import partial_deriv  # noqa: E402
from optical_flow.utils.occlusion import detect_occlusion  # noqa: E402
from optical_flow.utils.weighted_median import denoise_color_weighted_medfilt2  # noqa: E402
from optical_flow.robust.robust_function import RobustFunction  # noqa: E402
from optical_flow.methods.base import BaseOpticalFlow  # noqa: E402
from optical_flow.methods import HSOpticalFlow, BAOpticalFlow, ClassicNLOpticalFlow  # noqa: E402
from scipy.ndimage import median_filter, gaussian_filter  # noqa: E402

assert optical_flow.__file__.startswith("/root/reference"), optical_flow.__file__

CROP = (slice(150, 214), slice(230, 310))      # 64 x 80 window of RubberWhale with a motion boundary
CROP_SMALL = (slice(150, 198), slice(230, 294))  # 48 x 64

PRESETS = ["hs-brightness", "hs", "ba-brightness", "ba", "classic-c-brightness", "classic-c",
           "classic++", "classic+nl-fast", "classic+nl"]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def smooth_flow(rng, H, W, mag, sigma=6.0):
    f = gaussian_filter(rng.standard_normal((H, W, 2)), (sigma, sigma, 0))
    f /= np.abs(f).max()
    return f * mag


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote %s (%.1f KB)" % (name, os.path.getsize(path) / 1024.0))


def gen_data():
    im1, im2, tu, tv = read_flow_file("RubberWhale", 10)
    save("rubberwhale_10_11.npz", im1=im1.astype(np.uint8), im2=im2.astype(np.uint8),
         tu=tu.astype(np.float32), tv=tv.astype(np.float32))
    return im1, im2, tu, tv


def gen_stages(im1, im2):
    rng = np.random.default_rng(1234)
    out = {}
    c1 = im1[CROP].copy()
    c2 = im2[CROP].copy()
    H, W = c1.shape[:2]
    out["rgb1"] = c1.astype(np.uint8)
    out["rgb2"] = c2.astype(np.uint8)
    # a1: colour conversion
    g1, g2 = _rgb2gray(c1), _rgb2gray(c2)
    out["gray1"], out["gray2"] = g1, g2
    lab = _rgb2lab(c1)
    out["lab_raw"] = lab.copy()
    for j in range(3):
        lab[:, :, j] = scale_image(lab[:, :, j], 0, 255)
    out["lab_scaled"] = lab
    images = np.stack([g1, g2], axis=2)
    # a3: scale_image
    out["scale_0_255"] = scale_image(images, 0, 255)
    out["scale_const"] = scale_image(np.full((4, 5), 3.0), 0, 255)
    # a4: ROF texture
    out["rof_100"] = structure_texture_decomposition_rof(images, 1.0 / 8, 100, 0.95)
    out["rof_7"] = structure_texture_decomposition_rof(images, 1.0 / 8, 7, 0.95)
    tex = out["rof_100"]
    # a5: pyramids (gaussian kernels as BaseOpticalFlow._build_pyramid makes them)
    ope = HSOpticalFlow()
    for tag, arr in (("tex", tex), ("lab", lab)):
        for sp, lv in ((2.0, 3), (1.25, 2)):
            pyr = ope._build_pyramid(arr, lv, sp)
            for l, p in enumerate(pyr):
                out["pyr_%s_%g_%d" % (tag, sp, l)] = p
    out["gauss_5_1"] = fspecial_gaussian(5, 1.0)
    out["gauss_3_0.79"] = fspecial_gaussian(3, np.sqrt(1.25) / np.sqrt(2))
    odd = rng.random((37, 53, 2)) * 255           # odd sizes exercise the rounding rule
    pyr = ope._build_pyramid(odd, 3, 2.0)
    out["pyr_odd_in"] = odd
    for l, p in enumerate(pyr):
        out["pyr_odd_%d" % l] = p
    pyr = ope._build_pyramid(odd, 3, 1.25)
    for l, p in enumerate(pyr):
        out["pyr_odd125_%d" % l] = p
    # a6: resample_flow
    f_small = smooth_flow(rng, 32, 40, 2.0)
    out["resample_in_a"] = f_small
    out["resample_out_a"] = resample_flow(f_small, (64, 80))
    f_odd = smooth_flow(rng, 51, 64, 2.0)
    out["resample_in_b"] = f_odd
    out["resample_out_b"] = resample_flow(f_odd, (64, 80))
    out["resample_out_c"] = resample_flow(f_odd, (19, 23))   # downsample
    # a7/a8: partial_deriv, three interpolators, flows that leave the image at the border
    h = np.array([1, -8, 0, 8, -1]) / 12.0
    flows = {"zero": np.zeros((H, W, 2)), "smooth": smooth_flow(rng, H, W, 3.0),
             "big": smooth_flow(rng, H, W, 9.0, sigma=3.0),
             "int": np.round(smooth_flow(rng, H, W, 3.0))}
    tex_lvl = tex
    for fname, uv in flows.items():
        out["pd_uv_" + fname] = uv
        for interp in ("bi-cubic", "cubic", "bi-linear"):
            for blend in (0.5,) if fname != "smooth" else (0.5, 0.3):
                It, Ix, Iy = partial_deriv(tex_lvl, uv, interp, h, blend)
                key = "pd_%s_%s_%g" % (fname, interp, blend)
                out[key + "_It"], out[key + "_Ix"], out[key + "_Iy"] = It, Ix, Iy
    # a9: penalties
    x = np.concatenate([np.linspace(-3, 3, 61), [0.0, 1e-8, -1e-8, 1e-3, 25.0, -40.0]])
    out["pen_x"] = x
    specs = {"quadratic": (0.7,), "lorentzian": (0.03,), "charbonnier": (1e-3,),
             "generalized_charbonnier": (1e-3, 0.45), "geman_mcclure": (0.5,), "huber": (0.8,),
             "tukey": (1.7,), "gaussian": (1.3,), "tdist": (5.0, 0.1), "tdist_unnorm": (5.0, 0.1)}
    for m, p in specs.items():
        rf = RobustFunction(m, *p)
        out["pen_%s_p" % m] = np.asarray(p, dtype=float)
        out["pen_%s_0" % m] = rf.evaluate(x)
        out["pen_%s_1" % m] = rf.deriv(x)
        out["pen_%s_2" % m] = rf.deriv_over_x(x)
    # a13: 5x5 median, scipy 'reflect'
    uvm = smooth_flow(rng, H, W, 3.0) + 0.05 * rng.standard_normal((H, W, 2))
    uvm[10:20, 10:30, 0] = 1.25                      # ties
    out["med_in"] = uvm
    out["med_out"] = np.stack([median_filter(uvm[:, :, 0], size=[5, 5], mode="reflect"),
                               median_filter(uvm[:, :, 1], size=[5, 5], mode="reflect")], axis=2)
    out["med3_out"] = np.stack([median_filter(uvm[:, :, 0], size=[3, 3], mode="reflect"),
                                median_filter(uvm[:, :, 1], size=[3, 3], mode="reflect")], axis=2)
    tiny = rng.standard_normal((3, 4, 2))
    out["med_tiny_in"] = tiny
    out["med_tiny_out"] = np.stack([median_filter(tiny[:, :, 0], size=[5, 5], mode="reflect"),
                                    median_filter(tiny[:, :, 1], size=[5, 5], mode="reflect")], axis=2)
    # a14: occlusion
    uvo = flows["smooth"]
    occ = detect_occlusion(uvo, tex_lvl)
    out["occ_uv"], out["occ_out"] = uvo, occ
    occ_big = detect_occlusion(flows["big"], tex_lvl)
    out["occ_big_out"] = occ_big
    # a15: colour/occlusion weighted median (15x15) -- the reference's Python per-pixel loop
    t0 = time.time()
    wm = denoise_color_weighted_medfilt2(uvm, lab, occ, 7, [5, 5], 7, False)
    out["wmed_uv"], out["wmed_out"] = uvm, wm
    wm1 = denoise_color_weighted_medfilt2(uvm, g1, occ_big, 7, [5, 5], 7, False)   # C=1 (gray input quirk 10)
    out["wmed_gray_out"] = wm1
    wm3 = denoise_color_weighted_medfilt2(uvm[:40, :36], lab[:40, :36], occ[:40, :36], 3, [5, 5], 4.0, False)
    out["wmed_hsz3_out"] = wm3
    print("weighted medians took %.1f s" % (time.time() - t0))
    save("stages.npz", **out)
    return out


def _configure(name):
    ope = load_of_method(name)
    ope.display = False
    return ope


def gen_systems(stages):
    """flow_operator goldens: A@probe, b, diag(A) and the direct solve, at the reference's own call sites."""
    rng = np.random.default_rng(99)
    tex = stages["rof_100"]
    gray = stages["scale_0_255"]
    H, W = tex.shape[:2]
    uv = smooth_flow(rng, H, W, 2.0) + 0.02 * rng.standard_normal((H, W, 2))
    probe = rng.standard_normal(2 * H * W)
    h = np.array([1, -8, 0, 8, -1]) / 12.0

    def to_hw2(vec):         # reference vectorisation is column-major [u(:); v(:)]
        return vec.reshape((H, W, 2), order="F")

    def from_hw2(arr):
        return arr.reshape(-1, order="F")

    res = {"uv": uv, "probe": to_hw2(probe)}
    # HS (hs-brightness: cubic spline warp, lambda 10)
    ope = _configure("hs-brightness")
    ope.images = gray
    A, b, _, _ = ope.flow_operator(uv)
    x = ope._solve_linear_system(A, b, uv.shape)
    res.update(hs_Ap=to_hw2(A @ probe), hs_b=to_hw2(b), hs_diag=to_hw2(A.diagonal()), hs_x=x)
    # BA / Classic+NL: reproduce compute_flow_base's GNC blend at alpha = 1, .5, 0
    for tag, preset, imgs in (("ba", "ba", tex), ("cnl", "classic+nl", tex), ("cpp", "classic++", tex),
                              ("cc", "classic-c", tex)):
        ope = _configure(preset)
        ope.images = imgs
        It, Ix, Iy = partial_deriv(imgs, uv, ope.interpolation_method, h, 0.5)
        res[tag + "_It"], res[tag + "_Ix"], res[tag + "_Iy"] = It, Ix, Iy
        import copy
        qua = copy.copy(ope)
        qua.lambda_ = ope.lambda_q
        if isinstance(ope, ClassicNLOpticalFlow):
            qua.rho_spatial_u = [RobustFunction("quadratic", r.param[0]) for r in ope.rho_spatial_u]
            qua.rho_spatial_v = [RobustFunction("quadratic", r.param[0]) for r in ope.rho_spatial_v]
            qua.rho_data = RobustFunction("quadratic", ope.rho_data.param[0])
        else:
            ta = ope.rho_data.param[0] / ope.rho_spatial_u[0].param[0]
            qua.rho_spatial_u = [RobustFunction("quadratic", 1) for _ in ope.rho_spatial_u]
            qua.rho_spatial_v = [RobustFunction("quadratic", 1) for _ in ope.rho_spatial_v]
            qua.rho_data = RobustFunction("quadratic", ta)
        duv = np.zeros_like(uv)
        Aq, bq, _, _ = qua.flow_operator(uv, duv, It, Ix, Iy)
        Ar, br, _, _ = ope.flow_operator(uv, duv, It, Ix, Iy)
        for alpha in (1.0, 0.5, 0.0):
            A = alpha * Aq + (1 - alpha) * Ar
            b = alpha * bq + (1 - alpha) * br
            x = ope._solve_linear_system(A, b, uv.shape)
            k = "%s_a%g" % (tag, alpha)
            res[k + "_Ap"], res[k + "_b"] = to_hw2(A @ probe), to_hw2(b)
            res[k + "_diag"], res[k + "_x"] = to_hw2(A.diagonal()), x
        # max_linear > 1 path: non-zero duv enters both the weights and It_lin
        duv2 = 0.1 * smooth_flow(rng, H, W, 1.0)
        A2, b2, _, _ = ope.flow_operator(uv, duv2, It, Ix, Iy)
        res[tag + "_duv"] = duv2
        res[tag + "_lin_Ap"], res[tag + "_lin_b"] = to_hw2(A2 @ probe), to_hw2(b2)
    save("systems.npz", **res)


class Tape:
    """Records uv at the start of selected warp iterations and that iteration's operator outputs."""

    def __init__(self, every=1):
        self.rec = {}
        self.n = 0
        self.every = every


def gen_e2e(im1, im2, presets):
    c1, c2 = im1[CROP].copy(), im2[CROP].copy()
    for name in presets:
        t0 = time.time()
        uv = quiet(estimate_flow, c1, c2, name)
        extra = {}
        if name == "classic++":       # chaotic at default iteration counts: also the well-posed variant
            extra["uv_maxiters3"] = quiet(estimate_flow, c1, c2, name, {"max_iters": 3})
        if name in ("hs-brightness", "classic+nl-fast"):
            g1 = _rgb2gray(c1)
            g2 = _rgb2gray(c2)
            extra["uv_gray_input"] = quiet(estimate_flow, g1, g2, name)
        if name == "classic+nl-fast":
            extra["uv_pcg_default"] = quiet(estimate_flow, c1, c2, name, {"solver": "pcg"})
        save("e2e_%s.npz" % name.replace("+", "p"), uv=uv, **extra)
        print("  %s: %.1f s, max|uv|=%.3f" % (name, time.time() - t0, np.abs(uv).max()))


def gen_tape(im1, im2, name, params=None, tagsuffix=""):
    """Teacher-forcing tape: patch the names where the drivers bound them."""
    import optical_flow.methods.classic_nl as m_cnl
    import optical_flow.methods.ba as m_ba
    import optical_flow.methods.hs as m_hs
    c1, c2 = im1[CROP_SMALL].copy(), im2[CROP_SMALL].copy()
    steps = []
    cur = {}

    orig_pd = {m: m.partial_deriv for m in (m_cnl, m_ba, m_hs)}
    orig_solve = BaseOpticalFlow._solve_linear_system

    def mk_pd(orig):
        def pd(images, uv, *a, **k):
            It, Ix, Iy = orig(images, uv, *a, **k)
            cur.clear()
            cur.update(images=images.copy(), uv_in=uv.copy(), It=It, Ix=Ix, Iy=Iy)
            return It, Ix, Iy
        return pd

    def solve(self, A, b, uv_shape, x0=None):
        x = orig_solve(self, A, b, uv_shape, x0)
        cur.update(b=b.reshape(uv_shape, order="F"), x=x.copy(), alpha=float(getattr(self, "alpha", 1.0)))
        steps.append(dict(cur))
        return x

    for m in orig_pd:
        m.partial_deriv = mk_pd(orig_pd[m])
    BaseOpticalFlow._solve_linear_system = solve
    try:
        uv = quiet(estimate_flow, c1, c2, name, params)
    finally:
        for m, f in orig_pd.items():
            m.partial_deriv = f
        BaseOpticalFlow._solve_linear_system = orig_solve
    out = {"rgb1": c1.astype(np.uint8), "rgb2": c2.astype(np.uint8), "uv_final": uv, "nsteps": len(steps)}
    # keep every step's uv_in / x (small) and the heavy fields for all steps (48x64 is small enough)
    for i, s in enumerate(steps):
        for k, v in s.items():
            out["s%03d_%s" % (i, k)] = np.asarray(v)
    save("tape_%s%s.npz" % (name.replace("+", "p"), tagsuffix), **out)


def gen_multichannel(im1, im2):
    """Two-channel frames (R, G of the RubberWhale crop): estimate_flow concatenates them into a (H, W, 4) image stack
    (interface.py:46-52) and every data-term quantity becomes a channel mean (derivatives.py:208-233,265-292;
    classic_nl.py:330-343; ba.py:254-267; hs.py:176-181; occlusion.py:47-54)."""
    rng = np.random.default_rng(7)
    c1, c2 = im1[CROP][:, :, :2].astype(float).copy(), im2[CROP][:, :, :2].astype(float).copy()
    H, W = c1.shape[:2]
    images = np.concatenate([c1, c2], axis=2)
    res = {"c1": c1, "c2": c2}
    tex = structure_texture_decomposition_rof(images, 1.0 / 8, 100, 0.95)
    res["rof_100"] = tex
    res["scale_0_255"] = scale_image(images, 0, 255)
    uv = smooth_flow(rng, H, W, 2.0) + 0.02 * rng.standard_normal((H, W, 2))
    uv[:6, :, 0] -= 4.0           # drive some samples out of bounds
    res["uv"] = uv
    h = np.array([1, -8, 0, 8, -1]) / 12.0
    for tag, interp in (("bicubic", "bi-cubic"), ("cubic", "cubic"), ("bilinear", "bi-linear")):
        It, Ix, Iy = partial_deriv(tex, uv, interp, h, 0.5)
        res["pd_%s_It" % tag], res["pd_%s_Ix" % tag], res["pd_%s_Iy" % tag] = It, Ix, Iy
    res["occ"] = detect_occlusion(uv, tex)
    probe = rng.standard_normal(2 * H * W)
    res["probe"] = probe.reshape((H, W, 2), order="F")
    duv = 0.1 * smooth_flow(rng, H, W, 1.0)
    res["duv"] = duv
    for tag, preset in (("cnl", "classic+nl"), ("ba", "ba"), ("hs", "hs")):
        ope = _configure(preset)
        ope.images = tex
        if preset == "hs":
            A, b, _, _ = ope.flow_operator(uv)
        else:
            It, Ix, Iy = partial_deriv(tex, uv, ope.interpolation_method, h, 0.5)
            A, b, _, _ = ope.flow_operator(uv, duv, It, Ix, Iy)
        res[tag + "_Ap"] = (A @ probe).reshape((H, W, 2), order="F")
        res[tag + "_b"] = b.reshape((H, W, 2), order="F")
        res[tag + "_x"] = ope._solve_linear_system(A, b, uv.shape)
    for preset, params in (("hs-brightness", None), ("hs", None), ("ba-brightness", {"max_iters": 3}),
                           ("classic+nl-fast", None)):
        t0 = time.time()
        res["e2e_" + preset] = quiet(estimate_flow, c1, c2, preset, params)
        print("  multichannel %s: %.1f s" % (preset, time.time() - t0))
    save("multichannel.npz", **res)


def gen_sor(im1, im2, stages):
    """Legacy solver='sor' (lexicographic SOR, omega 1.9, tol 1e-2, base.py:138-172): systems on a 24x32 window and
    end-to-end flows on a 32x40 window (the reference's per-row Python loop is too slow for more)."""
    rng = np.random.default_rng(5)
    sl = (slice(10, 34), slice(20, 52))
    tex, gray = stages["rof_100"][sl], stages["scale_0_255"][sl]
    H, W = tex.shape[:2]
    uv = smooth_flow(rng, H, W, 1.5) + 0.02 * rng.standard_normal((H, W, 2))
    h = np.array([1, -8, 0, 8, -1]) / 12.0
    res = {"tex": tex, "gray": gray, "uv": uv}
    ope = _configure("hs-brightness")
    ope.images = gray
    ope.solver = "sor"
    A, b, _, _ = ope.flow_operator(uv)
    res["hs_x"] = ope._solve_linear_system(A, b, uv.shape)
    for tag, preset in (("cnl", "classic+nl"), ("ba", "ba")):
        ope = _configure(preset)
        ope.images = tex
        ope.solver = "sor"
        It, Ix, Iy = partial_deriv(tex, uv, ope.interpolation_method, h, 0.5)
        res[tag + "_It"], res[tag + "_Ix"], res[tag + "_Iy"] = It, Ix, Iy
        A, b, _, _ = ope.flow_operator(uv, np.zeros_like(uv), It, Ix, Iy)
        t0 = time.time()
        res[tag + "_x"] = ope._solve_linear_system(A, b, uv.shape)
        print("  sor %s system: %.1f s" % (tag, time.time() - t0))
    c = (slice(160, 192), slice(240, 280))
    c1, c2 = im1[c].copy(), im2[c].copy()
    res["rgb1"], res["rgb2"] = c1.astype(np.uint8), c2.astype(np.uint8)
    for preset, params in (("hs-brightness", {"solver": "sor"}), ("ba-brightness", {"solver": "sor", "max_iters": 2})):
        t0 = time.time()
        res["e2e_" + preset] = quiet(estimate_flow, c1, c2, preset, params)
        print("  sor e2e %s: %.1f s" % (preset, time.time() - t0))
    save("sor.npz", **res)


def gen_eval(tu, tv):
    """Evaluation / export edges: metrics, Middlebury colour coding, .flo bytes."""
    import tempfile
    from optical_flow.viz.flow_color import flow_to_color
    from optical_flow.io.flo_io import write_flo
    rng = np.random.default_rng(3)
    sl = (slice(100, 196), slice(200, 330))
    gt = np.stack([tu[sl], tv[sl]], axis=2).astype(float)          # contains unknown-flow pixels (1e10)
    H, W = gt.shape[:2]
    est = np.where(np.abs(gt) < 1e9, gt, 0.0) + 0.3 * smooth_flow(rng, H, W, 1.0) + 0.01 * rng.standard_normal((H, W, 2))
    res = {"gt": gt, "est": est}
    for border in (0, 5):
        res["metrics_b%d" % border] = np.array(flow_angular_error(gt[:, :, 0], gt[:, :, 1], est[:, :, 0], est[:, :, 1], border))
    big = est * 4.0
    big[3:9, 4:20] = 1e10                                           # unknown pixels -> black
    res["big"] = big
    res["color_auto"] = flow_to_color(est)
    res["color_max2"] = flow_to_color(big, max_flow=2.0)
    res["color_auto_unknown"] = flow_to_color(big)
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "a.flo")
        write_flo(est, fn)
        res["flo_bytes"] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
    save("eval.npz", **res)


def gen_full(im1, im2, tu, tv):
    t0 = time.time()
    uv = quiet(estimate_flow, im1, im2, "classic+nl-fast")
    dt = time.time() - t0
    aae, std, aepe = flow_angular_error(tu, tv, uv[:, :, 0], uv[:, :, 1], 0)
    print("RubberWhale classic+nl-fast: %.1f s  AAE %.6f STD %.6f AEPE %.6f" % (dt, aae, std, aepe))
    save("rubberwhale_full.npz", uv=uv, aae=aae, std=std, aepe=aepe, seconds=dt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also run full-resolution RubberWhale (about 5 min)")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    im1, im2, tu, tv = gen_data()
    if args.only in (None, "stages", "systems"):
        stages = gen_stages(im1, im2)
        if args.only in (None, "systems"):
            gen_systems(stages)
    if args.only in (None, "e2e"):
        gen_e2e(im1, im2, PRESETS)
    if args.only in (None, "tape"):
        gen_tape(im1, im2, "classic+nl-fast")
        gen_tape(im1, im2, "classic++", {"max_iters": 3}, "_mi3")
        gen_tape(im1, im2, "hs-brightness")
        gen_tape(im1, im2, "ba", {"max_iters": 2}, "_mi2")
    if args.only in (None, "multichannel"):
        gen_multichannel(im1, im2)
    if args.only in (None, "eval"):
        gen_eval(tu, tv)
    if args.only in (None, "sor"):
        gen_sor(im1, im2, dict(np.load(os.path.join(HERE, "stages.npz"))))
    if args.full or args.only == "full":
        gen_full(im1, im2, tu, tv)


if __name__ == "__main__":
    main()
