"""Pins the CPU oracle (oracle/flow_oracle.py) to the reference: every function is checked against fixtures that
tests/golden/gen_golden.py produced by running the unmodified reference (numpy 2.3.5 / scipy 1.18.1), and the
restated scipy pieces (correlate, cubic-spline map_coordinates, median_filter) against scipy itself.  CPU only."""
import numpy as np
import pytest

import flow_oracle as fo
from conftest import assert_close, load_golden, maxabs

TOL = 1e-11      # relative to O(1..255) data: summation-order noise only


def test_colour_conversion(stages, crop_rgb):
    np.testing.assert_array_equal(fo.rgb2gray(crop_rgb[0]), stages["gray1"])
    np.testing.assert_array_equal(fo.rgb2gray(crop_rgb[1]), stages["gray2"])
    lab = fo.rgb2lab(crop_rgb[0])
    assert_close(lab, stages["lab_raw"], 1e-12, "lab")
    for j in range(3):
        lab[:, :, j] = fo.scale_image(lab[:, :, j], 0, 255)
    assert_close(lab, stages["lab_scaled"], 1e-12, "lab scaled")


def test_scale_and_rof(stages):
    images = np.stack([stages["gray1"], stages["gray2"]], 2)
    assert_close(fo.scale_image(images, 0, 255), stages["scale_0_255"], 0, "scale")
    assert_close(fo.scale_image(np.full((4, 5), 3.0), 0, 255), stages["scale_const"], 0, "const")
    assert_close(fo.rof_texture(images, 1 / 8, 7, 0.95), stages["rof_7"], TOL, "rof7")
    assert_close(fo.rof_texture(images, 1 / 8, 100, 0.95), stages["rof_100"], TOL, "rof100")


def test_pyramids_and_resample(stages):
    assert_close(fo.pyramid_kernel(2.0), stages["gauss_5_1"], 1e-16, "gauss 5")
    assert_close(fo.pyramid_kernel(1.25), stages["gauss_3_0.79"], 1e-16, "gauss 3")
    for tag, key in (("tex", "rof_100"), ("lab", "lab_scaled")):
        for sp, lv in ((2.0, 3), (1.25, 2)):
            for l, p in enumerate(fo.build_pyramid(stages[key], lv, sp)):
                assert_close(p, stages["pyr_%s_%g_%d" % (tag, sp, l)], TOL, "pyr")
    for sp, pre in ((2.0, "pyr_odd_"), (1.25, "pyr_odd125_")):
        for l, p in enumerate(fo.build_pyramid(stages["pyr_odd_in"], 3, sp)):
            assert_close(p, stages[pre + str(l)], TOL, "odd pyr")
    assert_close(fo.resample_flow(stages["resample_in_a"], (64, 80)), stages["resample_out_a"], 1e-14, "res a")
    assert_close(fo.resample_flow(stages["resample_in_b"], (64, 80)), stages["resample_out_b"], 1e-14, "res b")
    assert_close(fo.resample_flow(stages["resample_in_b"], (19, 23)), stages["resample_out_c"], 1e-14, "res c")
    assert fo.auto_pyramid_levels(388, 584, 2.0) == 5 and fo.auto_pyramid_levels(1024, 1024, 2.0) == 7
    assert [fo.level_size(n, 0.5) for n in (388, 194, 97, 49)] == [194, 97, 49, 25]
    assert [fo.level_size(n, 0.8) for n in (388, 584)] == [310, 467]


@pytest.mark.parametrize("interp", ["bi-cubic", "cubic", "bi-linear"])
def test_partial_deriv(stages, interp):
    for flow in ("zero", "smooth", "big", "int"):
        for blend in ((0.5,) if flow != "smooth" else (0.5, 0.3)):
            It, Ix, Iy = fo.partial_deriv(stages["rof_100"], stages["pd_uv_" + flow], interp, fo.DERIV5, blend)
            key = "pd_%s_%s_%g" % (flow, interp, blend)
            np.testing.assert_array_equal(It == 0, stages[key + "_It"] == 0)
            assert_close(It, stages[key + "_It"], TOL, key)
            assert_close(Ix, stages[key + "_Ix"], TOL, key)
            assert_close(Iy, stages[key + "_Iy"], TOL, key)


def test_restated_scipy_pieces(stages):
    """correlate / spline prefilter + evaluation / median_filter restatements agree with scipy 1.18.1 itself."""
    from scipy import ndimage
    rng = np.random.default_rng(3)
    img = rng.random((23, 31)) * 255
    k = rng.random((5, 5))
    assert_close(fo.correlate_reflect(img, k), ndimage.correlate(img, k, mode="reflect"), 1e-10, "correlate")
    assert_close(fo.bspline_prefilter(img), ndimage.spline_filter(img, order=3, mode="mirror"), 1e-10, "prefilter")
    ys = rng.uniform(-1.5, 24.5, (40, 40))
    xs = rng.uniform(-1.5, 32.5, (40, 40))
    ys[0, :5] = [0.0, 22.0, 22.0, 5.0, 21.999999]
    xs[0, :5] = [0.0, 30.0, 3.25, 30.0, 29.5]
    want = ndimage.map_coordinates(img, [ys, xs], order=3, mode="constant", cval=np.nan)
    got = fo.bspline_eval(fo.bspline_prefilter(img), ys, xs)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    assert_close(np.nan_to_num(got), np.nan_to_num(want), 1e-10, "cubic map_coordinates")
    want = ndimage.map_coordinates(img, [ys, xs], order=1, mode="constant", cval=np.nan)
    got = fo.bilinear_eval_nan(img, ys, xs)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    assert_close(np.nan_to_num(got), np.nan_to_num(want), 1e-10, "linear map_coordinates")
    for sz in (3, 5, 7):
        np.testing.assert_array_equal(fo.median_filter_reflect(img, sz, sz), ndimage.median_filter(img, size=[sz, sz], mode="reflect"))


def test_penalties(stages):
    x = stages["pen_x"]
    for m in fo.PENALTY_KINDS:
        for t in range(3):
            assert_close(fo.penalty(m, stages["pen_%s_p" % m], t, x), stages["pen_%s_%d" % (m, t)], 1e-12, m)
    with pytest.raises(ValueError):
        fo.penalty("quadratic", [1.0], 3, x)
    with pytest.raises(ValueError):
        fo.penalty("nope", [1.0], 0, x)


def test_filters(stages):
    np.testing.assert_array_equal(fo.median_uv(stages["med_in"]), stages["med_out"])
    np.testing.assert_array_equal(fo.median_uv(stages["med_in"], (3, 3)), stages["med3_out"])
    np.testing.assert_array_equal(fo.median_uv(stages["med_tiny_in"]), stages["med_tiny_out"])
    assert_close(fo.detect_occlusion(stages["occ_uv"], stages["rof_100"]), stages["occ_out"], 1e-14, "occ")
    assert_close(fo.detect_occlusion(stages["pd_uv_big"], stages["rof_100"]), stages["occ_big_out"], 1e-14, "occ big")
    np.testing.assert_array_equal(fo.weighted_median_filter(stages["wmed_uv"], stages["lab_scaled"], stages["occ_out"], 7, 7),
                                  stages["wmed_out"])
    np.testing.assert_array_equal(fo.weighted_median_filter(stages["wmed_uv"], stages["gray1"], stages["occ_big_out"], 7, 7),
                                  stages["wmed_gray_out"])
    np.testing.assert_array_equal(fo.weighted_median_filter(stages["wmed_uv"][:40, :36], stages["lab_scaled"][:40, :36],
                                                            stages["occ_out"][:40, :36], 3, 4.0), stages["wmed_hsz3_out"])


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_linear_systems(systems, stages):
    uv, probe = systems["uv"], systems["probe"]
    It, Ix, Iy = fo.partial_deriv(stages["scale_0_255"], uv, "cubic")
    s = fo.assemble_hs(uv, It, Ix, Iy, 10.0)
    assert _rel(fo.apply_operator(s, probe), systems["hs_Ap"]) < 1e-13
    assert _rel(np.stack([s["bu"], s["bv"]], 2), systems["hs_b"]) < 1e-12
    assert _rel(fo.operator_diag(s), systems["hs_diag"]) < 1e-13
    assert_close(fo.solve_system(s), systems["hs_x"], 1e-10, "hs x")
    for tag, name in (("ba", "ba"), ("cnl", "classic+nl"), ("cpp", "classic++"), ("cc", "classic-c")):
        spec = fo._spec(fo.preset(name))
        d = [systems[tag + k] for k in ("_It", "_Ix", "_Iy")]
        for alpha in (1.0, 0.5, 0.0):
            s = fo.assemble(uv, np.zeros_like(uv), d[0], d[1], d[2], spec, alpha)
            k = "%s_a%g" % (tag, alpha)
            assert _rel(fo.apply_operator(s, probe), systems[k + "_Ap"]) < 1e-13, k
            assert _rel(np.stack([s["bu"], s["bv"]], 2), systems[k + "_b"]) < 1e-10, k
            assert _rel(fo.operator_diag(s), systems[k + "_diag"]) < 1e-13, k
        assert_close(fo.solve_system(s), systems[k + "_x"], 1e-9, k + " x")
        s = fo.assemble(uv, systems[tag + "_duv"], d[0], d[1], d[2], spec, 0.0)
        assert _rel(fo.apply_operator(s, probe), systems[tag + "_lin_Ap"]) < 1e-13
        assert _rel(np.stack([s["bu"], s["bv"]], 2), systems[tag + "_lin_b"]) < 1e-10


@pytest.mark.parametrize("preset", ["hs-brightness", "ba", "classic+nl-fast"])
def test_end_to_end(crop_rgb, preset):
    g = load_golden("e2e_%s.npz" % preset.replace("+", "p"))
    assert_close(fo.estimate_flow(crop_rgb[0], crop_rgb[1], preset), g["uv"], 1e-9, preset)


def test_unknown_names():
    with pytest.raises(ValueError):
        fo.preset("classic-x")
    with pytest.raises(ValueError):
        fo.partial_deriv(np.zeros((8, 8, 2)), np.zeros((8, 8, 2)), "nearest")


# ---------------------------------------------------------------------------------------------------------------
# multi-channel frames: goldens from the live reference on two-channel inputs
# ---------------------------------------------------------------------------------------------------------------
def test_multichannel_oracle_stages():
    import flow_oracle as fo
    g = load_golden("multichannel.npz")
    tex, uv = g["rof_100"], g["uv"]
    images = np.concatenate([g["c1"], g["c2"]], axis=2)
    assert maxabs(fo.rof_texture(images), tex) <= 1e-12
    for tag, interp in (("bicubic", "bi-cubic"), ("cubic", "cubic"), ("bilinear", "bi-linear")):
        It, Ix, Iy = fo.partial_deriv(tex, uv, interp)
        assert It.shape == tex.shape[:2] + (2,)
        for got, key in ((It, "It"), (Ix, "Ix"), (Iy, "Iy")):
            assert maxabs(got, g["pd_%s_%s" % (tag, key)]) <= 1e-11, (tag, key)
    assert maxabs(fo.detect_occlusion(uv, tex), g["occ"]) <= 1e-13


def test_multichannel_oracle_systems_and_e2e():
    import flow_oracle as fo
    g = load_golden("multichannel.npz")
    tex, uv = g["rof_100"], g["uv"]

    def rel(a, b):
        return np.abs(a - b).max() / np.abs(b).max()
    for tag, preset in (("cnl", "classic+nl"), ("ba", "ba")):
        p = fo.preset(preset)
        It, Ix, Iy = fo.partial_deriv(tex, uv, p["interp"])
        sys_ = fo.assemble(uv, g["duv"], It, Ix, Iy, fo._spec(p), 0.0)
        assert rel(fo.apply_operator(sys_, g["probe"]), g[tag + "_Ap"]) <= 1e-10
        assert rel(np.stack([sys_["bu"], sys_["bv"]], 2), g[tag + "_b"]) <= 1e-10
    It, Ix, Iy = fo.partial_deriv(tex, uv, "cubic")
    sys_ = fo.assemble_hs(uv, It, Ix, Iy, fo.preset("hs")["lam"])
    assert rel(fo.apply_operator(sys_, g["probe"]), g["hs_Ap"]) <= 1e-12
    for preset, params in (("hs-brightness", None), ("ba-brightness", {"max_iters": 3}), ("classic+nl-fast", None)):
        got = fo.estimate_flow(g["c1"], g["c2"], preset, params)
        assert maxabs(got, g["e2e_" + preset]) <= 1e-4, preset


def test_sor_oracle_vs_reference():
    """Anti-diagonal (wavefront) SOR == the reference's lexicographic per-row loop (base.py:138-172)."""
    import flow_oracle as fo
    g = load_golden("sor.npz")
    uv = g["uv"]
    It, Ix, Iy = fo.partial_deriv(g["gray"], uv, "cubic")
    assert maxabs(fo.sor_solve(fo.assemble_hs(uv, It, Ix, Iy, 10.0)), g["hs_x"]) <= 1e-11
    for tag, preset in (("cnl", "classic+nl"), ("ba", "ba")):
        p = fo.preset(preset)
        s = fo.assemble(uv, np.zeros_like(uv), g[tag + "_It"], g[tag + "_Ix"], g[tag + "_Iy"], fo._spec(p), 0.0)
        assert maxabs(fo.sor_solve(s), g[tag + "_x"]) <= 1e-12


def test_weighted_median_tie_fuzz():
    """Decision boundary of weighted_median_1d (weighted_median.py:15-21) on fixtures built to sit on it
    (tests/golden/gen_golden_wmfuzz.py, outputs of the unmodified reference).  The reference's np.argsort is unstable, so
    among EQUAL flow values its summation order -- and the last bit of its cumsum -- is an implementation detail; the
    oracle must agree bit-exactly wherever the answer does not depend on that (weighted_median_admissible: lo == hi), stay
    inside [lo, hi] elsewhere, and agree everywhere when every sum is exact (power-of-two weights, case 0)."""
    g = load_golden("wmed_fuzz.npz")
    boundary = 0
    for i in range(int(g["ncases"])):
        args = (g["c%d_uv" % i], g["c%d_col" % i], g["c%d_occ" % i], int(g["c%d_hsz" % i]), 7.0)
        ref = g["c%d_out" % i]
        lo, hi = fo.weighted_median_admissible(*args)
        got = fo.weighted_median_filter(*args)
        assert ((ref >= lo) & (ref <= hi)).all(), "case %d: the reference itself leaves the admissible interval" % i
        assert ((got >= lo) & (got <= hi)).all(), "case %d" % i
        same = lo == hi
        np.testing.assert_array_equal(got[same], ref[same], err_msg="case %d" % i)
        if i == 0:
            np.testing.assert_array_equal(got, ref, err_msg="exact-arithmetic case")
        boundary += int((~same).sum())
    assert boundary > 50, "the fixtures no longer reach the decision boundary"
