"""Stage-level parity of the CUDA kernels (through the C ABI) against the goldens produced by the live reference
(tests/golden/gen_golden.py) and against the oracle on seeded inputs.  Tolerances: bit-exact for the integer-valued
/ selection stages (gray quantisation, pyramid sizes, medians, weighted median), <= 1e-9 absolute (values are
O(1..255), i.e. ~1e-11 relative) for floating-point stages whose summation order differs from NumPy's."""
import numpy as np
import pytest

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu

FP_TOL = 1e-9


@pytest.fixture(scope="module")
def of():
    import optical_flow
    from optical_flow import interface, _lib
    from optical_flow.utils import image_processing, pyramid, warping, derivatives, occlusion, weighted_median
    _lib.default_context()            # fails loudly without a B200
    return dict(pkg=optical_flow, interface=interface, ip=image_processing, pyr=pyramid, warp=warping,
                der=derivatives, occ=occlusion, wm=weighted_median)


def test_rgb2gray_bit_exact(of, stages, crop_rgb):
    np.testing.assert_array_equal(of["interface"]._rgb2gray(crop_rgb[0]), stages["gray1"])
    np.testing.assert_array_equal(of["interface"]._rgb2gray(crop_rgb[1]), stages["gray2"])
    rng = np.random.default_rng(0)
    im = rng.integers(0, 256, (40, 50, 3)).astype(float) + rng.uniform(-0.49, 0.49, (40, 50, 3))
    import flow_oracle as fo
    np.testing.assert_array_equal(of["interface"]._rgb2gray(im), fo.rgb2gray(im))


def test_rgb2lab(of, stages, crop_rgb):
    assert_close(of["interface"]._rgb2lab(crop_rgb[0]), stages["lab_raw"], 1e-10, "lab raw")
    assert_close(of["interface"]._rgb2lab(crop_rgb[0], True), stages["lab_scaled"], 1e-9, "lab scaled")


def test_scale_image(of, stages):
    images = np.stack([stages["gray1"], stages["gray2"]], 2)
    assert_close(of["ip"].scale_image(images, 0, 255), stages["scale_0_255"], 1e-12, "scale 0..255")
    assert_close(of["ip"].scale_image(np.full((4, 5), 3.0), 0, 255), stages["scale_const"], 0, "constant image")


def test_rof_texture(of, stages):
    images = np.stack([stages["gray1"], stages["gray2"]], 2)
    assert_close(of["ip"].structure_texture_decomposition_rof(images, 1 / 8, 7, 0.95), stages["rof_7"], FP_TOL, "rof 7 it")
    assert_close(of["ip"].structure_texture_decomposition_rof(images, 1 / 8, 100, 0.95), stages["rof_100"], FP_TOL, "rof 100 it")


@pytest.mark.parametrize("tag,key", [("tex", "rof_100"), ("lab", "lab_scaled")])
@pytest.mark.parametrize("spacing,levels", [(2.0, 3), (1.25, 2)])
def test_pyramid(of, stages, tag, key, spacing, levels):
    import flow_oracle as fo
    pyr = of["pyr"].compute_image_pyramid(stages[key], fo.pyramid_kernel(spacing), levels, 1.0 / spacing)
    assert len(pyr) == levels
    for l, p in enumerate(pyr):
        want = stages["pyr_%s_%g_%d" % (tag, spacing, l)]
        assert p.shape == want.shape          # integer pyramid: sizes bit-exact
        assert_close(p, want, 0 if l == 0 else FP_TOL, "pyramid %s %g level %d" % (tag, spacing, l))


def test_pyramid_odd_sizes(of, stages):
    import flow_oracle as fo
    for sp, pre in ((2.0, "pyr_odd_"), (1.25, "pyr_odd125_")):
        pyr = of["pyr"].compute_image_pyramid(stages["pyr_odd_in"], fo.pyramid_kernel(sp), 3, 1.0 / sp)
        for l, p in enumerate(pyr):
            assert p.shape == stages[pre + str(l)].shape
            assert_close(p, stages[pre + str(l)], FP_TOL, "odd pyramid %g level %d" % (sp, l))


def test_resample_flow(of, stages):
    assert_close(of["warp"].resample_flow(stages["resample_in_a"], (64, 80)), stages["resample_out_a"], 1e-13, "2x up")
    assert_close(of["warp"].resample_flow(stages["resample_in_b"], (64, 80)), stages["resample_out_b"], 1e-13, "1.25x up")
    assert_close(of["warp"].resample_flow(stages["resample_in_b"], (19, 23)), stages["resample_out_c"], 1e-13, "down")
    same = of["warp"].resample_flow(stages["resample_in_a"], (32, 40))
    np.testing.assert_array_equal(same, stages["resample_in_a"])


@pytest.mark.parametrize("interp", ["bi-cubic", "cubic", "bi-linear"])
@pytest.mark.parametrize("flow", ["zero", "smooth", "big", "int"])
def test_partial_deriv(of, stages, interp, flow):
    uv = stages["pd_uv_" + flow]
    for blend in ((0.5,) if flow != "smooth" else (0.5, 0.3)):
        It, Ix, Iy = of["der"].partial_deriv(stages["rof_100"], uv, interp, np.array([1, -8, 0, 8, -1]) / 12.0, blend)
        key = "pd_%s_%s_%g" % (flow, interp, blend)
        # the out-of-bounds pattern (exact zeros) must be identical, values within FP_TOL
        np.testing.assert_array_equal(It == 0, stages[key + "_It"] == 0)
        assert_close(It, stages[key + "_It"], FP_TOL, key + " It")
        assert_close(Ix, stages[key + "_Ix"], FP_TOL, key + " Ix")
        assert_close(Iy, stages[key + "_Iy"], FP_TOL, key + " Iy")


def test_partial_deriv_bad_interp(of, stages):
    with pytest.raises(ValueError):
        of["der"].partial_deriv(stages["rof_100"], stages["pd_uv_zero"], "nearest")


def test_penalties_all(of, stages):
    from optical_flow.robust.robust_function import RobustFunction
    x = stages["pen_x"]
    for m in ["quadratic", "lorentzian", "charbonnier", "generalized_charbonnier", "geman_mcclure", "huber", "tukey",
              "gaussian", "tdist", "tdist_unnorm"]:
        rf = RobustFunction(m, *stages["pen_%s_p" % m])
        for t, fn in enumerate((rf.evaluate, rf.deriv, rf.deriv_over_x)):
            want = stages["pen_%s_%d" % (m, t)]
            got = fn(x)
            rel = np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want)))
            assert rel <= 1e-13, "%s d_type %d: rel err %.3e" % (m, t, rel)
    with pytest.raises(ValueError):
        RobustFunction("nope", 1.0)
    from optical_flow.robust import penalties
    with pytest.raises(ValueError):
        penalties.quadratic(x, [1.0], 3)
    with pytest.raises(NotImplementedError):
        penalties.mixture(x, [1.0], 0)


def test_median_bit_exact(of, stages):
    np.testing.assert_array_equal(of["wm"].median_filter_uv(stages["med_in"], [5, 5]), stages["med_out"])
    np.testing.assert_array_equal(of["wm"].median_filter_uv(stages["med_in"], [3, 3]), stages["med3_out"])
    np.testing.assert_array_equal(of["wm"].median_filter_uv(stages["med_tiny_in"], [5, 5]), stages["med_tiny_out"])
    import flow_oracle as fo
    rng = np.random.default_rng(5)
    f = rng.standard_normal((37, 71, 2))
    np.testing.assert_array_equal(of["wm"].median_filter_uv(f, [7, 7]), fo.median_uv(f, (7, 7)))
    with pytest.raises(ValueError):
        of["wm"].median_filter_uv(f, [4, 4])


def test_occlusion(of, stages):
    assert_close(of["occ"].detect_occlusion(stages["occ_uv"], stages["rof_100"]), stages["occ_out"], 1e-13, "occ smooth")
    assert_close(of["occ"].detect_occlusion(stages["pd_uv_big"], stages["rof_100"]), stages["occ_big_out"], 1e-13, "occ big")


def test_weighted_median_bit_exact(of, stages):
    wm = of["wm"].denoise_color_weighted_medfilt2
    np.testing.assert_array_equal(wm(stages["wmed_uv"], stages["lab_scaled"], stages["occ_out"], 7, [5, 5], 7), stages["wmed_out"])
    np.testing.assert_array_equal(wm(stages["wmed_uv"], stages["gray1"], stages["occ_big_out"], 7, [5, 5], 7), stages["wmed_gray_out"])
    np.testing.assert_array_equal(wm(stages["wmed_uv"][:40, :36], stages["lab_scaled"][:40, :36], stages["occ_out"][:40, :36],
                                     3, [5, 5], 4.0), stages["wmed_hsz3_out"])
    # no usable colour image -> plain median fallback (weighted_median.py:42-47)
    np.testing.assert_array_equal(wm(stages["med_in"], np.ones((1, 1, 3)), stages["occ_out"], 7, [5, 5], 7), stages["med_out"])


def test_weighted_median_other_windows(of, stages):
    import flow_oracle as fo
    rng = np.random.default_rng(11)
    uv = stages["wmed_uv"][:30, :33]
    col = stages["lab_scaled"][:30, :33]
    occ = rng.uniform(0.05, 1.0, (30, 33))
    for hsz in (1, 2, 4, 5, 9):
        got = of["wm"].denoise_color_weighted_medfilt2(uv, col, occ, hsz, [5, 5], 5.0)
        np.testing.assert_array_equal(got, fo.weighted_median_filter(uv, col, occ, hsz, 5.0), err_msg="hsz=%d" % hsz)


def test_weighted_median_tie_fuzz():
    """The CUDA weighted median on the tie / near-tie fixtures (outputs of the unmodified reference, see
    test_oracle_vs_golden.py::test_weighted_median_tie_fuzz for the criterion): bit-exact wherever the reference's answer
    does not depend on its unstable sort's order of equal values, inside the admissible interval elsewhere, bit-exact
    everywhere when the sums are exact (power-of-two weights)."""
    import flow_oracle as fo
    from optical_flow.utils.weighted_median import denoise_color_weighted_medfilt2
    g = load_golden("wmed_fuzz.npz")
    for i in range(int(g["ncases"])):
        uv, col, occ, hsz = g["c%d_uv" % i], g["c%d_col" % i], g["c%d_occ" % i], int(g["c%d_hsz" % i])
        ref = g["c%d_out" % i]
        got = denoise_color_weighted_medfilt2(uv, col, occ, hsz, [5, 5], 7, False)
        lo, hi = fo.weighted_median_admissible(uv, col, occ, hsz, 7.0)
        assert ((got >= lo) & (got <= hi)).all(), "case %d: %d outside" % (i, int(((got < lo) | (got > hi)).sum()))
        same = lo == hi
        np.testing.assert_array_equal(got[same], ref[same], err_msg="case %d" % i)
        if i == 0:
            np.testing.assert_array_equal(got, ref, err_msg="exact-arithmetic case")


@pytest.mark.parametrize("shape", [(40, 300), (300, 40), (257, 129)])
def test_spline_prefilter_segments_vs_oracle(shape):
    """Cubic-spline partial_deriv on lines longer than one prefilter segment (128 samples + 64 of warm-up, warp.cu): rows,
    columns and both, against the oracle's whole-line recursion (scipy semantics)."""
    import flow_oracle as fo
    from optical_flow.utils.derivatives import partial_deriv
    H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    images = np.round(rng.random((H, W, 2)) * 255)
    yy, xx = np.mgrid[0:H, 0:W].astype(float)
    uv = np.stack([1.3 * np.sin(xx / 17.0) + 0.4, 0.9 * np.cos(yy / 13.0) - 0.2], axis=2)
    h = np.array([1, -8, 0, 8, -1]) / 12.0
    got = partial_deriv(images, uv, "cubic", h, 0.5)
    want = fo.partial_deriv(images, uv, "cubic", h, 0.5)
    for g, w, name in zip(got, want, ("It", "Ix", "Iy")):
        assert_close(g, w, 1e-9, "segmented prefilter %dx%d %s" % (H, W, name))
