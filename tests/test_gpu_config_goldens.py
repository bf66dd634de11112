"""Reference-pinned parity at the BASELINE.json configs and at the timed bench workload.

The goldens (tests/golden/cfg_*.npz) were produced by tests/golden/gen_golden_configs.py, which RUNS THE UNMODIFIED
REFERENCE on the build host; wall seconds of those runs are in profiles/cpu_fullsize.json.  Synthetic inputs are
regenerated here (tests/synth.py / bench.py, quantised to uint8 grey levels) and checked against the stored crc32, so a
mismatch of the inputs is reported as such, never as a flow error.

Tolerances: north_star's 1e-3 px max-abs for final flows (the goldens are stored as float32: +1e-6 px), 0.5 % for
AAE / AEPE against the Middlebury ground truth, 1e-9 for operator-level quantities (data O(255)), 1e-6 px for the
teacher-forced increments.
  config 1  'classic+nl-fast' on all eight Middlebury sequences with ground truth (RubberWhale 584x388 ... Urban 640x480)
  config 2  'hs-brightness' at 512x512 (the 1024x1024 reference run takes ~1 h: BASELINE.md section 3)
  config 3  'ba' (3 GNC stages x 10 iterations + ROF) at 270x480; cubic-spline partial_deriv at the full 1920x1080
  config 4  is config 1's preset with fullVersion (identical arithmetic in the reference, tests/test_gpu_configs.py)
  config 5  'classic++' max_iters=3 at 540x960 end to end + teacher-forced warp iterations at that size;
            partial_deriv / flow_operator at the full 3840x2160 (64-bit indexing)
  bench     the timed workload: synthetic 640x480 RGB pair seed 3, 'classic+nl-fast', through estimate_flow_batch
"""
import os
import sys
import zlib

import numpy as np
import pytest

from conftest import ROOT, assert_close, load_golden
import synth

pytestmark = pytest.mark.gpu

E2E_TOL = 1e-3 + 1e-6


def crc(*arrs):
    c = 0
    for a in arrs:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return c


def quant_gray_pair(h, w, seed, disc=False):
    im1, im2, flow = synth.gray_pair(h, w, seed, disc=disc)
    q = lambda im: np.clip(np.floor(im + 0.5), 0, 255)  # noqa: E731
    return q(im1), q(im2), flow


def tf_flow(H, W, flow):
    """the analytic teacher-forcing flow of gen_golden_configs.py"""
    yy, xx = np.mgrid[0:H, 0:W].astype(float)
    rip = np.stack([0.3 * np.sin(xx / 37.0) * np.cos(yy / 29.0), 0.25 * np.cos(xx / 41.0 + 0.5) * np.sin(yy / 31.0)], axis=2)
    return flow * 0.9 + rip


def _golden_or_skip(name):
    path = os.path.join(ROOT, "tests", "golden", name)
    if not os.path.exists(path):
        pytest.skip("golden %s not generated" % name)
    return load_golden(name)


def _check_inputs(g, *arrs, key="input_crc"):
    assert crc(*arrs) == int(g[key]), ("regenerated synthetic input differs from the one the reference was run on "
                                       "(numpy/scipy version?) -- the golden cannot be compared")


# ---------------------------------------------------------------------------------------------------------------
def test_bench_workload_vs_reference():
    """The pair bench.py times (seed 3) -- through the same public call, estimate_flow_batch with uint8 RGB frames."""
    sys.path.insert(0, ROOT)
    import bench
    from optical_flow import estimate_flow_batch
    g = _golden_or_skip("cfg_bench640.npz")
    im1, im2, _ = bench.synth_pair(480, 640, 3)
    _check_inputs(g, im1, im2)
    other = bench.synth_pair(480, 640, 4)
    uv = estimate_flow_batch(np.stack([im1, other[0]]), np.stack([im2, other[1]]), "classic+nl-fast")
    assert_close(uv[0], g["uv"].astype(np.float64), E2E_TOL, "bench workload 640x480 seed 3 vs reference")


def test_bench_workload_concurrent_groups_identical():
    """Concurrent sub-batches (b200flow_ctx_set_split) must not change a single bit of any pair's flow."""
    sys.path.insert(0, ROOT)
    import bench
    from optical_flow import _lib, estimate_flow_batch
    pairs = [bench.synth_pair(240, 320, 10 + k) for k in range(4)]
    ims1 = np.stack([p[0] for p in pairs])
    ims2 = np.stack([p[1] for p in pairs])
    ctx = _lib.default_context()
    ctx.set_split(1)
    ref = estimate_flow_batch(ims1, ims2, "classic+nl-fast").copy()
    try:
        for groups in (2,):        # the device holds two 256-thread solver CTAs per SM: two concurrent groups
            ctx.set_split(groups)
            uv = estimate_flow_batch(ims1, ims2, "classic+nl-fast")
            np.testing.assert_array_equal(uv, ref, err_msg="%d concurrent groups" % groups)
    finally:
        ctx.set_split(1)


def test_config2_hs_brightness_512():
    from optical_flow import estimate_flow
    g = _golden_or_skip("cfg_hs512.npz")
    im1, im2, _ = quant_gray_pair(512, 512, 0)
    _check_inputs(g, im1, im2)
    uv = estimate_flow(im1, im2, "hs-brightness")
    assert_close(uv, g["uv"].astype(np.float64), E2E_TOL, "config 2 hs-brightness 512x512 vs reference")


def test_config3_ba_270x480():
    from optical_flow import estimate_flow
    g = _golden_or_skip("cfg_ba270.npz")
    im1, im2, _ = quant_gray_pair(270, 480, 1, disc=True)
    _check_inputs(g, im1, im2)
    uv = estimate_flow(im1, im2, "ba")
    assert_close(uv, g["uv"].astype(np.float64), E2E_TOL, "config 3 ba 270x480 (3 x 10 iterations) vs reference")


def test_config5_classicpp_540x960_short():
    from optical_flow import estimate_flow
    g = _golden_or_skip("cfg_cpp540.npz")
    im1, im2, _ = quant_gray_pair(540, 960, 2)
    _check_inputs(g, im1, im2)
    uv = estimate_flow(im1, im2, "classic++", {"max_iters": 3, "exact_rtol": 1e-11})
    assert_close(uv, g["uv"].astype(np.float64), E2E_TOL, "config 5 classic++ max_iters=3 540x960 vs reference")


def test_config5_classicpp_540x960_teacher_forced():
    """One warp iteration at size from a known flow, GNC alpha = 1 (quadratic) and alpha = 0 (generalized Charbonnier):
    ROF texture, Hermite partial_deriv, flow_operator, solve -- against the reference's spsolve."""
    from optical_flow import load_of_method
    from optical_flow.utils.derivatives import partial_deriv
    from optical_flow.utils.image_processing import structure_texture_decomposition_rof
    g = _golden_or_skip("cfg_cpp540_tf.npz")
    im1, im2, flow = quant_gray_pair(540, 960, 2)
    _check_inputs(g, im1, im2)
    uv_in = tf_flow(540, 960, flow)
    _check_inputs(g, uv_in, key="uv_in_crc")
    ope = load_of_method("classic++")
    ope.exact_rtol = 1e-11
    tex = structure_texture_decomposition_rof(np.stack([im1, im2], axis=2), 1.0 / 8, 100, ope.alp)
    assert abs(tex.sum() - float(g["tex_sum"])) <= 1e-9 * float(g["tex_abs_sum"])
    It, Ix, Iy = partial_deriv(tex, uv_in, ope.interpolation_method, ope.deriv_filter, ope.blend)
    for name, arr in (("It", It), ("Ix", Ix), ("Iy", Iy)):
        assert_close(arr[::8, ::8], g[name + "_s8"], 1e-9, "config 5 540x960 %s (stride-8 grid)" % name)
    ope.images = tex
    qua = ope._qua()
    zero = np.zeros_like(uv_in)
    for alpha, A in ((1.0, qua.flow_operator(uv_in, zero, It, Ix, Iy)[0]), (0.0, ope.flow_operator(uv_in, zero, It, Ix, Iy)[0])):
        assert_close(np.asarray(A.b).reshape(uv_in.shape, order="F")[::8, ::8], g["b_a%g_s8" % alpha], 1e-9,
                     "config 5 540x960 rhs alpha=%g" % alpha)
        x = ope._solve_linear_system(A, A.b, uv_in.shape)
        assert_close(x, g["x_a%g" % alpha].astype(np.float64), 1e-6, "config 5 540x960 teacher-forced increment alpha=%g" % alpha)


@pytest.mark.parametrize("tag,h,w,seed,preset,disc,stride", [
    ("ba1080_stage", 1080, 1920, 1, "ba", True, 8),
    ("cpp4k_stage", 2160, 3840, 2, "classic++", False, 16),
])
def test_operator_level_at_full_size(tag, h, w, seed, preset, disc, stride):
    """partial_deriv and flow_operator at the FULL size of configs 3 and 5 against the reference (stride grid, the last
    rows / columns in full, and whole-array sums): the B-spline prefilter at line length 1920, 64-bit indexing at 4K."""
    from optical_flow import load_of_method
    from optical_flow.utils.derivatives import partial_deriv
    g = _golden_or_skip("cfg_%s.npz" % tag)
    im1, im2, flow = quant_gray_pair(h, w, seed, disc=disc)
    _check_inputs(g, im1, im2)
    uv_in = tf_flow(h, w, flow)
    _check_inputs(g, uv_in, key="uv_in_crc")
    probe = np.random.default_rng(1000 + seed).standard_normal((h, w, 2))
    _check_inputs(g, probe, key="probe_crc")
    images = np.stack([im1, im2], axis=2)
    ope = load_of_method(preset)
    It, Ix, Iy = partial_deriv(images, uv_in, ope.interpolation_method, ope.deriv_filter, ope.blend)
    ope.images = images
    A = ope.flow_operator(uv_in, np.zeros_like(uv_in), It, Ix, Iy)[0]
    Ap = A.matvec(probe.reshape(-1, order="F")).reshape(uv_in.shape, order="F")
    b = np.asarray(A.b).reshape(uv_in.shape, order="F")
    s = (slice(None, None, stride), slice(None, None, stride))
    for name, arr, tol in (("It", It, 1e-9), ("Ix", Ix, 1e-9), ("Iy", Iy, 1e-9), ("Ap", Ap, None), ("b", b, None)):
        # A @ probe and b are sums of products of O(1e4) weights with O(1) differences: relative to the largest entry
        scale = max(1.0, float(np.abs(g[name + "_grid"]).max()), float(np.abs(g[name + "_lastrows"]).max()))
        t = tol if tol is not None else 1e-10 * scale
        assert_close(arr[s], g[name + "_grid"], t, "%s %s grid" % (tag, name))
        assert_close(arr[-3:], g[name + "_lastrows"], t, "%s %s last rows" % (tag, name))
        assert_close(arr[:, -3:], g[name + "_lastcols"], t, "%s %s last columns" % (tag, name))
        assert abs(np.abs(arr).sum() - float(g[name + "_abs_sum"])) <= 1e-9 * float(g[name + "_abs_sum"]), name


MB_SEQS = ["Dimetrodon", "Grove2", "Grove3", "Hydrangea", "RubberWhale", "Urban2", "Urban3", "Venus"]


@pytest.mark.parametrize("seq", MB_SEQS)
def test_config1_middlebury_sequences(seq):
    """classic+nl-fast on every Middlebury training sequence that has ground truth: final flow within 1e-3 px of the
    reference's, AAE / AEPE against the .flo ground truth within 0.5 % of the reference's."""
    from optical_flow import estimate_flow, flow_angular_error
    g = _golden_or_skip("cfg_mb_%s.npz" % seq)
    uv = estimate_flow(g["im1"].astype(float), g["im2"].astype(float), "classic+nl-fast")
    assert_close(uv, g["uv"].astype(np.float64), E2E_TOL, "Middlebury %s classic+nl-fast final flow" % seq)
    aae, std, aepe = flow_angular_error(g["tu"].astype(float), g["tv"].astype(float), uv[:, :, 0], uv[:, :, 1], 0)
    assert abs(aae - float(g["aae"])) <= 0.005 * float(g["aae"]), (aae, float(g["aae"]))
    assert abs(aepe - float(g["aepe"])) <= 0.005 * float(g["aepe"]), (aepe, float(g["aepe"]))
