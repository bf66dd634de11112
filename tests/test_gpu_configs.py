"""BASELINE.json configs 2-5 on the GPU.

(1) Seeded synthetic inputs at sizes the CPU oracle finishes in seconds (ragged, non-multiple-of-tile sizes): final
    flow of the CUDA path against the oracle, 1e-3 px (north_star's fp64 tolerance).
(2) The configs at their FULL sizes (1024^2 hs-brightness, 1920x1080 ba, Middlebury-size classic+nl-full batch,
    3840x2160 classic++) through size-independent properties: every solve converged to the fp64 true-residual target,
    the flow is finite and reproduces the known synthetic motion (AEPE), batch items equal single runs bit for bit,
    repeated runs are bit-identical, and classic+nl-full == classic+nl (the reference ignores fullVersion)."""
import numpy as np
import pytest

from conftest import assert_close
import synth

pytestmark = pytest.mark.gpu


def _estimate_with_stats(im1, im2, preset, params=None):
    from optical_flow import load_of_method
    from optical_flow import interface
    ope_holder = {}
    orig = interface.load_of_method

    def spy(name):
        ope_holder["ope"] = orig(name)
        return ope_holder["ope"]
    interface.load_of_method = spy
    try:
        uv = interface.estimate_flow(im1, im2, preset, params)
    finally:
        interface.load_of_method = orig
    return uv, ope_holder["ope"].last_stats


# ---------------------------------------------------------------------------------------------------------------
# (1) oracle comparison on seeded synthetic inputs, ragged sizes
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("preset,h,w,params", [
    ("hs-brightness", 67, 91, None),
    ("hs", 50, 77, None),
    ("ba", 45, 67, {"max_iters": 3}),
    ("ba-brightness", 61, 53, {"max_iters": 2}),
    ("classic+nl-fast", 59, 83, None),
])
def test_seeded_vs_oracle(preset, h, w, params):
    import flow_oracle as fo
    from optical_flow import estimate_flow
    im1, im2, _ = synth.gray_pair(h, w, seed=h + w, disc=preset.startswith("ba"))
    im1, im2 = np.floor(im1 + 0.5), np.floor(im2 + 0.5)          # 8-bit-valued like real frames
    want = fo.estimate_flow(im1, im2, preset, params)
    got = estimate_flow(im1, im2, preset, params)
    assert_close(got, want, 1e-3, "seeded %s %dx%d vs oracle" % (preset, w, h))


def test_seeded_color_vs_oracle():
    """RGB input (Lab-guided weighted median) on an odd-sized synthetic colour pair."""
    import bench
    import flow_oracle as fo
    from optical_flow import estimate_flow
    a, b, _ = bench.synth_pair(57, 75, 5)
    want = fo.estimate_flow(a.astype(float), b.astype(float), "classic+nl-fast")
    got = estimate_flow(a.astype(float), b.astype(float), "classic+nl-fast")
    assert_close(got, want, 1e-3, "seeded classic+nl-fast RGB 75x57 vs oracle")


# ---------------------------------------------------------------------------------------------------------------
# (2) full-size configs through properties
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.slow
def test_config2_hs_brightness_1024():
    im1, im2, flow = synth.gray_pair(1024, 1024, seed=0)
    uv, st = _estimate_with_stats(im1, im2, "hs-brightness")
    assert uv.shape == (1024, 1024, 2) and np.isfinite(uv).all()
    assert st["not_converged"] == 0 and st["solves"] > 0
    epe = synth.interior_epe(uv, flow)
    assert epe < 0.05, "AEPE vs the known affine flow %.4f px" % epe
    uv2, _ = _estimate_with_stats(im1, im2, "hs-brightness")
    np.testing.assert_array_equal(uv, uv2)                          # run-to-run deterministic


@pytest.mark.slow
def test_config3_ba_1080p():
    im1, im2, flow = synth.gray_pair(1080, 1920, seed=1, disc=True)
    uv, st = _estimate_with_stats(im1, im2, "ba", {"max_iters": 3})
    assert uv.shape == (1080, 1920, 2) and np.isfinite(uv).all()
    assert st["not_converged"] == 0
    yy, xx = np.mgrid[0:1080, 0:1920]
    far = (xx - 959.5) ** 2 + (yy - 539.5) ** 2 > (1080 / 6.0 + 40) ** 2
    far[:8] = far[-8:] = False
    far[:, :8] = far[:, -8:] = False
    epe_bg = float(np.sqrt(((uv - flow) ** 2).sum(-1))[far].mean())
    assert epe_bg < 0.1, "background AEPE %.4f px" % epe_bg
    inside = (xx - 959.5) ** 2 + (yy - 539.5) ** 2 < (1080 / 6.0 - 20) ** 2
    epe_fg = float(np.sqrt(((uv - flow) ** 2).sum(-1))[inside].mean())
    assert epe_fg < 0.5, "foreground-disc AEPE %.4f px" % epe_fg


@pytest.mark.slow
def test_config4_classic_nl_full_batch():
    """Middlebury-size RGB pairs, 'classic+nl-full' (== 'classic+nl': fullVersion is ignored, weighted_median.py:24,62)."""
    import bench
    from optical_flow import estimate_flow, estimate_flow_batch
    pairs = [bench.synth_pair(388, 584, 20 + k) for k in range(3)]
    ims1 = np.stack([p[0] for p in pairs])
    ims2 = np.stack([p[1] for p in pairs])
    uv, st = estimate_flow_batch(ims1, ims2, "classic+nl-full", params={"max_iters": 3}, return_stats=True)
    assert st["not_converged"] == 0 and np.isfinite(uv).all()
    uv_nl = estimate_flow_batch(ims1, ims2, "classic+nl", params={"max_iters": 3})
    np.testing.assert_array_equal(uv, uv_nl)
    single = estimate_flow(ims1[1].astype(float), ims2[1].astype(float), "classic+nl-full", {"max_iters": 3})
    assert_close(uv[1], single, 1e-6, "config 4: batch item vs single run")
    for k in range(3):
        epe = synth.interior_epe(uv[k], pairs[k][2])
        assert epe < 0.08, "pair %d AEPE %.4f px" % (k, epe)


@pytest.mark.slow
def test_config5_classicpp_4k():
    im1, im2, flow = synth.gray_pair(2160, 3840, seed=2)
    uv, st = _estimate_with_stats(im1, im2, "classic++", {"max_iters": 2})
    assert uv.shape == (2160, 3840, 2) and np.isfinite(uv).all()
    assert st["not_converged"] == 0
    epe = synth.interior_epe(uv, flow, margin=16)
    assert epe < 0.1, "AEPE vs the known affine flow %.4f px" % epe
