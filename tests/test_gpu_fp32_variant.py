"""The fp32 VARIANT north_star asks for (solver_precision='fp32': the IC-preconditioned solver entirely in fp32, stopped on its
own fp32 residual at 1e-6, no fp64 residual replacement; the rest of the pipeline stays fp64).  It is NOT parity-grade -- a
1e-6 solve leaves ~1e-4 px, which the discrete weighted-median selections amplify at a few pixels -- so it is judged by the
statistical protocol of SURVEY section 7: (i) teacher-forced, one solve from a given flow: within 1e-3 px of the fp64-grade
solve; (ii) end to end: AEPE against the KNOWN synthetic flow within 0.5 % of the default path's, and the |d uv| histogram
against the default path concentrated near zero (median <= 1e-4 px, 99 % <= 2e-2 px)."""
import numpy as np
import pytest

from conftest import ROOT
import synth

pytestmark = pytest.mark.gpu


def test_fp32_variant_teacher_forced_solve(stages):
    from optical_flow import load_of_method
    from optical_flow.utils.derivatives import partial_deriv
    import flow_oracle as fo
    im1, im2 = stages["rgb1"].astype(float), stages["rgb2"].astype(float)
    images = np.stack([fo.rgb2gray(im1), fo.rgb2gray(im2)], axis=2)
    uv = np.zeros(images.shape[:2] + (2,))
    outs = {}
    for prec in ("mixed", "fp32"):
        ope = load_of_method("classic+nl-fast")
        ope.solver_precision = prec
        ope.images = images
        It, Ix, Iy = partial_deriv(images, uv, ope.interpolation_method, ope.deriv_filter, ope.blend)
        A = ope.flow_operator(uv, np.zeros_like(uv), It, Ix, Iy)[0]
        outs[prec] = (ope._solve_linear_system(A, A.b, uv.shape), ope.last_stats)
    x64, st64 = outs["mixed"]
    x32, st32 = outs["fp32"]
    assert st32["pcg_iters"] < st64["pcg_iters"], (st32, st64)
    assert 1e-9 < st32["relres"] < 1e-4, st32                 # the reported fp64 TRUE residual of the fp32 solve
    assert np.abs(x32 - x64).max() <= 1e-3, np.abs(x32 - x64).max()


def test_fp32_variant_statistical_parity():
    import sys
    sys.path.insert(0, ROOT)
    import bench
    from optical_flow import estimate_flow_batch
    pairs = [bench.synth_pair(240, 320, 30 + k) for k in range(4)]
    ims1 = np.stack([p[0] for p in pairs])
    ims2 = np.stack([p[1] for p in pairs])
    flow = np.stack([p[2] for p in pairs])
    uv64 = estimate_flow_batch(ims1, ims2, "classic+nl-fast").copy()
    uv32, st = estimate_flow_batch(ims1, ims2, "classic+nl-fast", {"solver_precision": "fp32"}, return_stats=True)
    assert st["not_converged"] == 0
    epe = lambda uv: float(np.sqrt(((uv - flow) ** 2).sum(-1))[:, 8:-8, 8:-8].mean())  # noqa: E731
    e64, e32 = epe(uv64), epe(uv32)
    assert abs(e32 - e64) <= 0.005 * e64, (e32, e64)
    d = np.abs(uv32 - uv64).max(axis=-1).ravel()
    assert np.median(d) <= 1e-4 and np.quantile(d, 0.99) <= 2e-2, (np.median(d), np.quantile(d, 0.99), d.max())
