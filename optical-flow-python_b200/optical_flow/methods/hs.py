"""Horn-Schunck driver (reference: methods/hs.py).  compute_flow runs the whole coarse-to-fine loop --
pre-processing, pyramid, <= max_warping_iters warps per level, 5x5 medians, early exit on ||x|| < 1e-3 --
inside one b200flow_estimate call."""
import numpy as np

from optical_flow import _lib
from optical_flow.methods.base import BaseOpticalFlow, FlowOperator, METHOD_CODES
from optical_flow.robust.robust_function import RobustFunction
from optical_flow.utils.derivatives import partial_deriv


class HSOpticalFlow(BaseOpticalFlow):
    """Quadratic data term + Laplacian (unit-weight five-point) smoothness."""

    _method_code = METHOD_CODES['hs']

    def __init__(self):
        super().__init__()
        self.lambda_ = 80
        self.lambda_q = 80
        self.gnc_iters = 1
        self.pyramid_levels = 4
        self.pyramid_spacing = 2.0
        self.max_warping_iters = 10
        self.interpolation_method = 'cubic'
        self.sigmaD2 = 1.0
        self.sigmaS2 = 1.0
        self.mf_iter = 1
        method = 'quadratic'
        self.rho_spatial_u = [RobustFunction(method, 1), RobustFunction(method, 1)]
        self.rho_spatial_v = [RobustFunction(method, 1), RobustFunction(method, 1)]
        self.rho_data = RobustFunction(method, 1)

    def compute_flow(self, init=None, gt=None):
        """(H, W, 2) flow.  As in the reference (hs.py:73) the pyramid depth is always the automatic one."""
        images = _lib.f64(self.images)
        self.pyramid_levels = self._auto_pyramid_levels(images)
        P = self._c_params(levels=self.pyramid_levels)
        if self.pyramid_levels < 1:
            P.pyramid_levels, P.auto_level = 0, 1     # min(H, W) < 16: no level runs, flow = init (+ final median)
        self._apply_solver(P)
        return self._run(P, images, None, init, log_style='hs')

    def _copy_with_images(self, images):
        small = self._level_copy()
        small.images = images
        small.pyramid_levels = 1
        return small

    def compute_flow_base(self, uv):
        """Warps at a single level on self.images as they are (hs.py:109-142): no pre-processing, no final median."""
        P = self._c_params(levels=1)
        P.texture = -1
        P.final_median = 0
        self._apply_solver(P)
        return self._run(P, self.images, None, uv, log_style='hs_base')

    def flow_operator(self, uv, duv=None, It=None, Ix=None, Iy=None):
        """(A, b, None, True) with A matrix-free (hs.py:144-203); derivatives are recomputed from uv as in the reference."""
        uv = _lib.f64(uv)
        It, Ix, Iy = partial_deriv(self.images, uv, self.interpolation_method, self.deriv_filter)
        A = FlowOperator(self, uv, _lib.f64(np.zeros_like(uv)), It, Ix, Iy, [(1.0, self)])
        return A, A.b, None, True
