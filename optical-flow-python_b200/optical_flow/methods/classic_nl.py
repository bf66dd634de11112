"""Classic+NL driver: BA-style GNC/IRLS plus the non-local term (occlusion-aware, colour-weighted 15x15 median
after every solve) -- reference: methods/classic_nl.py."""
import time

import numpy as np

from optical_flow import _lib
from optical_flow.methods.base import BaseOpticalFlow, METHOD_CODES
from optical_flow.robust.robust_function import RobustFunction


class ClassicNLOpticalFlow(BaseOpticalFlow):
    _method_code = METHOD_CODES['classic_nl']

    def __init__(self):
        super().__init__()
        self.lambda2 = 0.1
        self.lambda3 = 1.0
        self.interpolation_method = 'bi-cubic'
        self.gnc_iters = 3
        self.alpha = 1.0
        self.max_iters = 10
        self.max_linear = 1
        method, a, sig = 'generalized_charbonnier', 0.45, 1e-3
        self.rho_spatial_u = [RobustFunction(method, sig, a), RobustFunction(method, sig, a)]
        self.rho_spatial_v = [RobustFunction(method, sig, a), RobustFunction(method, sig, a)]
        self.rho_data = RobustFunction(method, sig, a)
        # attributes the reference carries but never reads on this path -- kept for compatibility
        self.seg = None
        self.mfT = 15
        self.imfsz = [7, 7]
        self.filter_weight = None
        self.hybrid = False
        self.area_hsz = 10
        self.affine_hsz = 4
        self.sigma_i = 7
        self.input_seg = None
        self.input_occ = None
        self.fullVersion = False

    def _qua(self):
        """Quadratic relaxation (classic_nl.py:212-226): quadratic penalties with the robust penalties' sigma."""
        qua = self._level_copy()
        qua.lambda_ = self.lambda_q
        qua.rho_spatial_u = [RobustFunction('quadratic', r.param[0]) for r in self.rho_spatial_u]
        qua.rho_spatial_v = [RobustFunction('quadratic', r.param[0]) for r in self.rho_spatial_v]
        qua.rho_data = RobustFunction('quadratic', self.rho_data.param[0])
        return qua

    def _color_for(self, shape_hw):
        col = self.color_images
        if col is None or np.size(col) < shape_hw[0] * shape_hw[1]:
            return None            # weighted_median.py:42-47: falls back to the plain median
        return col

    def compute_flow(self, init=None, gt=None):
        """Whole GNC / pyramid / warp loop in one device call (classic_nl.py:89-198).  Like the reference, the
        GNC alpha reached at the end is kept on the object (it is NOT restored, classic_nl.py:181-184)."""
        self._check_fc()
        images = _lib.f64(self.images)
        t0 = self._display_t0 = time.time()
        self._stage_lines = 0
        if self.auto_level:
            self.pyramid_levels = self._auto_pyramid_levels(images)
        P = self._c_params(levels=self.pyramid_levels)
        if self.pyramid_levels < 1:
            P.pyramid_levels, P.auto_level = 0, 1
        self._apply_solver(P)
        uv = self._run(P, images, self._color_for(images.shape[:2]), init, log_style='gnc')
        for ignc in range(int(self.gnc_iters)):
            if self.gnc_iters > 1:
                self.alpha = max(0, min(self.alpha, 1 - (ignc + 1) / (self.gnc_iters - 1)))
        # the reference prints one such line after EVERY stage, display or not (classic_nl.py:186-198); with display on the
        # earlier stages' lines come interleaved from the device log (_print_display_log)
        for k in range(self._stage_lines, int(self.gnc_iters) - 1):
            self._print_stage_done(k)
        msg = f"GNC stage {int(self.gnc_iters)} finished, {(time.time() - t0) / 60:.2f} minutes passed"
        if gt is not None:
            from optical_flow.evaluation.metrics import flow_angular_error
            aae, stdae, aepe = flow_angular_error(gt[:, :, 0], gt[:, :, 1], uv[:, :, 0], uv[:, :, 1], 0)
            msg += f"  AAE {aae:.3f} STD {stdae:.3f} EPE {aepe:.3f}"
        if int(self.gnc_iters) >= 1:
            print(msg)
        return uv

    def compute_flow_base(self, uv):
        """max_iters warps with the non-local term on self.images / self.color_images (classic_nl.py:200-277)."""
        P = self._c_params(levels=1)
        P.texture = -1
        P.gnc_iters = 1
        self._apply_solver(P)
        return self._run(P, self.images, self._color_for(np.shape(self.images)[:2]), uv, log_style='gnc_base')

    def flow_operator(self, uv, duv, It, Ix, Iy):
        return self._gnc_flow_operator(uv, duv, It, Ix, Iy)
