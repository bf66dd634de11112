"""Black-Anandan driver: robust penalties, graduated non-convexity, 5x5 median per warp (reference: methods/ba.py)."""
import time

from optical_flow import _lib
from optical_flow.methods.base import BaseOpticalFlow, METHOD_CODES
from optical_flow.robust.robust_function import RobustFunction


class BAOpticalFlow(BaseOpticalFlow):
    _method_code = METHOD_CODES['ba']

    def __init__(self):
        super().__init__()
        self.gnc_iters = 3
        self.alpha = 1.0
        self.max_iters = 10
        self.max_linear = 1
        self.interpolation_method = 'cubic'
        method = 'lorentzian'
        self.rho_spatial_u = [RobustFunction(method, 0.03), RobustFunction(method, 0.03)]
        self.rho_spatial_v = [RobustFunction(method, 0.03), RobustFunction(method, 0.03)]
        self.rho_data = RobustFunction(method, 1.5)

    def _qua(self):
        """Quadratic relaxation of compute_flow_base (ba.py:146-160): unit quadratic spatial terms, data term
        quadratic with sigma_data / sigma_spatial, lambda = lambda_q."""
        qua = self._level_copy()
        qua.lambda_ = self.lambda_q
        ta = self.rho_data.param[0] / self.rho_spatial_u[0].param[0]
        qua.rho_spatial_u = [RobustFunction('quadratic', 1) for _ in self.rho_spatial_u]
        qua.rho_spatial_v = [RobustFunction('quadratic', 1) for _ in self.rho_spatial_v]
        qua.rho_data = RobustFunction('quadratic', ta)
        return qua

    def _levels(self, images):
        if self.auto_level:
            self.pyramid_levels = self._auto_pyramid_levels(images)
        return self.pyramid_levels

    def compute_flow(self, init=None, gt=None):
        """GNC stages x pyramid levels x max_iters warps in one device call (ba.py:57-138).  self.alpha is left
        untouched, as the reference restores it."""
        self._check_fc()
        images = _lib.f64(self.images)
        t0 = self._display_t0 = time.time()
        self._stage_lines = 0
        P = self._c_params(levels=self._levels(images))
        if self.pyramid_levels < 1:
            P.pyramid_levels, P.auto_level = 0, 1
        self._apply_solver(P)
        uv = self._run(P, images, None, init, log_style='gnc')
        # ba.py:132-133: one "finished" line per stage, display or not (with display on, the earlier stages' lines were
        # already printed between the log lines)
        for k in range(self._stage_lines, int(self.gnc_iters) - 1):
            self._print_stage_done(k)
        if int(self.gnc_iters) >= 1:
            print(f"GNC stage {int(self.gnc_iters)} finished, {(time.time() - t0) / 60:.2f} minutes passed")
        return uv

    def compute_flow_base(self, uv):
        """max_iters IRLS warps on self.images at the current self.alpha (ba.py:140-206)."""
        P = self._c_params(levels=1)
        P.texture = -1
        P.gnc_iters = 1
        self._apply_solver(P)
        return self._run(P, self.images, None, uv, log_style='gnc_base')

    def flow_operator(self, uv, duv, It, Ix, Iy):
        """(A, b, None, iterative) for this object's own penalties and lambda_ (ba.py:208-302); A is matrix-free."""
        return self._gnc_flow_operator(uv, duv, It, Ix, Iy)
