"""Preset factory: method name -> configured driver (the preset VALUES are part of the drop-in contract;
reference: methods/config.py:10-176)."""
import numpy as np

from optical_flow.robust.robust_function import RobustFunction

_MEDIAN = [5, 5]


def _set_penalties(ope, kind, spatial, data):
    ope.spatial_filters = [np.array([[1, -1]]), np.array([[1], [-1]])]
    ope.rho_spatial_u = [RobustFunction(kind, *spatial), RobustFunction(kind, *spatial)]
    ope.rho_spatial_v = [RobustFunction(kind, *spatial), RobustFunction(kind, *spatial)]
    ope.rho_data = RobustFunction(kind, *data)


def _hs(lam, texture=False, display=False):
    from optical_flow.methods.hs import HSOpticalFlow
    ope = HSOpticalFlow()
    ope.median_filter_size = list(_MEDIAN)
    ope.texture = texture
    ope.lambda_ = ope.lambda_q = lam
    ope.display = display
    return ope


def _ba(kind, spatial, data, lam, texture, interp='cubic'):
    from optical_flow.methods.ba import BAOpticalFlow
    ope = BAOpticalFlow()
    ope.median_filter_size = list(_MEDIAN)
    ope.texture = texture
    ope.interpolation_method = interp
    _set_penalties(ope, kind, spatial, data)
    ope.lambda_ = ope.lambda_q = lam
    return ope


def _classic_nl(max_iters=10, gnc_iters=3, display=False, full=False):
    from optical_flow.methods.classic_nl import ClassicNLOpticalFlow
    ope = ClassicNLOpticalFlow()
    ope.texture = True
    ope.median_filter_size = list(_MEDIAN)
    ope.alp = 0.95
    ope.area_hsz = 7
    ope.sigma_i = 7
    ope.color_images = np.ones((1, 1, 3))       # placeholder: "use colour"; estimate_flow swaps in the Lab image
    ope.lambda_ = ope.lambda_q = 3
    ope.max_iters, ope.gnc_iters, ope.display, ope.fullVersion = max_iters, gnc_iters, display, full
    return ope


_PRESETS = {
    'classic+nl-fast': lambda: _classic_nl(max_iters=3, gnc_iters=2, display=True),
    'classic+nl': lambda: _classic_nl(),
    'classic+nl-full': lambda: _classic_nl(full=True),      # fullVersion is accepted and ignored, as upstream
    'hs-brightness': lambda: _hs(10),
    'hs': lambda: _hs(40, texture=True, display=True),
    'ba-brightness': lambda: _ba('lorentzian', (0.1,), (3.5,), 0.045, False),
    'ba': lambda: _ba('lorentzian', (0.03,), (1.5,), 0.06, True),
    'classic-l': lambda: _ba('lorentzian', (0.03,), (1.5,), 0.06, True),
    'classic-c-brightness': lambda: _ba('charbonnier', (1e-3,), (1e-3,), 3, False),
    'classic-c': lambda: _ba('charbonnier', (1e-3,), (1e-3,), 5, True),
    'classic++': lambda: _ba('generalized_charbonnier', (1e-3, 0.45), (1e-3, 0.45), 3, True, interp='bi-cubic'),
}


def load_of_method(method):
    """Configured optical-flow object for one of: classic+nl-fast, classic+nl, classic+nl-full, hs-brightness, hs,
    ba-brightness, ba, classic-l, classic-c-brightness, classic-c, classic++.  ('classic-c-a' / Alt-BA is outside
    the hot path this package accelerates: its upstream implementation diverges numerically)"""
    if method == 'classic-c-a':
        raise NotImplementedError("'classic-c-a' (Alt-BA) is not part of the B200 hot path (no usable oracle upstream)")
    if method not in _PRESETS:
        raise ValueError(f"Unknown optical flow method: '{method}'")
    return _PRESETS[method]()
