"""The ten robust penalties rho(x), rho'(x), rho'(x)/x, evaluated on the GPU (b200flow_robust_eval).

Same call convention as the reference's optical_flow/robust/penalties.py:18-345: `fn(x, sigma, d_type)` with
d_type 0 = value, 1 = first derivative, 2 = derivative over x (the IRLS weight).  `mixture` and
`spline_penalty` raise NotImplementedError exactly like the reference (penalties.py:348-373).
"""
import numpy as np

from optical_flow import _lib

KINDS = ["quadratic", "lorentzian", "charbonnier", "generalized_charbonnier", "geman_mcclure", "huber", "tukey",
         "gaussian", "tdist", "tdist_unnorm"]
_TWO_PARAM = {"generalized_charbonnier", "tdist", "tdist_unnorm"}


def make_penalty(kind, sigma):
    """(name, parameter array) -> the C struct the kernels take."""
    p = np.atleast_1d(np.asarray(sigma, dtype=float))
    if kind in _TWO_PARAM and p.size < 2:
        raise IndexError("penalty '%s' needs two parameters" % kind)
    return _lib.Penalty(KINDS.index(kind), float(p[0]), float(p[1]) if p.size > 1 else 0.0)


def _evaluate(kind, x, sigma, d_type):
    if d_type not in (0, 1, 2):
        raise ValueError(f"Unknown d_type: {d_type}")
    x = np.asarray(x, dtype=float)
    xin = _lib.f64(x).reshape(-1)
    out = np.empty_like(xin)
    ctx = _lib.default_context()
    ctx.call("b200flow_robust_eval", make_penalty(kind, sigma), int(d_type), _lib.ptr(xin), xin.size, _lib.ptr(out))
    return out.reshape(x.shape)


def _mk(kind):
    def fn(x, sigma, d_type):
        return _evaluate(kind, x, sigma, d_type)
    fn.__name__ = kind
    fn.__doc__ = "%s penalty; see module docstring for the (x, sigma, d_type) convention." % kind
    return fn


quadratic = _mk("quadratic")
lorentzian = _mk("lorentzian")
charbonnier = _mk("charbonnier")
generalized_charbonnier = _mk("generalized_charbonnier")
geman_mcclure = _mk("geman_mcclure")
huber = _mk("huber")
tukey = _mk("tukey")
gaussian = _mk("gaussian")
tdist = _mk("tdist")
tdist_unnorm = _mk("tdist_unnorm")


def mixture(x, sigma, d_type):
    raise NotImplementedError("Mixture penalty is not yet implemented. It requires a complex parameter structure: "
                              "sigma = [[weights], [component_funcs], [component_params]].")


def spline_penalty(x, sigma, d_type):
    raise NotImplementedError("Spline penalty is not yet implemented.")
