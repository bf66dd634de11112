"""RobustFunction: same public surface as the reference's optical_flow/robust/robust_function.py:30-145
(.method, .sigma, .param, evaluate, deriv, deriv_over_x, evaluate_log, repr); evaluation runs on the GPU."""
import numpy as np

from optical_flow.robust import penalties as _p

PENALTY_MAP = {k: getattr(_p, k) for k in _p.KINDS}


class RobustFunction:
    def __init__(self, method, *args):
        if method not in PENALTY_MAP:
            raise ValueError(f"Unknown penalty method '{method}'. Available: {list(PENALTY_MAP.keys())}")
        self.method = method
        self._func = PENALTY_MAP[method]
        if method in ("generalized_charbonnier", "tdist", "tdist_unnorm") and len(args) >= 2:
            self.sigma = np.array([args[0], args[1]], dtype=float)
        elif len(args) > 0:
            self.sigma = np.atleast_1d(np.asarray(args[0], dtype=float))
        else:
            self.sigma = np.array([1.0])

    @property
    def param(self):
        return self.sigma

    def evaluate(self, x):
        return self._func(np.asarray(x, dtype=float), self.sigma, 0)

    def deriv(self, x):
        return self._func(np.asarray(x, dtype=float), self.sigma, 1)

    def deriv_over_x(self, x):
        return self._func(np.asarray(x, dtype=float), self.sigma, 2)

    def evaluate_log(self, x):
        return self.evaluate(x)

    def c_struct(self):
        return _p.make_penalty(self.method, self.sigma)

    def __repr__(self):
        return f"RobustFunction('{self.method}', sigma={self.sigma})"
