"""Shared state and plumbing of the three method drivers (reference: methods/base.py).

The public attributes, their defaults and `parse_input_parameter` are those of the reference's BaseOpticalFlow
(base.py:21-85); everything numerical is delegated to libb200flow.so.  Additions (not in the reference):
  exact_rtol / exact_maxiter  relative-residual target and iteration cap of the matrix-free block-Jacobi PCG that
                              stands in for the direct `solver='backslash'` solve.  Measured on RubberWhale 584x388
                              classic+nl-fast against the reference's SuperLU pipeline: 1e-8 -> 3.5e-3 px (a median
                              selection flips), 1e-9 -> 3.6e-5 px.  Round 2, all eight Middlebury sequences: 1e-10 is not
                              enough on five of them (up to 1.7e-2 px at a few pixels), the default is 1e-12 (<= 6e-5 px)
  solver_precision            'mixed' (default): the PCG keeps its Krylov vectors in fp32 and accumulates the solution and
                              the periodically recomputed TRUE residual in fp64 ("reliable updates"), so convergence
                              is still declared on the fp64 residual ||b - A x|| <= exact_rtol ||b||; 'fp64': every
                              vector in fp64 (same criterion, 1.9x the memory traffic).  'mixed' is preconditioned by a
                              tile-local 2x2-block incomplete Cholesky IC(0) (~2.2x fewer iterations than block Jacobi);
                              'mixed-jacobi' keeps the round-1 block-Jacobi preconditioner (reported variant);
                              'fp32': the fp32 VARIANT north_star asks for -- the IC solver entirely in fp32, stopped when
                              its own fp32 residual reaches fp32_rtol (1e-6), no fp64 residual replacement; everything
                              else stays fp64.  Not parity-grade (judged statistically, tests/test_gpu_fp32_variant.py)
  last_stats                  dict of solver / launch statistics of the most recent compute_flow call
"""
import copy
from abc import ABC, abstractmethod

import numpy as np

from optical_flow import _lib

# solver_precision -> B200FLOW_SOLVER_* of the 'backslash' stand-in (include/b200flow.h)
_EXACT_SOLVERS = {'mixed': 4, 'mixed-jacobi': 0, 'fp64': 2, 'fp32': 5}
from optical_flow.robust.robust_function import RobustFunction
from optical_flow.utils.derivatives import INTERP_CODES
from optical_flow.utils.image_processing import fspecial_gaussian
from optical_flow.utils.pyramid import compute_image_pyramid

METHOD_CODES = {'hs': 0, 'ba': 1, 'classic_nl': 2}


class FlowOperator:
    """Matrix-free stand-in for the sparse `A` the reference's flow_operator returns: supports `A @ x`,
    `.matvec`, `.diagonal()`, `.shape`, and the GNC blend `alpha * A_q + (1 - alpha) * A_r` of two operators built
    from the same (uv, duv, It, Ix, Iy).  Vectors use the reference's column-major [u(:); v(:)] ordering."""

    def __init__(self, owner, uv, duv, It, Ix, Iy, terms):
        self._owner, self._uv, self._duv, self._It, self._Ix, self._Iy = owner, uv, duv, It, Ix, Iy
        self.terms = terms                      # list of (coefficient, driver object carrying rho_* / lambda_)
        H, W = uv.shape[:2]
        self.hw = (H, W)
        self.nc = 1 if np.ndim(It) == 2 else It.shape[2]     # channels per frame (derivative planes are (H, W, nc))
        self.shape = (2 * H * W, 2 * H * W)
        self.dtype = np.dtype(float)
        self._b = None

    # -- linear combinations ------------------------------------------------------------------------
    def __rmul__(self, c):
        return FlowOperator(self._owner, self._uv, self._duv, self._It, self._Ix, self._Iy,
                            [(float(c) * k, o) for k, o in self.terms])

    __mul__ = __rmul__

    def __add__(self, other):
        if not isinstance(other, FlowOperator) or other._uv is not self._uv or other._It is not self._It:
            return NotImplemented
        return FlowOperator(self._owner, self._uv, self._duv, self._It, self._Ix, self._Iy, self.terms + other.terms)

    # -- device calls ---------------------------------------------------------------------------------
    def _each(self, x):
        H, W = self.hw
        ctx = _lib.default_context()
        tot_ax = tot_b = tot_d = 0.0
        for coef, obj in self.terms:
            P = obj._c_params(single=True)
            ax, b, d = np.empty((H, W, 2)), np.empty((H, W, 2)), np.empty((H, W, 2))
            ctx.call("b200flow_operator_apply_mc", P, 0.0, _lib.ptr(self._uv), _lib.ptr(self._duv), _lib.ptr(self._It),
                     _lib.ptr(self._Ix), _lib.ptr(self._Iy), H, W, self.nc, _lib.ptr(x),
                     _lib.ptr(ax) if x is not None else None, _lib.ptr(b), _lib.ptr(d))
            if x is not None:
                tot_ax = tot_ax + coef * ax
            tot_b = tot_b + coef * b
            tot_d = tot_d + coef * d
        return tot_ax, tot_b, tot_d

    def matvec(self, x):
        H, W = self.hw
        xi = _lib.f64(np.asarray(x, dtype=float).reshape((H, W, 2), order='F'))
        return self._each(xi)[0].reshape(-1, order='F')

    def __matmul__(self, x):
        return self.matvec(x)

    dot = matvec

    def diagonal(self):
        return self._each(None)[2].reshape(-1, order='F')

    @property
    def b(self):
        if self._b is None:
            self._b = self._each(None)[1].reshape(-1, order='F')
        return self._b

    def tocsc(self):
        """The operator as a scipy.sparse CSC matrix (what the reference's flow_operator returns), recovered from the
        matrix-free device operator by probing: the 2N x 2N matrix couples (u, v) of a pixel with itself and its four
        neighbours, so 3 x 3 pixel colourings x 2 components = 18 products A @ e give every entry exactly."""
        from scipy import sparse
        H, W = self.hw
        N = H * W
        yy, xx = np.mgrid[0:H, 0:W]
        colour = (yy % 3) * 3 + (xx % 3)
        rows, cols, vals = [], [], []
        pix = np.arange(N).reshape(H, W, order='F')              # column-major pixel index of the reference's vectorisation
        for comp in range(2):
            for c in range(9):
                e = np.zeros((H, W, 2))
                e[:, :, comp][colour == c] = 1.0
                y = self.matvec(e.reshape(-1, order='F')).reshape((H, W, 2), order='F')
                # a response at pixel q (component k) comes from the unique probe pixel of colour c within distance 1 of q
                for k in range(2):
                    for dy, dx in ((0, 0), (0, 1), (0, -1), (1, 0), (-1, 0)):
                        if (dy or dx) and k != comp:
                            continue                                 # the two components couple only inside a pixel
                        src = np.zeros((H, W), dtype=bool)
                        src[colour == c] = True
                        qy, qx = yy + dy, xx + dx                    # q = p + d for every probe pixel p
                        ok = src & (qy >= 0) & (qy < H) & (qx >= 0) & (qx < W)
                        v = y[qy[ok], qx[ok], k]
                        nz = v != 0.0
                        rows.append((k * N + pix[qy[ok], qx[ok]])[nz])
                        cols.append((comp * N + pix[yy[ok], xx[ok]])[nz])
                        vals.append(v[nz])
        return sparse.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=self.shape)

    tocsr = lambda self: self.tocsc().tocsr()      # noqa: E731

    def solve(self, owner, rhs=None):
        """A x = b on the device (b = the assembled right-hand side, or the caller's rhs); x as (H, W, 2)."""
        H, W = self.hw
        if len(self.terms) == 1 and abs(self.terms[0][0] - 1.0) < 1e-15:
            P, alpha = self.terms[0][1]._c_params(single=True), 0.0
        elif len(self.terms) == 2 and abs(self.terms[0][0] + self.terms[1][0] - 1.0) < 1e-12:
            (a, q), (_, r) = self.terms
            P, alpha = r._c_params(qua=q), a
        else:
            raise NotImplementedError("only A, or alpha*A_q + (1-alpha)*A_r, can be solved on the device")
        owner._apply_solver(P)
        x = np.empty((H, W, 2))
        iters = _lib.C.c_int(0)
        rel = _lib.C.c_double(0.0)
        if rhs is None:
            _lib.default_context().call("b200flow_solve_increment_mc", P, float(alpha), _lib.ptr(self._uv),
                                        _lib.ptr(self._duv), _lib.ptr(self._It), _lib.ptr(self._Ix), _lib.ptr(self._Iy), H, W,
                                        self.nc, _lib.ptr(x), _lib.C.byref(iters), _lib.C.byref(rel), allow_noconv=True)
        else:
            r = _lib.f64(np.asarray(rhs, dtype=float).reshape((H, W, 2), order='F'))
            _lib.default_context().call("b200flow_solve_rhs_mc", P, float(alpha), _lib.ptr(self._uv),
                                        _lib.ptr(self._duv), _lib.ptr(self._It), _lib.ptr(self._Ix), _lib.ptr(self._Iy), H, W,
                                        self.nc, _lib.ptr(r), _lib.ptr(x), _lib.C.byref(iters), _lib.C.byref(rel),
                                        allow_noconv=True)
        owner.last_stats = {"pcg_iters": iters.value, "relres": rel.value}
        return x


class BaseOpticalFlow(ABC):
    """Base class for variational optical flow estimation (attribute-compatible with the reference)."""

    _method_code = None

    def __init__(self):
        self.images = None
        self.lambda_ = 1.0
        self.lambda_q = 1.0
        self.solver = 'backslash'
        self.pcg_rtol = 1e-3
        self.pcg_maxiter = 200
        self.sor_max_iters = 10000
        self.interpolation_method = 'cubic'
        self.deriv_filter = np.array([1, -8, 0, 8, -1]) / 12.0
        self.blend = 0.5
        self.texture = False
        self.fc = False
        self.median_filter_size = None
        self.limit_update = True
        self.display = False
        self.color_images = None
        self.auto_level = True
        self.alp = 0.95
        self.pyramid_levels = 4
        self.pyramid_spacing = 2.0
        self.gnc_iters = 1
        self.gnc_pyramid_levels = 2
        self.gnc_pyramid_spacing = 1.25
        self.alpha = 1.0
        self.max_iters = 10
        self.max_linear = 1
        self.spatial_filters = [np.array([[1, -1]]), np.array([[1], [-1]])]
        method = 'quadratic'
        self.rho_spatial_u = [RobustFunction(method, 1), RobustFunction(method, 1)]
        self.rho_spatial_v = [RobustFunction(method, 1), RobustFunction(method, 1)]
        self.rho_data = RobustFunction(method, 1)
        self._cached_conv_mats = {}
        # additions
        # 1e-12: measured on the eight Middlebury sequences with ground truth (tests/test_gpu_config_goldens.py) -- at 1e-10
        # five of them end up with weighted-median selections that differ from the reference's (up to 1.7e-2 px at a
        # few pixels: the systems are ill-conditioned, a 1e-10 residual still leaves ~1e-6 px of error); at 1e-12 every
        # sequence is within 6e-5 px of the reference, at 1e-13 within 8e-6 px
        self.exact_rtol = 1e-12
        self.exact_maxiter = 20000
        self.solver_precision = 'mixed'
        self.fp32_rtol = 1e-6
        self.last_stats = None

    def parse_input_parameter(self, params):
        """dict or [key, value, key, value, ...]; 'lambda' aliases lambda_; unknown keys are ignored (base.py:65-85)."""
        if isinstance(params, dict):
            items = list(params.items())
        elif isinstance(params, (list, tuple)):
            items = [(params[i], params[i + 1]) for i in range(0, len(params) - 1, 2)]
        else:
            return
        for key, val in items:
            attr = 'lambda_' if key == 'lambda' else key
            if hasattr(self, attr):
                setattr(self, attr, val)

    # ---- parameter marshalling ----------------------------------------------------------------------
    def _qua(self):
        """The quadratic GNC stand-in of this object (overridden per driver)."""
        return self

    def _apply_solver(self, P):
        solver = str(self.solver).lower()
        if solver == 'backslash':
            prec = str(self.solver_precision).lower()
            if prec not in _EXACT_SOLVERS:
                raise ValueError(f"Unknown solver_precision: {self.solver_precision}")
            P.solver, P.tol, P.maxit = _EXACT_SOLVERS[prec], float(self.exact_rtol), int(self.exact_maxiter)
            if prec == 'fp32':          # the fp32 variant stops on its own (iterated) residual: fp32 cannot deliver 1e-12
                P.tol = float(self.fp32_rtol)
        elif solver == 'pcg':
            P.solver, P.tol, P.maxit = 1, float(self.pcg_rtol), int(self.pcg_maxiter)
        elif solver == 'sor':      # base.py:109-110: _sor_solve(A, b, 1.9, self.sor_max_iters, 1e-2)
            P.solver, P.tol, P.maxit = 3, 1e-2, int(self.sor_max_iters)
        else:
            raise ValueError(f"Unknown solver: {self.solver}")

    def _c_params(self, single=False, qua=None, levels=None):
        """Fill the C parameter block from the public attributes.  single=True: flow_operator semantics (this
        object's own penalties and lambda_, no GNC blend)."""
        if self.interpolation_method not in INTERP_CODES:
            raise ValueError(f"Unknown interpolation method: {self.interpolation_method}")
        P = _lib.Params()
        P.method = self._method_code
        P.interp = INTERP_CODES[self.interpolation_method]
        P.texture = 1 if self.texture else 0
        P.gnc_iters = int(self.gnc_iters)
        P.max_iters = int(self.max_iters)
        P.max_linear = int(self.max_linear)
        P.max_warping_iters = int(getattr(self, 'max_warping_iters', self.max_iters))
        P.limit_update = 1 if self.limit_update else 0
        P.pyramid_levels = int(levels if levels is not None else self.pyramid_levels)
        P.auto_level = 0
        P.gnc_pyramid_levels = int(self.gnc_pyramid_levels)
        P.pyramid_spacing = float(self.pyramid_spacing)
        P.gnc_pyramid_spacing = float(self.gnc_pyramid_spacing)
        P.lambda_ = float(self.lambda_)
        q = qua if qua is not None else (self if single else self._qua())
        P.lambda_q = float(q.lambda_) if (single or qua is not None) else float(self.lambda_q)
        P.alpha0 = float(self.alpha)
        P.alp = float(self.alp)
        P.blend = float(self.blend)
        h = np.asarray(self.deriv_filter, dtype=float).reshape(-1)
        if h.size != 5:
            raise ValueError("deriv_filter must have 5 taps")
        for i in range(5):
            P.deriv_filter[i] = h[i]
        P.sigmaD2 = float(getattr(self, 'sigmaD2', 1.0))
        P.sigmaS2 = float(getattr(self, 'sigmaS2', 1.0))
        for i in range(2):
            P.rho_su[i] = self.rho_spatial_u[i].c_struct()
            P.rho_sv[i] = self.rho_spatial_v[i].c_struct()
            P.qua_su[i] = q.rho_spatial_u[i].c_struct()
            P.qua_sv[i] = q.rho_spatial_v[i].c_struct()
        P.rho_d = self.rho_data.c_struct()
        P.qua_d = q.rho_data.c_struct()
        mfs = self.median_filter_size
        if mfs is None:
            P.median_h = P.median_w = 0
        elif hasattr(mfs, '__len__'):
            P.median_h, P.median_w = int(mfs[0]), int(mfs[1])
        else:
            P.median_h = P.median_w = int(mfs)
        P.mf_iter = int(getattr(self, 'mf_iter', 1))
        P.area_hsz = int(getattr(self, 'area_hsz', 0))
        P.sigma_i = float(getattr(self, 'sigma_i', 7.0))
        P.occ_sigma_d, P.occ_sigma_i = 0.3, 20.0
        P.rof_iters, P.rof_theta = 100, 1.0 / 8
        P.final_median = 1
        P.solver, P.tol, P.maxit = _EXACT_SOLVERS['mixed'], float(self.exact_rtol), int(self.exact_maxiter)
        return P

    def _check_fc(self):
        if self.fc and not self.texture:
            raise NotImplementedError("fc=True (Gaussian high-pass preprocessing) is not built; no preset uses it")

    def _run(self, P, images, color, init, log_style=None):
        """One b200flow_estimate call for a single pair.  log_style ('gnc', 'hs', or their single-level forms 'gnc_base',
        'hs_base') + self.display: the reference's per-level / per-iteration display lines, printed from the device log
        after the call (the whole loop is one device call)."""
        images = _lib.f64(images)
        if images.ndim != 3 or images.shape[2] < 2 or images.shape[2] % 2:
            raise ValueError("images must be (H, W, 2C): C channels of frame 1 followed by C channels of frame 2")
        H, W = images.shape[:2]
        nc = images.shape[2] // 2
        Cn = 0
        if color is not None:
            color = _lib.f64(color)
            # the library reads H*W*C values: a guidance image of another size would be read past its end.  The reference
            # resizes it with skimage (weighted_median.py:49-56); skimage is not a dependency here, so say so instead.
            if color.ndim not in (2, 3) or color.shape[:2] != (H, W):
                raise ValueError("color_images must be (%d, %d) or (%d, %d, C); got %s" % (H, W, H, W, color.shape))
            Cn = 1 if color.ndim == 2 else color.shape[2]
        init = None if init is None else _lib.f64(init)
        if init is not None and init.shape != (H, W, 2):
            raise ValueError("init must be (%d, %d, 2); got %s" % (H, W, init.shape))
        uv = np.empty((H, W, 2))
        st = _lib.Stats()
        ctx = _lib.default_context()
        show = bool(self.display) and log_style is not None
        if show:
            ctx.set_log(True)
        try:
            ctx.call("b200flow_estimate_mc", P, 1, H, W, nc, Cn, _lib.ptr(images), _lib.ptr(color), _lib.ptr(init),
                     _lib.ptr(uv), _lib.C.byref(st))
            if show:
                self._print_display_log(ctx.get_log(), log_style)
        finally:
            if show:
                ctx.set_log(False)
        self.last_stats = st.as_dict()
        return uv

    _display_t0 = None     # set by compute_flow: start of the run, for the "minutes passed" lines
    _stage_lines = 0       # "GNC stage k finished" lines printed so far in this run

    def _print_display_log(self, rows, style):
        """The lines the reference prints under display=True, in its order and format: classic_nl.py:141-152,255-256,
        186-198; ba.py:100-114,189-190,132-133; hs.py:80-81,123-124 (HS stops a level at the first ||x|| < 1e-3, hs.py:126)."""
        import time
        hs = style.startswith('hs')
        full = not style.endswith('_base')
        last_g = last_l = None
        gated = False
        for g, l, i, j, v in rows:
            g, l, i, j = int(g), int(l), int(i), int(j)
            if full and not hs and g != last_g:
                if last_g is not None:
                    self._print_stage_done(last_g)
                print(f"GNC stage: {g + 1}")
                last_l = None
            if (g, l) != (last_g, last_l):
                gated = False
                if full:
                    print(f"Pyramid level: {l + 1}" if hs else f"  Pyramid level: {l + 1}")
            last_g, last_l = g, l
            if hs:
                if gated:
                    continue
                print(f"  Iteration: {i + 1}  (norm: {v:.6f})")
                gated = v < 1e-3
            else:
                print(f"    Iter: {i + 1} {j + 1} (delta: {v:.6f})")
        if full and not hs and last_g is not None and last_g + 1 < int(self.gnc_iters):
            self._print_stage_done(last_g)     # the last stage's line is the caller's (it may carry AAE / EPE)

    def _print_stage_done(self, g):
        import time
        t0 = self._display_t0 or time.time()
        self._stage_lines = g + 1
        print(f"GNC stage {g + 1} finished, {(time.time() - t0) / 60:.2f} minutes passed")

    # ---- reference-compatible helpers ---------------------------------------------------------------
    def _solve_linear_system(self, A, b, uv_shape, x0=None):
        """A must be the FlowOperator returned by flow_operator (base.py:87-114 semantics: solution reshaped to
        uv_shape).  Unknown solver names raise ValueError like the reference."""
        solver = str(self.solver).lower()
        if solver not in ('pcg', 'backslash', 'sor'):
            raise ValueError(f"Unknown solver: {self.solver}")
        if not isinstance(A, FlowOperator):
            raise TypeError("_solve_linear_system needs the FlowOperator returned by flow_operator (no sparse matrices "
                            "are built on this path)")
        custom = b is not None and not (b is A.b or np.array_equal(np.asarray(b).reshape(-1), A.b))
        return A.solve(self, rhs=b if custom else None).reshape(uv_shape)

    def _build_pyramid(self, images, levels, spacing):
        smooth_sigma = np.sqrt(spacing) / np.sqrt(2)
        ksize = 2 * round(1.5 * smooth_sigma) + 1
        return compute_image_pyramid(images, fspecial_gaussian(int(ksize), smooth_sigma), levels, 1.0 / spacing)

    def _auto_pyramid_levels(self, images):
        min_dim = min(images.shape[0], images.shape[1])
        return 1 + int(np.floor(np.log(min_dim / 16.0) / np.log(self.pyramid_spacing)))

    def clear_conv_cache(self):
        self._cached_conv_mats = {}

    def _gnc_flow_operator(self, uv, duv, It, Ix, Iy):
        uv, It, Ix, Iy = _lib.f64(uv), _lib.f64(It), _lib.f64(Ix), _lib.f64(Iy)
        duv = _lib.f64(np.zeros_like(uv) if duv is None else duv)
        A = FlowOperator(self, uv, duv, It, Ix, Iy, [(1.0, self)])
        return A, A.b, None, True

    def _level_copy(self):
        return copy.copy(self)

    @abstractmethod
    def compute_flow(self, init=None, gt=None):
        pass

    @abstractmethod
    def compute_flow_base(self, uv):
        pass

    @abstractmethod
    def flow_operator(self, uv, duv, It, Ix, Iy):
        pass
