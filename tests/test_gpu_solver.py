"""Parity of the matrix-free operator, right-hand side, preconditioner diagonal and the persistent PCG solve
against the reference's assembled sparse system and its SuperLU solution (goldens: tests/golden/systems.npz,
recorded at the reference's own call sites for HS / BA / Classic+NL / classic++ / classic-c at alpha = 1, .5, 0)."""
import numpy as np
import pytest

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


def _rel(got, want):
    return float(np.max(np.abs(got - want)) / max(1e-300, np.max(np.abs(want))))


def _blend(ope, alpha, uv, duv, It, Ix, Iy):
    qua = ope._qua()
    Aq, bq, _, _ = qua.flow_operator(uv, duv, It, Ix, Iy)
    Ar, br, _, _ = ope.flow_operator(uv, duv, It, Ix, Iy)
    if alpha == 1:
        return Aq, bq
    if alpha == 0:
        return Ar, br
    return alpha * Aq + (1 - alpha) * Ar, alpha * bq + (1 - alpha) * br


def _f(a):          # goldens are stored (H, W, 2); the reference's vectors are column-major [u(:); v(:)]
    return a.reshape(-1, order="F")


def test_hs_system(systems, stages):
    from optical_flow import load_of_method
    ope = load_of_method("hs-brightness")
    ope.images = stages["scale_0_255"]
    uv = systems["uv"]
    A, b, _, it = ope.flow_operator(uv)
    assert A.shape == (2 * uv.shape[0] * uv.shape[1],) * 2 and it is True
    assert _rel(A @ _f(systems["probe"]), _f(systems["hs_Ap"])) < 1e-11
    assert _rel(b, _f(systems["hs_b"])) < 1e-9
    assert _rel(A.diagonal(), _f(systems["hs_diag"])) < 1e-12
    ope.exact_rtol = 1e-10
    x = ope._solve_linear_system(A, b, uv.shape)
    assert ope.last_stats["relres"] <= 1e-10
    assert_close(x, systems["hs_x"], 1e-6, "HS solve vs spsolve (pcg %r)" % (ope.last_stats,))


@pytest.mark.parametrize("tag,preset", [("ba", "ba"), ("cnl", "classic+nl"), ("cpp", "classic++"), ("cc", "classic-c")])
@pytest.mark.parametrize("alpha", [1.0, 0.5, 0.0])
@pytest.mark.parametrize("precision", ["mixed", "mixed-jacobi", "fp64"])
def test_gnc_system(systems, tag, preset, alpha, precision):
    from optical_flow import load_of_method
    ope = load_of_method(preset)
    ope.solver_precision = precision
    uv = systems["uv"]
    It, Ix, Iy = systems[tag + "_It"], systems[tag + "_Ix"], systems[tag + "_Iy"]
    A, b = _blend(ope, alpha, uv, np.zeros_like(uv), It, Ix, Iy)
    k = "%s_a%g" % (tag, alpha)
    assert _rel(A @ _f(systems["probe"]), _f(systems[k + "_Ap"])) < 1e-11, "A @ probe"
    assert _rel(b, _f(systems[k + "_b"])) < 1e-9, "rhs"
    assert _rel(A.diagonal(), _f(systems[k + "_diag"])) < 1e-12, "diag"
    ope.exact_rtol = 1e-10
    x = ope._solve_linear_system(A, b, uv.shape)
    # the direct solve itself is only accurate to ~cond * eps; weights span 6+ decades for the Charbonnier family
    assert_close(x, systems[k + "_x"], 2e-6, "%s %s solve vs spsolve (pcg %r)" % (k, precision, ope.last_stats))
    # the convergence criterion is the fp64 TRUE residual in both precisions: recompute it here with the assembled matrix
    true_rel = float(np.linalg.norm(b - A @ _f(x)) / np.linalg.norm(b))
    assert true_rel <= 1.5e-10, "%s %s: true relative residual %.2e" % (k, precision, true_rel)


@pytest.mark.parametrize("tag,preset", [("ba", "ba"), ("cnl", "classic+nl")])
def test_relinearised_operator(systems, tag, preset):
    """max_linear > 1 path: non-zero duv enters the IRLS weights and It_lin (classic_nl.py:312-346)."""
    from optical_flow import load_of_method
    ope = load_of_method(preset)
    A, b, _, _ = ope.flow_operator(systems["uv"], systems[tag + "_duv"], systems[tag + "_It"], systems[tag + "_Ix"],
                                   systems[tag + "_Iy"])
    assert _rel(A @ _f(systems["probe"]), _f(systems[tag + "_lin_Ap"])) < 1e-11
    assert _rel(b, _f(systems[tag + "_lin_b"])) < 1e-9


def test_solver_names():
    from optical_flow import load_of_method
    ope = load_of_method("ba")
    ope.solver = "cholesky"
    with pytest.raises(ValueError):
        ope._apply_solver(ope._c_params())


def test_solver_precision_names():
    from optical_flow import load_of_method
    ope = load_of_method("ba")
    ope.solver_precision = "fp16"
    with pytest.raises(ValueError):
        ope._apply_solver(ope._c_params())


@pytest.mark.parametrize("precision", ["mixed", "mixed-jacobi", "fp64"])
def test_pcg_determinism_and_batch(systems, precision):
    """Same system solved twice gives bit-identical x (fixed-order reductions, no fp atomics)."""
    from optical_flow import load_of_method
    ope = load_of_method("classic+nl")
    ope.solver_precision = precision
    uv = systems["uv"]
    A, b, _, _ = ope.flow_operator(uv, np.zeros_like(uv), systems["cnl_It"], systems["cnl_Ix"], systems["cnl_Iy"])
    x1 = ope._solve_linear_system(A, b, uv.shape)
    x2 = ope._solve_linear_system(A, b, uv.shape)
    np.testing.assert_array_equal(x1, x2)


@pytest.mark.parametrize("precision", ["mixed", "mixed-jacobi", "fp64"])
def test_exact_solver_iteration_cap(systems, precision):
    """exact_maxiter too small: the solve stops at the cap, reports it (iterations, residual above tol) and still
    returns the finite iterate it reached -- no hang, no exception from the persistent kernel."""
    from optical_flow import load_of_method
    ope = load_of_method("classic+nl")
    ope.solver_precision = precision
    ope.exact_maxiter = 5
    uv = systems["uv"]
    A, b, _, _ = ope.flow_operator(uv, np.zeros_like(uv), systems["cnl_It"], systems["cnl_Ix"], systems["cnl_Iy"])
    x = ope._solve_linear_system(A, b, uv.shape)
    assert np.isfinite(x).all()
    assert ope.last_stats["pcg_iters"] == 5 and ope.last_stats["relres"] > 1e-10
    true_rel = float(np.linalg.norm(b - A @ _f(x)) / np.linalg.norm(b))
    assert abs(true_rel - ope.last_stats["relres"]) <= 0.05 * true_rel + 1e-12     # the reported residual is the TRUE one
    assert true_rel < 1.0


def test_ic_preconditioner_halves_the_iterations(systems):
    """The tile-local block-IC(0) preconditioner of the default solver against block Jacobi on the same systems (alpha = 0:
    generalized Charbonnier weights spanning decades; alpha = 1: quadratic): same solution, at most 60 % of the iterations."""
    from optical_flow import load_of_method
    uv = systems["uv"]
    for alpha in (0.0, 1.0):
        its, xs = {}, {}
        for prec in ("mixed", "mixed-jacobi"):
            ope = load_of_method("classic+nl")
            ope.solver_precision = prec
            A, b = _blend(ope, alpha, uv, np.zeros_like(uv), systems["cnl_It"], systems["cnl_Ix"], systems["cnl_Iy"])
            xs[prec] = ope._solve_linear_system(A, b, uv.shape)
            its[prec] = ope.last_stats["pcg_iters"]
        assert_close(xs["mixed"], xs["mixed-jacobi"], 2e-6, "IC vs Jacobi solution, alpha=%g" % alpha)
        assert its["mixed"] <= 0.6 * its["mixed-jacobi"], its


@pytest.mark.parametrize("precision", ["mixed", "mixed-jacobi", "fp64"])
def test_zero_motion_gives_zero_increment(precision):
    """Identical frames: It is rounding noise (1e-13), so the right-hand side is ~0: the solvers must return the ~0
    increment (measured 5e-15 px) without dividing 0 by 0 in their scalar updates."""
    from optical_flow import load_of_method
    from optical_flow.utils.derivatives import partial_deriv
    rng = np.random.default_rng(5)
    im = rng.random((45, 70)) * 255
    images = np.stack([im, im], axis=2)
    ope = load_of_method("ba")
    ope.solver_precision = precision
    ope.images = images
    uv = np.zeros((45, 70, 2))
    It, Ix, Iy = partial_deriv(images, uv, ope.interpolation_method, ope.deriv_filter, 0.5)
    A, b, _, _ = ope.flow_operator(uv, np.zeros_like(uv), It, Ix, Iy)
    x = ope._solve_linear_system(A, b, uv.shape)
    assert np.isfinite(x).all() and np.abs(x).max() < 1e-12


@pytest.mark.parametrize("h,w", [(19, 37), (9, 33), (26, 70)])
def test_exact_solvers_on_ragged_sizes_vs_dense_solve(h, w):
    """Sizes that are not multiples of the solver's 8 x 32 strips (partial strips, a single strip row, one extra
    column): the operator is probed column by column into a dense matrix, solved with LAPACK, and compared with the
    three device solvers."""
    import synth
    from optical_flow import load_of_method
    from optical_flow.utils.derivatives import partial_deriv
    im1, im2, _ = synth.gray_pair(h, w, seed=h * w)
    images = np.stack([im1, im2], axis=2)
    ope = load_of_method("classic+nl")
    ope.images = images
    rng = np.random.default_rng(h + w)
    uv = 0.3 * rng.standard_normal((h, w, 2))
    It, Ix, Iy = partial_deriv(images, uv, ope.interpolation_method, ope.deriv_filter, 0.5)
    A, b, _, _ = ope.flow_operator(uv, np.zeros_like(uv), It, Ix, Iy)
    n = 2 * h * w
    dense = np.empty((n, n))
    e = np.zeros(n)
    for k in range(n):
        e[k] = 1.0
        dense[:, k] = A @ e
        e[k] = 0.0
    assert np.abs(dense - dense.T).max() <= 1e-12 * np.abs(dense).max()          # symmetric operator
    want = np.linalg.solve(dense, b).reshape((h, w, 2), order="F")
    for precision in ("mixed", "mixed-jacobi", "fp64"):
        ope.solver_precision = precision
        x = ope._solve_linear_system(A, b, (h, w, 2))
        assert_close(x, want, 2e-6, "%dx%d %s solve vs dense LAPACK solve (pcg %r)" % (w, h, precision, ope.last_stats))
        assert ope.last_stats["relres"] <= 1e-10


# ---------------------------------------------------------------------------------------------------------------
# the reference's own approximate solver modes
# ---------------------------------------------------------------------------------------------------------------
def test_reference_pcg_mode_e2e(stages):
    """params={'solver': 'pcg'}: scipy cg with the Jacobi preconditioner, rtol 1e-3, maxiter 200 (base.py:116-136).
    Approximate by design (0.14 px off the direct solve on this crop), but the SAME approximation as the reference's."""
    from optical_flow import estimate_flow
    g = load_golden("e2e_classicpnl-fast.npz")
    uv = estimate_flow(stages["rgb1"].astype(float), stages["rgb2"].astype(float), "classic+nl-fast", {"solver": "pcg"})
    assert_close(uv, g["uv_pcg_default"], 1e-5, "solver='pcg' end to end vs the reference's solver='pcg'")
    assert np.abs(uv - g["uv"]).max() > 1e-2          # and it really is the approximate mode, not the exact one


@pytest.mark.parametrize("tag,preset", [("hs", "hs-brightness"), ("cnl", "classic+nl"), ("ba", "ba")])
def test_sor_system(tag, preset):
    """solver='sor': lexicographic SOR (omega 1.9, tol 1e-2) swept along anti-diagonals reproduces the reference's
    iterates, hence its (unconverged, approximate) answer."""
    from optical_flow import load_of_method
    g = load_golden("sor.npz")
    ope = load_of_method(preset)
    ope.solver = "sor"
    uv = g["uv"]
    if tag == "hs":
        ope.images = g["gray"]
        A, b, _, _ = ope.flow_operator(uv)
    else:
        ope.images = g["tex"]
        A, b, _, _ = ope.flow_operator(uv, np.zeros_like(uv), g[tag + "_It"], g[tag + "_Ix"], g[tag + "_Iy"])
    x = ope._solve_linear_system(A, b, uv.shape)
    assert_close(x, g[tag + "_x"], 1e-9, "solver='sor' %s system vs the reference's SOR" % tag)


@pytest.mark.parametrize("preset,params", [("hs-brightness", {"solver": "sor"}),
                                           ("ba-brightness", {"solver": "sor", "max_iters": 2})])
def test_sor_e2e(preset, params):
    from optical_flow import estimate_flow
    g = load_golden("sor.npz")
    uv = estimate_flow(g["rgb1"].astype(float), g["rgb2"].astype(float), preset, params)
    assert_close(uv, g["e2e_" + preset], 1e-3, "solver='sor' estimate_flow(%s)" % preset)


def test_operator_as_sparse_matrix_and_custom_rhs(systems):
    """API edges of the reference's flow_operator / _solve_linear_system (base.py:87-114): the operator as a scipy.sparse
    matrix (FlowOperator.tocsc, recovered exactly by 18 probing products) and a solve with the caller's own right-hand side."""
    from scipy import sparse
    from optical_flow import load_of_method
    ope = load_of_method("classic+nl")
    uv = systems["uv"]
    It, Ix, Iy = systems["cnl_It"], systems["cnl_Ix"], systems["cnl_Iy"]
    A, b, _, _ = ope.flow_operator(uv, np.zeros_like(uv), It, Ix, Iy)
    M = A.tocsc()
    assert sparse.issparse(M) and M.shape == A.shape
    assert abs(M - M.T).max() <= 1e-12 * abs(M).max(), "symmetric"
    assert M.nnz <= 6 * A.shape[0]
    probe = _f(systems["probe"])
    assert _rel(M @ probe, _f(systems["cnl_a0_Ap"])) < 1e-11
    assert _rel(M.diagonal(), _f(systems["cnl_a0_diag"])) < 1e-12
    rng = np.random.default_rng(11)
    rhs = rng.standard_normal(b.shape)
    x = ope._solve_linear_system(A, rhs, uv.shape)
    true_rel = float(np.linalg.norm(rhs - M @ _f(x)) / np.linalg.norm(rhs))
    assert true_rel <= 1e-10, true_rel


@pytest.mark.parametrize("a_s,a_d,sig_s,sig_d", [(0.45, 0.45, 1e-3, 1e-3), (0.3, 0.3, 0.01, 0.5), (0.9, 0.9, 1.0, 1.0),
                                                 (0.45, 0.25, 1e-3, 1e-3), (0.2, 0.45, 1e-3, 1e-2)])
def test_generalized_charbonnier_weights_table_vs_oracle(a_s, a_d, sig_s, sig_d):
    """csrc/warp.cu pow_tab: the IRLS weight 2a (sig^2 + x^2)^(a-1) (penalties.py:121-128) through the exponent / mantissa
    table + binomial series, entry by entry against the oracle's numpy power over 30 decades of flow differences --
    including the exp/log fallback outside 2^+-64, exponents the table was not built for (a_s != a_d: only the data
    term's uses it) and the two-constant form of the Classic+NL-shaped sets.  Tolerance: 2e-14 RELATIVE per entry."""
    import flow_oracle as fo
    from optical_flow import load_of_method
    from optical_flow.robust import RobustFunction
    rng = np.random.default_rng(11)
    H, W = 37, 70                                       # ragged against the 32 x 8 tiles: halo edges on every side
    mag = 10.0 ** rng.uniform(-9, 5, size=(H, W, 2))
    uv = mag * rng.choice([-1.0, 1.0], size=(H, W, 2))
    uv[5, 7] = (3e12, -2e13)                            # y ~ 1e25 > 2^64: exp/log fallback
    uv[20, 33:36, 0] = 0.25                             # exact ties: delta = 0 -> y = sig^2
    duv = 10.0 ** rng.uniform(-6, 0, size=(H, W, 2)) * rng.choice([-1.0, 1.0], size=(H, W, 2))
    It = 10.0 ** rng.uniform(-8, 3, size=(H, W)) * rng.choice([-1.0, 1.0], size=(H, W))
    Ix, Iy = rng.normal(size=(H, W)) * 20, rng.normal(size=(H, W)) * 20
    ope = load_of_method("classic+nl")
    kind = "generalized_charbonnier"
    ope.rho_spatial_u = [RobustFunction(kind, sig_s, a_s), RobustFunction(kind, sig_s, a_s)]
    ope.rho_spatial_v = [RobustFunction(kind, sig_s, a_s), RobustFunction(kind, sig_s, a_s)]
    ope.rho_data = RobustFunction(kind, sig_d, a_d)
    A, b, _, _ = ope.flow_operator(uv, duv, It, Ix, Iy)
    M = A.tocsr()
    rs, rd = (kind, (sig_s, a_s)), (kind, (sig_d, a_d))
    q = ("quadratic", (1.0,))
    spec = dict(rho_su=[rs, rs], rho_sv=[rs, rs], rho_d=rd, qua_su=[q, q], qua_sv=[q, q], qua_d=q, lam=ope.lambda_,
                lam_q=ope.lambda_)
    ref = fo.assemble(uv, duv, It, Ix, Iy, spec, 0.0)
    N = H * W
    pix = np.arange(N).reshape(H, W, order="F")

    def entries(i, j):
        return -np.asarray(M[i.ravel(), j.ravel()]).reshape(i.shape)

    worst = 0.0
    for comp, (kh, kv) in enumerate((("wuh", "wuv"), ("wvh", "wvv"))):
        got_h = entries(comp * N + pix[:, :-1], comp * N + pix[:, 1:])
        got_v = entries(comp * N + pix[:-1, :], comp * N + pix[1:, :])
        for got, want in ((got_h, ref[kh][:, :-1]), (got_v, ref[kv][:-1, :])):
            assert np.all(want > 0)
            worst = max(worst, float(np.max(np.abs(got - want) / want)))
    assert worst <= 2e-14, "edge weights: worst relative error %.2e" % worst
    a12 = np.asarray(M[pix.ravel(), N + pix.ravel()]).reshape(H, W)
    err = float(np.max(np.abs(a12 - ref["a12"]) / np.abs(ref["a12"])))
    assert err <= 2e-14, "data weights (through a12): worst relative error %.2e" % err
    assert _rel(b, np.stack([ref["bu"], ref["bv"]], axis=2).reshape(-1, order="F")) < 1e-12
