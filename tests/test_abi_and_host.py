"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200flow.h declares (no compute calls
without a GPU), the ctypes mirror of b200flow_params matches the header field for field, presets / parameter
parsing / error behaviour follow the reference, and the product fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _header():
    return open(os.path.join(ROOT, "include", "b200flow.h")).read()


def test_library_exports_every_declared_symbol():
    from optical_flow import _lib
    lib = _lib.load_library()
    declared = sorted(set(re.findall(r"\b(b200flow_[a-z0-9_]+)\s*\(", _header())))
    assert declared, "no prototypes found"
    for name in declared:
        assert hasattr(lib, name), "libb200flow.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared
    assert lib.b200flow_abi_version() == 2


def test_params_struct_matches_header():
    from optical_flow import _lib
    h = _header()
    end = h.index("} b200flow_params;")
    body = h[h.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(int|double|b200flow_penalty)\s+", "", decl)
        for n in decl.split(","):
            names.append(re.sub(r"\[.*\]", "", n).strip())
    mirror = [n for n, _ in _lib.Params._fields_]
    assert [n.rstrip("_") for n in mirror] == [n.rstrip("_") for n in names]
    import ctypes
    assert ctypes.sizeof(_lib.Penalty) == 24


def test_no_cpu_fallback():
    """Without a GPU every operator must raise instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from optical_flow import estimate_flow
    from optical_flow.utils.derivatives import partial_deriv
    with pytest.raises(RuntimeError, match="no CPU fallback|cannot create a context"):
        estimate_flow(np.zeros((32, 32)), np.zeros((32, 32)), "hs-brightness")
    with pytest.raises(RuntimeError):
        partial_deriv(np.zeros((16, 16, 2)), np.zeros((16, 16, 2)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "optical-flow-python_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(d, f)).read()
                assert "flow_oracle" not in src and "import oracle" not in src, f


def test_presets_match_reference_table():
    """SURVEY.md App. C / methods/config.py:10-176."""
    from optical_flow import load_of_method
    from optical_flow.methods import HSOpticalFlow, BAOpticalFlow, ClassicNLOpticalFlow
    o = load_of_method("classic+nl-fast")
    assert isinstance(o, ClassicNLOpticalFlow) and (o.max_iters, o.gnc_iters, o.display) == (3, 2, True)
    assert o.texture and o.interpolation_method == "bi-cubic" and o.area_hsz == 7 and o.sigma_i == 7
    assert o.lambda_ == 3 and o.lambda_q == 3 and o.median_filter_size == [5, 5] and o.color_images is not None
    assert o.rho_data.method == "generalized_charbonnier" and list(o.rho_data.param) == [1e-3, 0.45]
    o = load_of_method("classic+nl")
    assert (o.max_iters, o.gnc_iters, o.display) == (10, 3, False)
    assert load_of_method("classic+nl-full").fullVersion is True
    o = load_of_method("hs-brightness")
    assert isinstance(o, HSOpticalFlow) and o.lambda_ == 10 and not o.texture and o.max_warping_iters == 10
    o = load_of_method("hs")
    assert o.lambda_ == 40 and o.texture and o.display
    o = load_of_method("ba-brightness")
    assert isinstance(o, BAOpticalFlow) and o.lambda_ == 0.045 and o.rho_data.param[0] == 3.5 and not o.texture
    for name in ("ba", "classic-l"):
        o = load_of_method(name)
        assert o.lambda_ == 0.06 and o.rho_spatial_u[0].param[0] == 0.03 and o.rho_data.param[0] == 1.5 and o.texture
        assert o.interpolation_method == "cubic" and o.gnc_iters == 3 and o.max_iters == 10
    o = load_of_method("classic-c")
    assert o.lambda_ == 5 and o.rho_data.method == "charbonnier" and o.texture
    assert load_of_method("classic-c-brightness").lambda_ == 3
    o = load_of_method("classic++")
    assert o.interpolation_method == "bi-cubic" and o.rho_data.method == "generalized_charbonnier" and o.lambda_ == 3
    with pytest.raises(ValueError, match="Unknown optical flow method"):
        load_of_method("classic-x")


def test_parse_input_parameter_and_c_params():
    from optical_flow import load_of_method
    o = load_of_method("ba")
    o.parse_input_parameter({"lambda": 0.5, "max_iters": 4, "no_such_key": 1})
    assert o.lambda_ == 0.5 and o.max_iters == 4 and not hasattr(o, "no_such_key")
    o.parse_input_parameter(["gnc_iters", 2, "solver", "pcg"])
    assert o.gnc_iters == 2 and o.solver == "pcg"
    P = o._c_params(levels=3)
    o._apply_solver(P)
    assert (P.method, P.interp, P.gnc_iters, P.max_iters, P.pyramid_levels, P.solver) == (1, 1, 2, 4, 3, 1)
    assert P.tol == 1e-3 and P.maxit == 200                      # the reference's pcg_rtol / pcg_maxiter
    assert P.qua_su[0].kind == 0 and P.qua_su[0].p0 == 1.0 and abs(P.qua_d.p0 - 1.5 / 0.03) < 1e-12   # ba.py:150-160
    assert [P.deriv_filter[i] * 12 for i in range(5)] == [1, -8, 0, 8, -1]
    c = load_of_method("classic+nl")
    P = c._c_params(levels=5)
    assert P.qua_d.kind == 0 and P.qua_d.p0 == 1e-3 and P.rho_d.kind == 3 and P.rho_d.p1 == 0.45   # classic_nl.py:212-226
    assert (P.median_h, P.median_w, P.area_hsz) == (5, 5, 7)
    o.solver = "cholesky"
    with pytest.raises(ValueError, match="Unknown solver"):
        o._apply_solver(P)
    o.interpolation_method = "nearest"
    with pytest.raises(ValueError, match="Unknown interpolation"):
        o._c_params()
    assert c._auto_pyramid_levels(np.empty((388, 584, 2))) == 5


def test_robust_function_surface():
    from optical_flow.robust.robust_function import RobustFunction, PENALTY_MAP
    assert list(PENALTY_MAP) == ["quadratic", "lorentzian", "charbonnier", "generalized_charbonnier", "geman_mcclure",
                                 "huber", "tukey", "gaussian", "tdist", "tdist_unnorm"]
    rf = RobustFunction("generalized_charbonnier", 1e-3, 0.45)
    assert rf.method == "generalized_charbonnier" and list(rf.sigma) == [1e-3, 0.45] and rf.param is rf.sigma
    assert "generalized_charbonnier" in repr(rf)
    assert list(RobustFunction("quadratic").sigma) == [1.0]
    s = rf.c_struct()
    assert (s.kind, s.p0, s.p1) == (3, 1e-3, 0.45)
    with pytest.raises(ValueError, match="Unknown penalty method"):
        RobustFunction("nope", 1.0)


def test_flo_roundtrip_and_metrics(tmp_path):
    from optical_flow import read_flo, write_flo, flow_angular_error
    rng = np.random.default_rng(0)
    f = rng.standard_normal((7, 9, 2)).astype(np.float32)
    p = str(tmp_path / "a.flo")
    write_flo(f, p)
    np.testing.assert_array_equal(read_flo(p), f)
    with open(p, "r+b") as fh:
        fh.write(b"\x00\x00\x00\x00")
    with pytest.raises(ValueError, match="Invalid .flo file tag"):
        read_flo(p)
    with pytest.raises(ValueError):
        write_flo(np.zeros((3, 3)), p)
    u = rng.standard_normal((8, 8))
    v = rng.standard_normal((8, 8))
    aae, std, epe = flow_angular_error(u, v, u, v)
    assert aae < 1e-5 and epe == 0
    tu = u.copy()
    tu[0, 0] = 1e10                       # unknown-flow marker is masked out
    assert flow_angular_error(tu, v, u, v)[2] == 0


def test_shard_indices():
    from optical_flow.interface import shard_indices
    parts = [shard_indices(10, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == list(range(10)) and parts[1] == [1, 5, 9]


def test_pinned_result_buffers_are_recycled_only_when_released():
    """Context.pinned_empty (result buffers of estimate_flow_batch): a buffer goes back to the free list only when the
    array AND every view of it are gone; the allocator itself is faked here (no GPU)."""
    import ctypes as C
    import gc
    from optical_flow import _lib

    class Fake(_lib.Context):
        def __init__(self):
            self.n, self.keep = 0, []

        def call(self, name, nbytes, pref):
            assert name == "b200flow_host_alloc"
            self.n += 1
            b = C.create_string_buffer(nbytes.value)
            self.keep.append(b)
            C.cast(pref, C.POINTER(C.c_void_p))[0] = C.addressof(b)

        def __del__(self):
            pass

    f = Fake()
    a = f.pinned_empty((4, 5, 2))
    assert a.shape == (4, 5, 2) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"]
    a[:] = 1.5
    view, addr = a[1:], a.ctypes.data
    del a
    gc.collect()
    b = f.pinned_empty((4, 5, 2))                 # the view is alive: a NEW buffer
    assert b.ctypes.data != addr and f.n == 2 and float(view[0, 0, 0]) == 1.5
    del view
    gc.collect()
    c = f.pinned_empty((4, 5, 2))                 # released: recycled, no new allocation
    assert c.ctypes.data == addr and f.n == 2
    # a caller hoarding results cannot pin the host: past the cap the buffers are ordinary pageable arrays
    old_cap, _lib.PINNED_CAP_BYTES = _lib.PINNED_CAP_BYTES, 3 * 4 * 5 * 2 * 8
    try:
        d = f.pinned_empty((4, 5, 2))             # third owned buffer: still pinned
        e = f.pinned_empty((4, 5, 2))             # would be the fourth: pageable
        assert f.n == 3 and e.shape == (4, 5, 2) and e.flags["OWNDATA"] and not d.flags["OWNDATA"]
    finally:
        _lib.PINNED_CAP_BYTES = old_cap


def test_display_log_is_printed_in_the_reference_format(capsys):
    """Host side of display=True (methods/base.py _print_display_log): rows {stage, level, warp, linearisation, norm} from the
    device log become the reference's lines (classic_nl.py:141-152,255-256,186-198; hs.py:80-81,123-127) -- no GPU needed."""
    from optical_flow import load_of_method
    ope = load_of_method("classic+nl-fast")
    ope._stage_lines = 0
    rows = np.array([[0, 1, 0, 0, 8.5], [0, 1, 1, 0, 3.25], [0, 0, 0, 0, 12.0], [1, 0, 0, 0, 1.0], [1, 0, 0, 1, 0.5]])
    ope._print_display_log(rows, "gnc")
    out = capsys.readouterr().out.splitlines()
    assert out[:6] == ["GNC stage: 1", "  Pyramid level: 2", "    Iter: 1 1 (delta: 8.500000)", "    Iter: 2 1 (delta: 3.250000)",
                       "  Pyramid level: 1", "    Iter: 1 1 (delta: 12.000000)"]
    assert out[6].startswith("GNC stage 1 finished, ") and out[6].endswith(" minutes passed")
    assert out[7:] == ["GNC stage: 2", "  Pyramid level: 1", "    Iter: 1 1 (delta: 1.000000)", "    Iter: 1 2 (delta: 0.500000)"]
    assert ope._stage_lines == 1                    # the last stage's "finished" line is compute_flow's (it may carry AAE / EPE)
    ope._print_display_log(rows[:2], "gnc_base")   # compute_flow_base prints the iteration lines only
    assert capsys.readouterr().out.splitlines() == ["    Iter: 1 1 (delta: 8.500000)", "    Iter: 2 1 (delta: 3.250000)"]
    hs = load_of_method("hs")
    hs._print_display_log(np.array([[0, 1, 0, 0, 1.0], [0, 1, 1, 0, 5e-4], [0, 1, 2, 0, 0.3], [0, 0, 0, 0, 2.0]]), "hs")
    # Horn-Schunck leaves a level at the first ||x|| < 1e-3 (hs.py:126-127): later solves of that level are not printed
    assert capsys.readouterr().out.splitlines() == ["Pyramid level: 2", "  Iteration: 1  (norm: 1.000000)",
                                                    "  Iteration: 2  (norm: 0.000500)", "Pyramid level: 1",
                                                    "  Iteration: 1  (norm: 2.000000)"]
