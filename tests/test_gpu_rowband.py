"""Row-band split of one pair.  On the one-GPU test box the two ranks are two contexts of this process on
two host threads (b200flow_band_init(..., same_device=1): each rank's persistent solver takes one CTA per SM so that both
are resident, and the 'peer' block is an ordinary pointer instead of a CUDA IPC mapping).  The kernels, the halo loads, the
flag barriers and the exchange of the solution bands are the ones the multi-GPU job runs (bench.py --mode rowband).
Criterion: the band-wise solve is the SAME iteration as the single-GPU solve (bands are whole 8-row strips, so the tile-local
IC preconditioner is unchanged; only the summation order of the dot products differs) -- a teacher-forced solve within
1e-9 px and a short end-to-end run within 1e-6 px of the single-GPU result."""
import os
import threading

import numpy as np
import pytest

from conftest import assert_close
import synth

pytestmark = pytest.mark.gpu


def _run_two_ranks(fn, H, W):
    """fn(rank) is called on two threads, each with its own thread-local default context in row-band mode"""
    from optical_flow import _lib
    from optical_flow.rowband import RowBand, arena_bytes_for
    os.environ["B200FLOW_BAND_MIN_PIXELS"] = "1024"
    bar = threading.Barrier(2)
    exports, results, errors = [None, None], [None, None], []

    def worker(rank):
        try:
            ctx = _lib.default_context(0)
            rb = RowBand(ctx, rank, 2, arena_bytes_for(H, W), same_device=True)
            exports[rank] = rb.export()
            bar.wait(timeout=60)
            rb.connect(exports, same_process=True)
            bar.wait(timeout=60)
            try:
                results[rank] = fn(rank)
            finally:
                bar.wait(timeout=120)
                rb.close()
        except Exception as e:          # noqa: BLE001
            errors.append((rank, repr(e)))
            try:
                bar.abort()
            except Exception:           # noqa: BLE001
                pass
    th = [threading.Thread(target=worker, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    os.environ.pop("B200FLOW_BAND_MIN_PIXELS", None)
    assert not errors, errors
    assert all(not t.is_alive() for t in th), "a rank is stuck"
    return results


def test_rowband_teacher_forced_solve():
    from optical_flow import load_of_method
    from optical_flow.utils.derivatives import partial_deriv
    H, W = 136, 200
    im1, im2, flow = synth.gray_pair(H, W, seed=5)
    images = np.stack([im1, im2], axis=2)
    uv = 0.8 * flow

    def solve(rank=None):
        ope = load_of_method("classic++")
        ope.images = images
        It, Ix, Iy = partial_deriv(images, uv, ope.interpolation_method, ope.deriv_filter, ope.blend)
        A = ope.flow_operator(uv, np.zeros_like(uv), It, Ix, Iy)[0]
        x = ope._solve_linear_system(A, A.b, uv.shape)
        return x, ope.last_stats
    want, st = solve()
    got = _run_two_ranks(solve, H, W)
    for rank in range(2):
        x, st_r = got[rank]
        assert_close(x, want, 1e-9, "row-band solve, rank %d vs single GPU" % rank)
        assert abs(st_r["pcg_iters"] - st["pcg_iters"]) <= 2, (st_r, st)
    np.testing.assert_array_equal(got[0][0], got[1][0], err_msg="both ranks must hold the same bits")


def test_rowband_end_to_end_short():
    from optical_flow import estimate_flow
    H, W = 120, 176
    im1, im2, _ = synth.gray_pair(H, W, seed=6)
    want = estimate_flow(im1, im2, "ba", {"max_iters": 2})
    got = _run_two_ranks(lambda rank: estimate_flow(im1, im2, "ba", {"max_iters": 2}), H, W)
    for rank in range(2):
        assert_close(got[rank], want, 1e-6, "row-band ba max_iters=2, rank %d vs single GPU" % rank)
    np.testing.assert_array_equal(got[0], got[1])
