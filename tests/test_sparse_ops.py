"""Host-side sparse convolution / filtering matrices of the drop-in (optical_flow/utils/sparse_ops.py; reference
sparse_ops.py:6-220) against SciPy's own convolution and correlation on the same inputs.  Column-major vec throughout.
No GPU: these helpers are kept for API compatibility, the B200 path never builds a matrix."""
import numpy as np
import pytest
from scipy.ndimage import correlate
from scipy.signal import convolve2d

from optical_flow.utils.sparse_ops import convmtxn, make_convn_mat, make_imfilter_mat

FILTERS = [(1, 1), (1, 2), (2, 1), (3, 3), (2, 3), (5, 5), (1, 5)]
SIZES = [(4, 5), (7, 3), (6, 6)]


def _vec(a):
    return a.ravel(order="F")


@pytest.mark.parametrize("fs", FILTERS)
@pytest.mark.parametrize("sz", SIZES)
def test_convolution_matrices_match_convolve2d(fs, sz):
    rng = np.random.default_rng(fs[0] * 100 + fs[1] * 10 + sz[0])
    F, X = rng.standard_normal(fs), rng.standard_normal(sz)
    full = convolve2d(X, F, mode="full")
    M = convmtxn(F, sz)
    assert M.shape == (full.size, X.size)
    np.testing.assert_allclose(M @ _vec(X), _vec(full), atol=1e-12)
    np.testing.assert_allclose(make_convn_mat(F, sz, "full").toarray(), M.toarray())
    # 'same': centre crop with offset (f - 1) // 2 -- the MATLAB convention, which is convolve2d's for odd kernels
    oh, ow = (fs[0] - 1) // 2, (fs[1] - 1) // 2
    same = full[oh:oh + sz[0], ow:ow + sz[1]]
    np.testing.assert_allclose(make_convn_mat(F, sz, "same") @ _vec(X), _vec(same), atol=1e-12)
    if sz[0] >= fs[0] and sz[1] >= fs[1]:
        valid = convolve2d(X, F, mode="valid")
        Mv = make_convn_mat(F, sz, "valid")
        assert Mv.shape == (valid.size, X.size)
        np.testing.assert_allclose(Mv @ _vec(X), _vec(valid), atol=1e-12)
        # 'sameswap': the valid result written into a same-sized array at offset (f - 1) // 2, zeros elsewhere
        want = np.zeros(sz)
        want[oh:oh + valid.shape[0], ow:ow + valid.shape[1]] = valid
        Ms = make_convn_mat(F, sz, "valid", pad="sameswap")
        assert Ms.shape == (X.size, X.size)
        np.testing.assert_allclose(Ms @ _vec(X), _vec(want), atol=1e-12)


def test_valid_larger_than_image_is_empty_and_bad_shape_raises():
    assert make_convn_mat(np.ones((5, 5)), (3, 4), "valid").shape == (0, 12)
    with pytest.raises(ValueError, match="Unknown shape"):
        make_convn_mat(np.ones((2, 2)), (3, 3), "bogus")
    assert convmtxn(np.zeros((2, 2)), (3, 3)).nnz == 0


@pytest.mark.parametrize("fs", [(1, 1), (3, 3), (1, 3), (5, 1), (5, 5), (3, 5)])
@pytest.mark.parametrize("boundary,mode", [("replicate", "nearest"), ("0", "constant"), ("symmetric", "reflect")])
def test_imfilter_matrix_matches_ndimage_correlate(fs, boundary, mode):
    """imfilter(..., 'corr') == scipy.ndimage.correlate with the matching boundary mode (odd kernels: same centre)."""
    rng = np.random.default_rng(fs[0] * 7 + fs[1])
    sz = (6, 7)
    F, X = rng.standard_normal(fs), rng.standard_normal(sz)
    M = make_imfilter_mat(F, sz, boundary=boundary)
    assert M.shape == (X.size, X.size)
    np.testing.assert_allclose((M @ _vec(X)).reshape(sz, order="F"), correlate(X, F, mode=mode, cval=0.0), atol=1e-12)


def test_imfilter_identity_and_row_sums():
    M = make_imfilter_mat(np.array([[1.0]]), (4, 4))
    np.testing.assert_allclose(M @ np.arange(16.0), np.arange(16.0))
    F = np.array([[0.0, 1.0, 0.0], [1.0, -4.0, 1.0], [0.0, 1.0, 0.0]])
    ones = np.ones(20)
    np.testing.assert_allclose(make_imfilter_mat(F, (4, 5), "replicate") @ ones, 0.0, atol=1e-14)   # constants are in the Laplacian's kernel
    assert np.abs(make_imfilter_mat(F, (4, 5), "0") @ ones).max() > 0.5                               # but not with a zero boundary
