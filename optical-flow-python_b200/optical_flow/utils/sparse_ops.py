"""Sparse convolution / filtering matrices (reference: optical_flow/utils/sparse_ops.py:6-220).

Host-side helpers kept for API compatibility: the reference's drivers build their linear systems from these
matrices, the B200 path does not (its operator is the matrix-free five-point stencil of csrc/solve*.cu), so nothing
here is on the hot path or touches the GPU.  Same names, argument meaning, vectorisation convention (column-major,
`vec(X) = X.ravel(order='F')`) and error behaviour as the reference.

Construction: a separable shift structure instead of per-entry index lists.  With column-major vectorisation
`vec(A X B^T) = (B kron A) vec(X)`, so a tap F[a, b] that moves rows by a and columns by b contributes
`F[a, b] * kron(C_b, R_a)` with two small 0/1 matrices; boundary handling only changes how R_a / C_b map an output
index to its source index.
"""
import numpy as np
from scipy import sparse


def _shift_full(n, k, a):
    """(n + k - 1) x n matrix with ones at (i + a, i): places a length-n signal at offset a of the full output."""
    return sparse.coo_matrix((np.ones(n), (np.arange(n) + a, np.arange(n))), shape=(n + k - 1, n)).tocsr()


def _select(n_total, start, count):
    """count x n_total matrix picking rows start .. start + count - 1."""
    return sparse.coo_matrix((np.ones(count), (np.arange(count), np.arange(count) + start)),
                             shape=(count, n_total)).tocsr()


def _embed(n_total, start, count):
    """n_total x count matrix writing a length-count signal at offset start (zeros elsewhere)."""
    return _select(n_total, start, count).T.tocsr()


def convmtxn(F, sz):
    """M such that M @ vec(X) = vec(conv2(X, F, 'full')), column-major vec (sparse_ops.py:6-56)."""
    F = np.atleast_2d(np.asarray(F, dtype=float))
    H, W = int(sz[0]), int(sz[1])
    fh, fw = F.shape
    M = sparse.csc_matrix(((H + fh - 1) * (W + fw - 1), H * W))
    for a, b in zip(*np.nonzero(F)):
        M = M + F[a, b] * sparse.kron(_shift_full(W, fw, b), _shift_full(H, fh, a), format="csc")
    return M.tocsc()


def make_convn_mat(F, sz, shape='full', pad=None):
    """Convolution matrix with output-shape control: 'full', 'same' (centre crop, offset (f - 1) // 2), 'valid';
    shape='valid' with pad='sameswap' writes the valid result into a same-sized output (zero rows elsewhere)
    (sparse_ops.py:59-118).  Unknown shape -> ValueError."""
    F = np.atleast_2d(np.asarray(F, dtype=float))
    H, W = int(sz[0]), int(sz[1])
    fh, fw = F.shape
    if shape not in ('full', 'same', 'valid'):
        raise ValueError(f"Unknown shape: {shape}")
    M_full = convmtxn(F, (H, W))
    if shape == 'full':
        return M_full
    Hf, Wf = H + fh - 1, W + fw - 1
    if shape == 'same':
        rows = sparse.kron(_select(Wf, (fw - 1) // 2, W), _select(Hf, (fh - 1) // 2, H), format="csr")
        return (rows @ M_full).tocsc()
    Hv, Wv = H - fh + 1, W - fw + 1
    if Hv <= 0 or Wv <= 0:
        return sparse.csc_matrix((0, H * W))
    rows = sparse.kron(_select(Wf, fw - 1, Wv), _select(Hf, fh - 1, Hv), format="csr")
    M_valid = rows @ M_full
    if pad == 'sameswap':
        place = sparse.kron(_embed(W, (fw - 1) // 2, Wv), _embed(H, (fh - 1) // 2, Hv), format="csr")
        return (place @ M_valid).tocsc()
    return M_valid.tocsc()


def _source_map(n, offset, boundary):
    """n x n 0/1 matrix: output index i reads source index i + offset under the boundary rule (rows of out-of-range
    outputs are empty for the zero boundary)."""
    i = np.arange(n)
    s = i + offset
    if boundary == '0':
        keep = (s >= 0) & (s < n)
        i, s = i[keep], s[keep]
    elif boundary == 'replicate':
        s = np.clip(s, 0, n - 1)
    elif boundary == 'symmetric':
        s = np.where(s < 0, -s - 1, s)
        s = np.where(s >= n, 2 * n - s - 1, s)
        s = np.clip(s, 0, n - 1)
    else:                      # the reference silently builds an empty matrix for any other string
        i, s = i[:0], s[:0]
    return sparse.coo_matrix((np.ones(len(i)), (i, s)), shape=(n, n)).tocsr()


def make_imfilter_mat(F, sz, boundary='replicate', shape='same'):
    """Matrix of MATLAB's imfilter(X, F, boundary, 'same', 'corr'): out(i, j) = sum F[a, b] X(i + a - hf, j + b - wf)
    with hf, wf = (size(F) - 1) // 2; boundary 'replicate' (nearest), '0' (zero) or 'symmetric' (mirror with the edge
    sample repeated); column-major vec (sparse_ops.py:128-220)."""
    F = np.atleast_2d(np.asarray(F, dtype=float))
    H, W = int(sz[0]), int(sz[1])
    hf, wf = (F.shape[0] - 1) // 2, (F.shape[1] - 1) // 2
    M = sparse.csc_matrix((H * W, H * W))
    for a, b in zip(*np.nonzero(F)):
        M = M + F[a, b] * sparse.kron(_source_map(W, b - wf, boundary), _source_map(H, a - hf, boundary), format="csc")
    return M.tocsc()
