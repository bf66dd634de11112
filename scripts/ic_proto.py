"""CPU prototype (design evidence, not product): tile-local block-IC(0)/MIC(0) preconditioner iteration counts."""
import os
import sys
import time
import numpy as np
import numba as nb

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mg_proto import make_systems, pcg, block_inv, apply_minv, fo  # noqa: E402


@nb.njit(cache=True)
def factor(d11, d12, d22, wuh, wuv, wvh, wvv, th, tw, mic):
    """pivot blocks P[i,j] (2x2 symmetric, stored inverted) of the block incomplete Cholesky in lexicographic order
    inside th x tw tiles; couplings that leave the tile are dropped.  mic: modified-IC weight (0 = IC)."""
    H, W = d11.shape
    i11 = np.zeros((H, W)); i12 = np.zeros((H, W)); i22 = np.zeros((H, W))
    for i in range(H):
        for j in range(W):
            p11 = d11[i, j]; p12 = d12[i, j]; p22 = d22[i, j]
            if j % tw != 0:
                # left neighbour (i, j-1), edge weights wuh[i,j-1], wvh[i,j-1]; off-diag block = -diag(wu, wv)
                wu = wuh[i, j - 1]; wv = wvh[i, j - 1]
                a11 = i11[i, j - 1]; a12 = i12[i, j - 1]; a22 = i22[i, j - 1]
                p11 -= wu * a11 * wu; p12 -= wu * a12 * wv; p22 -= wv * a22 * wv
                if mic > 0.0 and (i + 1) % th != 0 and i + 1 < H:
                    # fill-in that IC drops: left's coupling to its lower neighbour (i+1, j-1)
                    wu2 = wuv[i, j - 1]; wv2 = wvv[i, j - 1]
                    p11 -= mic * wu * a11 * wu2; p22 -= mic * wv * a22 * wv2
            if i % th != 0:
                wu = wuv[i - 1, j]; wv = wvv[i - 1, j]
                a11 = i11[i - 1, j]; a12 = i12[i - 1, j]; a22 = i22[i - 1, j]
                p11 -= wu * a11 * wu; p12 -= wu * a12 * wv; p22 -= wv * a22 * wv
                if mic > 0.0 and (j + 1) % tw != 0 and j + 1 < W:
                    wu2 = wuh[i - 1, j]; wv2 = wvh[i - 1, j]
                    p11 -= mic * wu * a11 * wu2; p22 -= mic * wv * a22 * wv2
            det = p11 * p22 - p12 * p12
            i11[i, j] = p22 / det; i12[i, j] = -p12 / det; i22[i, j] = p11 / det
    return i11, i12, i22


@nb.njit(cache=True)
def apply(i11, i12, i22, wuh, wuv, wvh, wvv, r, th, tw):
    H, W = i11.shape
    y = np.zeros((H, W, 2))
    # forward: (P + L) y = r  ->  y = P^-1 (r - L y)
    for i in range(H):
        for j in range(W):
            ru = r[i, j, 0]; rv = r[i, j, 1]
            if j % tw != 0:
                ru += wuh[i, j - 1] * y[i, j - 1, 0]; rv += wvh[i, j - 1] * y[i, j - 1, 1]
            if i % th != 0:
                ru += wuv[i - 1, j] * y[i - 1, j, 0]; rv += wvv[i - 1, j] * y[i - 1, j, 1]
            y[i, j, 0] = i11[i, j] * ru + i12[i, j] * rv
            y[i, j, 1] = i12[i, j] * ru + i22[i, j] * rv
    # backward: (I + P^-1 L^T) z = y
    z = np.zeros((H, W, 2))
    for i in range(H - 1, -1, -1):
        for j in range(W - 1, -1, -1):
            su = 0.0; sv = 0.0
            if (j + 1) % tw != 0 and j + 1 < W:
                su += wuh[i, j] * z[i, j + 1, 0]; sv += wvh[i, j] * z[i, j + 1, 1]
            if (i + 1) % th != 0 and i + 1 < H:
                su += wuv[i, j] * z[i + 1, j, 0]; sv += wvv[i, j] * z[i + 1, j, 1]
            z[i, j, 0] = y[i, j, 0] + i11[i, j] * su + i12[i, j] * sv
            z[i, j, 1] = y[i, j, 1] + i12[i, j] * su + i22[i, j] * sv
    return z


def main():
    systems = make_systems()
    for alpha in (1.0, 0.0):
        s = systems[alpha]
        b = np.stack([s["bu"], s["bv"]], axis=2)
        dg = fo.operator_diag(s)
        M = block_inv(s)
        x, it = pcg(s, b, lambda r: apply_minv(M, r))
        print("alpha=%g block-Jacobi: %d it" % (alpha, it))
        for th, tw in ((1 << 20, 1 << 20), (8, 32), (16, 32), (32, 32), (4, 64), (1, 1 << 20), (2, 1 << 20)):
            for mic in (0.0, 0.95):
                f = factor(dg[:, :, 0], s["a12"], dg[:, :, 1], s["wuh"], s["wuv"], s["wvh"], s["wvv"], th, tw, mic)
                t = time.time()
                try:
                    x2, it = pcg(s, b, lambda r: apply(*f, s["wuh"], s["wuv"], s["wvh"], s["wvv"], r, th, tw), maxit=600)
                    print("  tile %s x %s mic=%.2f: %3d it  |dx| %.1e  (%.1fs)" % (
                        "inf" if th > 1e5 else th, "inf" if tw > 1e5 else tw, mic, it, np.abs(x2 - x).max(), time.time() - t))
                except Exception as e:  # noqa
                    print("  tile", th, tw, mic, "failed", e)


if __name__ == "__main__":
    main()
