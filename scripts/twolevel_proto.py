"""CPU prototype (design evidence, not product): does a coarse-space correction on top of the tile-local block-IC(0)
preconditioner cut the PCG iteration count enough to pay for one more grid barrier per iteration?
Real Classic+NL system (RubberWhale 584x388, reference's final flow, alpha = 0 and 1), solved to 1e-12."""
import os
import sys
import time
import numpy as np
from scipy import sparse
from scipy.sparse.linalg import splu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mg_proto import make_systems, fo  # noqa: E402
from ic_proto import factor, apply  # noqa: E402


def pcg(A, b, prec, tol=1e-12, maxit=2000):
    x = np.zeros_like(b)
    r = b.copy()
    z = prec(r)
    p = z.copy()
    rz = r @ z
    bb = np.sqrt(b @ b)
    for k in range(1, maxit + 1):
        Ap = A @ p
        a = rz / (p @ Ap)
        x += a * p
        r -= a * Ap
        if np.sqrt(r @ r) <= tol * bb:
            return x, k
        z = prec(r)
        rz2 = r @ z
        p = z + (rz2 / rz) * p
        rz = rz2
    return x, maxit


def aggregation(H, W, T, kind="const"):
    """P: (2HW) x (2 nc) piecewise-constant (or bilinear hat) prolongation, per component"""
    hc, wc = -(-H // T), -(-W // T)
    ii, jj = np.indices((H, W))
    if kind == "const":
        col = ((ii // T) * wc + jj // T).ravel()
        P1 = sparse.csr_matrix((np.ones(H * W), (np.arange(H * W), col)), shape=(H * W, hc * wc))
    else:   # bilinear hats on the aggregate centres
        cy = (ii + 0.5) / T - 0.5
        cx = (jj + 0.5) / T - 0.5
        y0 = np.clip(np.floor(cy).astype(int), 0, hc - 1); x0 = np.clip(np.floor(cx).astype(int), 0, wc - 1)
        y1 = np.clip(y0 + 1, 0, hc - 1); x1 = np.clip(x0 + 1, 0, wc - 1)
        fy = np.clip(cy - y0, 0, 1); fx = np.clip(cx - x0, 0, 1)
        rows = np.tile(np.arange(H * W), 4)
        cols = np.concatenate([(y0 * wc + x0).ravel(), (y0 * wc + x1).ravel(), (y1 * wc + x0).ravel(), (y1 * wc + x1).ravel()])
        vals = np.concatenate([((1 - fy) * (1 - fx)).ravel(), ((1 - fy) * fx).ravel(), (fy * (1 - fx)).ravel(), (fy * fx).ravel()])
        P1 = sparse.csr_matrix((vals, (rows, cols)), shape=(H * W, hc * wc))
    return sparse.block_diag([P1, P1]).tocsr()


def main():
    systems = make_systems()
    for alpha in (0.0, 1.0):
        s = systems[alpha]
        H, W = s["a11"].shape
        A = fo.to_sparse(s).tocsr()
        b = np.concatenate([s["bu"].ravel(), s["bv"].ravel()])
        dg = fo.operator_diag(s)
        f = factor(dg[:, :, 0], s["a12"], dg[:, :, 1], s["wuh"], s["wuv"], s["wvh"], s["wvv"], 8, 8, 0.0)

        def ic(r):
            r2 = np.stack([r[:H * W].reshape(H, W), r[H * W:].reshape(H, W)], axis=2)
            z = apply(*f, s["wuh"], s["wuv"], s["wvh"], s["wvv"], r2, 8, 8)
            return np.concatenate([z[:, :, 0].ravel(), z[:, :, 1].ravel()])
        t = time.time()
        x0, it0 = pcg(A, b, ic)
        print("alpha=%g  tile-IC(8x8): %d it (%.0fs)" % (alpha, it0, time.time() - t), flush=True)
        for T in (8, 16, 32):
            for kind in ("const", "bilinear"):
                P = aggregation(H, W, T, kind)
                Ac = (P.T @ A @ P).tocsc()
                lu = splu(Ac)

                def additive(r):
                    return ic(r) + P @ lu.solve(P.T @ r)

                def adef2(r):     # z = M^-1 r + Q (r - A M^-1 r)
                    z = ic(r)
                    return z + P @ lu.solve(P.T @ (r - A @ z))
                for name, prec in (("additive", additive), ("adef2", adef2)):
                    t = time.time()
                    x, it = pcg(A, b, prec)
                    print("  T=%2d %-8s %-8s coarse %dx2: %3d it (x%.2f)  |dx| %.1e (%.0fs)" % (
                        T, kind, name, Ac.shape[0] // 2, it, it0 / it, np.abs(x - x0).max(), time.time() - t), flush=True)


if __name__ == "__main__":
    main()
