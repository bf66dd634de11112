"""CPU prototype (design evidence, not product): the tile-local block-IC(0) preconditioner exactly as the CUDA kernel
applies it -- 8 x 32 sub-tiles, fp32 factorisation and fp32 forward / backward sweeps in PUSH form -- inside an fp64
PCG, to check that single precision in the factor does not cost iterations.  Compare scripts/ic_proto.py (fp64)."""
import os
import sys
import numpy as np
import numba as nb

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mg_proto import make_systems, pcg, block_inv, apply_minv, fo  # noqa: E402

f32 = np.float32


@nb.njit(cache=True)
def factor32(d11, d12, d22, wh_u, wh_v, wv_u, wv_v, th, tw):
    H, W = d11.shape
    i11 = np.zeros((H, W), np.float32); i12 = np.zeros((H, W), np.float32); i22 = np.zeros((H, W), np.float32)
    for i in range(H):
        for j in range(W):
            p11 = d11[i, j]; p12 = d12[i, j]; p22 = d22[i, j]
            q11 = p11; q12 = p12; q22 = p22
            if j % tw != 0:
                wu = wh_u[i, j - 1]; wv = wh_v[i, j - 1]
                p11 = p11 - (wu * wu) * i11[i, j - 1]; p12 = p12 - (wu * wv) * i12[i, j - 1]
                p22 = p22 - (wv * wv) * i22[i, j - 1]
            if i % th != 0:
                wu = wv_u[i - 1, j]; wv = wv_v[i - 1, j]
                p11 = p11 - (wu * wu) * i11[i - 1, j]; p12 = p12 - (wu * wv) * i12[i - 1, j]
                p22 = p22 - (wv * wv) * i22[i - 1, j]
            det = p11 * p22 - p12 * p12
            if not (p11 > 0 and det > np.float32(1e-6) * p11 * p22):
                p11 = q11; p12 = q12; p22 = q22
                det = p11 * p22 - p12 * p12
            inv = np.float32(1.0) / det
            i11[i, j] = p22 * inv; i12[i, j] = -p12 * inv; i22[i, j] = p11 * inv
    return i11, i12, i22


@nb.njit(cache=True)
def apply32(i11, i12, i22, wh_u, wh_v, wv_u, wv_v, r, th, tw):
    H, W = i11.shape
    t = np.zeros((H, W, 2), np.float32)
    for i in range(H):
        for j in range(W):
            su = r[i, j, 0]; sv = r[i, j, 1]
            if j % tw != 0:
                su += wh_u[i, j - 1] * t[i, j - 1, 0]; sv += wh_v[i, j - 1] * t[i, j - 1, 1]
            if i % th != 0:
                su += wv_u[i - 1, j] * t[i - 1, j, 0]; sv += wv_v[i - 1, j] * t[i - 1, j, 1]
            t[i, j, 0] = i11[i, j] * su + i12[i, j] * sv
            t[i, j, 1] = i12[i, j] * su + i22[i, j] * sv
    z = np.zeros((H, W, 2), np.float32)
    for i in range(H - 1, -1, -1):
        for j in range(W - 1, -1, -1):
            su = np.float32(0.0); sv = np.float32(0.0)
            if (j + 1) % tw != 0 and j + 1 < W:
                su += wh_u[i, j] * z[i, j + 1, 0]; sv += wh_v[i, j] * z[i, j + 1, 1]
            if (i + 1) % th != 0 and i + 1 < H:
                su += wv_u[i, j] * z[i + 1, j, 0]; sv += wv_v[i, j] * z[i + 1, j, 1]
            z[i, j, 0] = t[i, j, 0] + (i11[i, j] * su + i12[i, j] * sv)
            z[i, j, 1] = t[i, j, 1] + (i12[i, j] * su + i22[i, j] * sv)
    return z


def trunc_bf16(a):
    return (a.astype(f32).view(np.uint32) & np.uint32(0xFFFF0000)).view(f32)


def main():
    systems = make_systems()
    for alpha in (1.0, 0.0):
        s = systems[alpha]
        b = np.stack([s["bu"], s["bv"]], axis=2)
        dg = fo.operator_diag(s)
        M = block_inv(s)
        x, it = pcg(s, b, lambda r: apply_minv(M, r))
        print("alpha=%g block-Jacobi: %d it" % (alpha, it))
        for name, cvt in (("fp32", lambda a: a.astype(f32)), ("bf16-trunc edges", trunc_bf16)):
            w = [cvt(s[k]) for k in ("wuh", "wvh", "wuv", "wvv")]
            for th, tw in ((8, 32), (8, 64), (16, 32)):
                f = factor32(dg[:, :, 0].astype(f32), s["a12"].astype(f32), dg[:, :, 1].astype(f32), *w, th, tw)
                x2, it = pcg(s, b, lambda r: apply32(*f, *w, r.astype(f32), th, tw).astype(float), maxit=700)
                print("  %s tile %dx%d: %3d it |dx| %.1e" % (name, th, tw, it, np.abs(x2 - x).max()))


if __name__ == "__main__":
    main()
