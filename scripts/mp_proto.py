"""CPU prototype (design evidence, not product): mixed-precision PCG with reliable updates -- fp32 vectors and
coefficients inside the iteration, fp64 solution accumulation and periodic fp64 true-residual replacement."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mg_proto import make_systems, pcg, block_inv, apply_minv, fo  # noqa: E402


def mp_pcg(s, b, rtol=1e-10, delta=0.1, maxit=3000, prec_dtype=np.float32):
    f32 = prec_dtype
    s32 = {k: s[k].astype(f32) for k in ("a11", "a12", "a22", "wuh", "wuv", "wvh", "wvv")}
    M = tuple(m.astype(f32) for m in block_inv(s))
    x = np.zeros_like(b)
    r = b.copy()
    bb = float((b * b).sum())
    rs = r.astype(f32)
    z = apply_minv(M, rs)
    p = z.copy()
    rz = float((rs.astype(np.float64) * z).sum())
    y = np.zeros_like(rs)
    maxr = np.sqrt(bb)
    nrel = 0
    for k in range(maxit):
        Ap = fo.apply_operator(s32, p)
        pap = float((p.astype(np.float64) * Ap).sum())
        a = rz / pap
        y += f32(a) * p
        rs -= f32(a) * Ap
        rr = float((rs.astype(np.float64) ** 2).sum())
        if rr < delta * delta * maxr * maxr or rr <= rtol * rtol * bb:
            x += y
            y[:] = 0
            r = b - fo.apply_operator(s, x)
            rr = float((r * r).sum())
            nrel += 1
            if rr <= rtol * rtol * bb:
                return x, k + 1, nrel
            rs = r.astype(f32)
            maxr = np.sqrt(rr)
        z = apply_minv(M, rs)
        rz2 = float((rs.astype(np.float64) * z).sum())
        p = z + f32(rz2 / rz) * p
        rz = rz2
    return x + y, maxit, nrel


if __name__ == "__main__":
    systems = make_systems()
    for alpha in (1.0, 0.0):
        s = systems[alpha]
        b = np.stack([s["bu"], s["bv"]], axis=2)
        M = block_inv(s)
        x, it = pcg(s, b, lambda r: apply_minv(M, r))
        print("alpha=%g fp64 block-Jacobi: %d it" % (alpha, it))
        for delta in (0.3, 0.1, 0.03, 0.01, 1e-3):
            x2, it2, nrel = mp_pcg(s, b, delta=delta)
            res = np.sqrt(((b - fo.apply_operator(s, x2)) ** 2).sum() / (b * b).sum())
            print("  mixed fp32/fp64 delta=%g: %d it, %d reliable updates, |dx| %.1e, true relres %.1e" % (
                delta, it2, nrel, np.abs(x2 - x).max(), res))
