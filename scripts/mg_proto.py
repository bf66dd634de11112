"""CPU prototype (design evidence, not product): iteration counts of PCG on real Classic+NL systems with
(a) the 2x2 block-Jacobi preconditioner of round 1 and (b) multilevel V-cycle preconditioners, to pick the
scheme that is worth writing in CUDA.  Systems: RubberWhale 584x388 full resolution, flow = the reference's
final flow (tests/golden/rubberwhale_full.npz), alpha = 1 (quadratic stage) and alpha = 0 (gen. Charbonnier)."""
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import flow_oracle as fo  # noqa: E402


def make_systems(scale=1.0):
    g = np.load(os.path.join(ROOT, "tests/golden/rubberwhale_10_11.npz"))
    uv = np.load(os.path.join(ROOT, "tests/golden/rubberwhale_full.npz"))["uv"] * scale
    im1, im2 = g["im1"].astype(float), g["im2"].astype(float)
    images = np.stack([fo.rgb2gray(im1), fo.rgb2gray(im2)], axis=2)
    p = fo.preset("classic+nl-fast")
    pre = fo._preprocess(p, images)
    spec = fo._spec(p)
    It, Ix, Iy = fo.partial_deriv(pre, uv, p["interp"], p["deriv_filter"], p["blend"])
    out = {}
    for alpha in (1.0, 0.0):
        out[alpha] = fo.assemble(uv, np.zeros_like(uv), It, Ix, Iy, spec, alpha)
    return out


KEYS = ("a11", "a12", "a22", "wuh", "wuv", "wvh", "wvv")


def pad2(a):
    H, W = a.shape
    return np.pad(a, ((0, H % 2), (0, W % 2)))


def coarsen(s, edge_scale):
    """2x2 aggregation.  Data block: sum over the aggregate.  Edges: sum of the two fine edges that cross the
    aggregate boundary (exact Galerkin for piecewise-constant P), times edge_scale (0.5 = rediscretisation)."""
    c = {}
    for k in ("a11", "a12", "a22"):
        a = pad2(s[k])
        c[k] = a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2]
    for k in ("wuh", "wvh"):
        a = pad2(s[k])
        c[k] = (a[0::2, 1::2] + a[1::2, 1::2]) * edge_scale
        c[k][:, -1] = 0.0
    for k in ("wuv", "wvv"):
        a = pad2(s[k])
        c[k] = (a[1::2, 0::2] + a[1::2, 1::2]) * edge_scale
        c[k][-1, :] = 0.0
    return c


def restrict_sum(r):
    out = []
    for ch in range(2):
        a = pad2(r[:, :, ch])
        out.append(a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2])
    return np.stack(out, axis=2)


def prolong_const(e, H, W):
    return np.repeat(np.repeat(e, 2, axis=0), 2, axis=1)[:H, :W]


def prolong_bilinear(e, H, W):
    """cell-centred bilinear: fine cell (2I+a, 2J+b) takes 9/16, 3/16, 3/16, 1/16 of the 4 nearest coarse cells
    (clamped at the border)."""
    Hc, Wc = e.shape[:2]
    ep = np.pad(e, ((1, 1), (1, 1), (0, 0)), mode="edge")
    out = np.zeros((2 * Hc, 2 * Wc, 2))
    for a in (0, 1):
        for b in (0, 1):
            di = -1 if a == 0 else 1
            dj = -1 if b == 0 else 1
            c = ep[1:-1, 1:-1]
            ci = ep[1 + di:Hc + 1 + di, 1:-1]
            cj = ep[1:-1, 1 + dj:Wc + 1 + dj]
            cij = ep[1 + di:Hc + 1 + di, 1 + dj:Wc + 1 + dj]
            out[a::2, b::2] = (9 * c + 3 * ci + 3 * cj + cij) / 16.0
    return out[:H, :W]


def restrict_bilinear_T(r):
    """transpose of prolong_bilinear (computed by explicit adjoint accumulation)"""
    H, W = r.shape[:2]
    Hc, Wc = (H + 1) // 2, (W + 1) // 2
    acc = np.zeros((Hc + 2, Wc + 2, 2))
    rp = np.zeros((2 * Hc, 2 * Wc, 2))
    rp[:H, :W] = r
    for a in (0, 1):
        for b in (0, 1):
            di = -1 if a == 0 else 1
            dj = -1 if b == 0 else 1
            f = rp[a::2, b::2]
            acc[1:-1, 1:-1] += 9 * f / 16
            acc[1 + di:Hc + 1 + di, 1:-1] += 3 * f / 16
            acc[1:-1, 1 + dj:Wc + 1 + dj] += 3 * f / 16
            acc[1 + di:Hc + 1 + di, 1 + dj:Wc + 1 + dj] += f / 16
    # fold the clamped border back (edge padding adjoint)
    acc[1, :] += acc[0, :]
    acc[-2, :] += acc[-1, :]
    acc[:, 1] += acc[:, 0]
    acc[:, -2] += acc[:, -1]
    return acc[1:-1, 1:-1]


def block_inv(s):
    dg = fo.operator_diag(s)
    d11, d22, a12 = dg[:, :, 0], dg[:, :, 1], s["a12"]
    det = d11 * d22 - a12 * a12
    ok = det > 1e-300
    det = np.where(ok, det, 1.0)
    m11 = np.where(ok, d22 / det, np.where(np.abs(d11) > 1e-12, 1 / np.where(d11 == 0, 1, d11), 0))
    m22 = np.where(ok, d11 / det, np.where(np.abs(d22) > 1e-12, 1 / np.where(d22 == 0, 1, d22), 0))
    m12 = np.where(ok, -a12 / det, 0)
    return m11, m12, m22


def apply_minv(M, r):
    m11, m12, m22 = M
    return np.stack([m11 * r[:, :, 0] + m12 * r[:, :, 1], m12 * r[:, :, 0] + m22 * r[:, :, 1]], axis=2)


class MG:
    def __init__(self, s, nlev, edge_scale=0.5, omega=0.8, nu=1, interp="const", coarse_sweeps=8, smoother="jacobi",
                 overcorrect=1.0):
        self.lv = [s]
        for _ in range(nlev - 1):
            if min(self.lv[-1]["a11"].shape) <= 4:
                break
            self.lv.append(coarsen(self.lv[-1], edge_scale))
        self.M = [block_inv(l) for l in self.lv]
        self.omega, self.nu, self.interp, self.cs, self.smoother, self.oc = omega, nu, interp, coarse_sweeps, smoother, overcorrect
        self.work = 0.0   # fine-level matvec equivalents

    def smooth(self, l, x, r, reverse=False):
        s, M = self.lv[l], self.M[l]
        n = s["a11"].size / self.lv[0]["a11"].size
        if self.smoother == "jacobi":
            if x is None:
                self.work += 0.5 * n
                return self.omega * apply_minv(M, r)
            self.work += 1.2 * n
            return x + self.omega * apply_minv(M, r - fo.apply_operator(s, x))
        # red-black block Gauss-Seidel (colour order reversed on the way up => symmetric V-cycle)
        H, W = s["a11"].shape
        ii, jj = np.indices((H, W))
        if x is None:
            x = np.zeros_like(r)
        for colour in ((1, 0) if reverse else (0, 1)):
            m = ((ii + jj) % 2 == colour)[:, :, None]
            x = np.where(m, x + apply_minv(M, r - fo.apply_operator(s, x)), x)
            self.work += 1.2 * n
        return x

    def vcycle(self, l, r):
        s = self.lv[l]
        if l == len(self.lv) - 1:
            x = None
            for _ in range(self.cs):
                x = self.smooth(l, x, r)
            for _ in range(self.cs):
                x = self.smooth(l, x, r, True)
            return x
        x = None
        for _ in range(self.nu):
            x = self.smooth(l, x, r)
        res = r - fo.apply_operator(s, x)
        self.work += 1.2 * s["a11"].size / self.lv[0]["a11"].size
        H, W = s["a11"].shape
        if self.interp == "const":
            ec = self.vcycle(l + 1, restrict_sum(res))
            x = x + self.oc * prolong_const(ec, H, W)
        else:
            ec = self.vcycle(l + 1, restrict_bilinear_T(res))
            x = x + self.oc * prolong_bilinear(ec, H, W)
        for _ in range(self.nu):
            x = self.smooth(l, x, r, True)
        return x

    def __call__(self, r):
        return self.vcycle(0, r)


def pcg(s, b, prec, rtol=1e-10, maxit=2000):
    x = np.zeros_like(b)
    r = b.copy()
    z = prec(r)
    p = z.copy()
    rz = float((r * z).sum())
    bb = float((b * b).sum())
    for k in range(maxit):
        Ap = fo.apply_operator(s, p)
        a = rz / float((p * Ap).sum())
        x += a * p
        r -= a * Ap
        rr = float((r * r).sum())
        if rr <= rtol * rtol * bb:
            return x, k + 1
        z = prec(r)
        rz2 = float((r * z).sum())
        p = z + (rz2 / rz) * p
        rz = rz2
    return x, maxit


def main():
    systems = make_systems()
    for alpha, s in systems.items():
        b = np.stack([s["bu"], s["bv"]], axis=2)
        print("alpha=%g  %dx%d  edge w range [%.3g, %.3g]  a11 range [%.3g, %.3g]" % (
            alpha, *s["a11"].shape, s["wuh"][:, :-1].min(), s["wuh"].max(), s["a11"].min(), s["a11"].max()))
        M = block_inv(s)
        t = time.time()
        x, it = pcg(s, b, lambda r: apply_minv(M, r))
        print("  block-Jacobi PCG: %4d iterations  (%.1fs)  cost %.0f matvec-equivalents (228 B/it = 2.6)" % (it, time.time() - t, it * 2.6))
        xref = x
        for kw in (
            dict(interp="const", edge_scale=0.5, omega=0.8, nu=1),
            dict(interp="const", edge_scale=0.5, omega=0.8, nu=2),
            dict(interp="const", edge_scale=1.0, omega=0.8, nu=1, overcorrect=1.0),
            dict(interp="const", edge_scale=1.0, omega=0.8, nu=1, overcorrect=1.8),
            dict(interp="bilinear", edge_scale=0.5, omega=0.8, nu=1),
            dict(interp="bilinear", edge_scale=0.5, omega=0.8, nu=2),
            dict(interp="const", edge_scale=0.5, nu=1, smoother="rbgs"),
            dict(interp="bilinear", edge_scale=0.5, nu=1, smoother="rbgs"),
        ):
            mg = MG(s, 7, **kw)
            t = time.time()
            x, it = pcg(s, b, mg, maxit=300)
            err = np.abs(x - xref).max()
            print("  MG-PCG %-70s: %3d it, levels %d, %.1f matvec-eq/it, total %.0f  |x-xref| %.1e (%.1fs)" % (
                kw, it, len(mg.lv), mg.work / max(it, 1), mg.work + it * 2.6, err, time.time() - t))


if __name__ == "__main__":
    main()
