#!/usr/bin/env python
"""Weighted-median tie / near-tie fuzz fixtures, produced by RUNNING THE UNMODIFIED REFERENCE's
denoise_color_weighted_medfilt2 (weighted_median.py:24-112) on inputs built to sit on its decision boundary
`cumw[k] >= total / 2` (weighted_median.py:15-21): flows quantised to 1/8 px (many equal values), weights from small sets
(exact powers of two: ties that are exact in any summation order; 0.1 / 0.2 / 0.3 / 0.7: ties that depend on the ROUNDING of
the reference's sequential cumsum in sorted order), constant or two-level colour guides.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden_wmfuzz.py      -> tests/golden/wmed_fuzz.npz
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(1, "/root/reference")
sys.dont_write_bytecode = True
import numpy as np  # noqa: E402
import optical_flow  # noqa: E402
from optical_flow.utils.weighted_median import denoise_color_weighted_medfilt2  # noqa: E402

assert optical_flow.__file__.startswith("/root/reference")


def cases():
    rng = np.random.default_rng(2024)
    H, W = 36, 44
    out = []
    for ci, (wset, hsz, colour) in enumerate([
            ([1.0, 0.5, 0.25, 0.125, 2.0 ** -6], 7, "const"),
            ([0.1, 0.2, 0.3, 0.7], 7, "const"),
            ([0.1, 0.2, 0.3, 0.7], 2, "const"),
            ([1.0], 7, "two"),
            ([0.3, 0.6, 0.9], 3, "two"),
            ([1.0, 1e-12], 7, "const"),            # half of the samples on the 1e-10 floor
            ([0.1, 0.2, 0.3, 0.7], 7, "three")]):
        uv = np.round(rng.normal(0.0, 0.4, (H, W, 2)) * 8) / 8          # 1/8-px grid, ~10 distinct values per window
        if ci % 2:
            uv[:, :, 1] = np.round(rng.normal(0.0, 0.15, (H, W)) * 8) / 8   # v: 3-4 distinct values only
        occ = rng.choice(np.asarray(wset), size=(H, W))
        if colour == "const":
            col = np.full((H, W, 3), 100.0)
        elif colour == "two":
            col = np.full((H, W, 3), 100.0)
            col[:, W // 2:, 0] += 7.0                                      # exp(-49 / 98) = exp(-0.5)
        else:
            col = np.full((H, W, 3), 100.0)
            col[:, :, 0] += 7.0 * rng.integers(0, 3, (H, W))              # levels 0, 7, 14 -> exp(-0.5 k^2)
        out.append((uv, col, occ, hsz))
    return out


def main():
    res = {}
    for i, (uv, col, occ, hsz) in enumerate(cases()):
        got = denoise_color_weighted_medfilt2(uv, col, occ, hsz, [5, 5], 7, False)
        res["c%d_uv" % i], res["c%d_col" % i], res["c%d_occ" % i] = uv, col, occ
        res["c%d_hsz" % i], res["c%d_out" % i] = hsz, got
        print("case %d: hsz %d, %d of %d outputs differ from the input" % (i, hsz, int((got != uv).any(axis=2).sum()), uv.shape[0] * uv.shape[1]))
    res["ncases"] = len(cases())
    np.savez_compressed(os.path.join(HERE, "wmed_fuzz.npz"), **res)
    print("wrote wmed_fuzz.npz")


if __name__ == "__main__":
    main()
