"""Diagnostic (round 2): where does the CUDA flow differ from the reference on the Middlebury sequences, and is it the
solver tolerance (discrete weighted-median selections amplify a 1e-9 px solve error) or the weighted median itself?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-python_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from optical_flow import estimate_flow
from optical_flow.utils.weighted_median import denoise_color_weighted_medfilt2
from optical_flow.utils.occlusion import detect_occlusion
import flow_oracle as fo

seqs = sys.argv[1:] or ["Grove3", "Urban3", "Hydrangea"]
for seq in seqs:
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg_mb_%s.npz" % seq))
    im1, im2 = g["im1"].astype(float), g["im2"].astype(float)
    ref = g["uv"].astype(np.float64)
    for rtol in (1e-10, 1e-12, 1e-13):
        t0 = time.time()
        uv = estimate_flow(im1, im2, "classic+nl-fast", {"exact_rtol": rtol})
        d = np.abs(uv - ref).max(axis=2)
        idx = np.unravel_index(np.argmax(d), d.shape)
        print("%s rtol %g: max %.3e at %s, >1e-3: %d px, >1e-4: %d, >1e-5: %d, median %.2e  (%.2fs)" % (
            seq, rtol, d.max(), idx, (d > 1e-3).sum(), (d > 1e-4).sum(), (d > 1e-5).sum(), np.median(d), time.time() - t0), flush=True)
    # weighted median in isolation on this sequence's Lab guide: GPU vs the oracle's restatement, bit-exact?
    lab = fo.rgb2lab(im1)
    for j in range(3):
        lab[:, :, j] = fo.scale_image(lab[:, :, j], 0, 255)
    H, W = ref.shape[:2]
    sl = (slice(H // 2 - 60, H // 2 + 60), slice(W // 2 - 80, W // 2 + 80))
    uvc = np.ascontiguousarray(ref[sl]); labc = np.ascontiguousarray(lab[sl])
    gray = np.stack([fo.rgb2gray(im1)[sl], fo.rgb2gray(im2)[sl]], axis=2)
    occ = fo.detect_occlusion(uvc, gray)
    a = denoise_color_weighted_medfilt2(uvc, labc, occ, 7, [5, 5], 7, False)
    b = fo.weighted_median_filter(uvc, labc, occ, 7, 7.0)
    nd = (a != b).any(axis=2).sum()
    print("%s weighted median GPU vs oracle on the centre 120x160: %d differing pixels, max %.3e" % (seq, nd, np.abs(a - b).max()), flush=True)
