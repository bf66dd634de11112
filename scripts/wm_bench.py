"""Micro-benchmark of the weighted-median kernel alone on a realistic input (640x480 synthetic pair, noisy flow)."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'optical-flow-python_b200')); sys.path.insert(0, ROOT)
import bench
from optical_flow.utils.weighted_median import denoise_color_weighted_medfilt2
from optical_flow.utils.occlusion import detect_occlusion
from optical_flow.interface import _rgb2lab, _rgb2gray
im1, im2, flow = bench.synth_pair(480, 640, 3)
rng = np.random.default_rng(0)
uv = flow + 0.05 * rng.standard_normal(flow.shape)
lab = _rgb2lab(im1.astype(float), True)
images = np.stack([_rgb2gray(im1.astype(float)), _rgb2gray(im2.astype(float))], 2)
occ = detect_occlusion(uv, images)
for i in range(3):
    t0 = time.time(); out = denoise_color_weighted_medfilt2(uv, lab, occ, 7, [5, 5], 7); dt = time.time() - t0
print('wm 640x480 wall %.2f ms' % (dt * 1e3), out.shape)
