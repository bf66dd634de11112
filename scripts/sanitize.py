"""compute-sanitizer target: small runs that touch every kernel of the round (TMA-tiled ROF, separable pyramid, segmented
spline prefilter at line lengths > 256, edge-sharing assembly, IC solver with reliable updates on a batch with staggered
convergence, weighted median).  Usage: compute-sanitizer --tool memcheck|racecheck python scripts/sanitize.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "optical-flow-python_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import synth
from optical_flow import estimate_flow, estimate_flow_batch
pairs = [bench.synth_pair(72, 104, 50 + k) for k in range(3)]
uv = estimate_flow_batch(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]), "classic+nl-fast")
print("classic+nl-fast batch ok", float(np.abs(uv).max()))
im1, im2, _ = synth.gray_pair(40, 300, seed=3)          # 300 columns: two prefilter segments along x
uv = estimate_flow(im1, im2, "ba", {"max_iters": 1, "gnc_iters": 2})
print("ba ok", float(np.abs(uv).max()))
uv = estimate_flow(im1[:, :299], im2[:, :299], "hs")    # odd width: scalar ROF path
print("hs ok", float(np.abs(uv).max()))
