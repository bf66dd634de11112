"""Diagnostic: which solve of config 3 (ba, 1920x1080) stops short of the 1e-12 true-residual target, and where.
Usage (GPU box): B200FLOW_TRACE=1 python scripts/diag_ba1080.py 2>&1 | grep -B1 NOT"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-python_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from optical_flow import estimate_flow
im1, im2, flow = synth.gray_pair(1080, 1920, seed=1, disc=True)
uv = estimate_flow(im1, im2, "ba", {"max_iters": 3})
