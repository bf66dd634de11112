"""Per-solve trace of BASELINE config 5 (classic++ 3840x2160, one pair): B200FLOW_TRACE=1 python scripts/trace_4k.py 2> trace.txt"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-python_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from optical_flow import _lib, estimate_flow
_lib.default_context(0).set_timing(True)
im1, im2, flow = synth.gray_pair(2160, 3840, seed=2)
estimate_flow(im1, im2, "classic++")
estimate_flow(im1, im2, "classic++")
print("=== traced run", file=sys.stderr)
estimate_flow(im1, im2, "classic++")
