"""Solver micro-benchmark (tuning / ncu helper): per-iteration time and algorithmic GB/s of the persistent PCG kernels on
synthetic batches.  Usage: python scripts/pcg_bench.py [--solver 0|2] [--iters 200] [--cases B,H,W ...]"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-python_b200"))
from optical_flow import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--solver", type=int, default=4, help="4 mixed + tile-local IC(0) (128 B/px/it), 0 mixed block-Jacobi (120 B/px/it), 2 all-fp64 (228 B/px/it)")
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cases", nargs="*", default=["16,480,640", "16,384,512", "16,240,320", "16,120,160", "16,60,80",
                                                "16,30,40", "1,480,640", "4,2160,3840", "64,388,584"])
args = ap.parse_args()
ctx = _lib.default_context(0)
bytes_px = {0: 120, 2: 228, 1: 228, 4: 128}[args.solver]
for case in args.cases:
    B, H, W = [int(v) for v in case.split(",")]
    ms, it = C.c_double(0), C.c_longlong(0)
    ctx.call("b200flow_debug_pcg_bench", B, H, W, args.solver, args.iters, args.reps, 4.0, C.byref(ms), C.byref(it))
    us_it = 1e3 * ms.value / max(1, it.value)
    gbs = B * H * W * bytes_px / (us_it * 1e-6) / 1e9
    print(json.dumps({"solver": args.solver, "B": B, "H": H, "W": W, "iters": it.value, "ms_per_solve": round(ms.value, 3),
                      "us_per_iter": round(us_it, 2), "algorithmic_GBps": round(gbs, 1)}))
