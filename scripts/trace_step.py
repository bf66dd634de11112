"""Debug helper: one traced classic+nl-fast step on the bench workload (B200FLOW_TRACE=1 prints per-solve statistics)."""
import os, sys
os.environ["B200FLOW_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "optical-flow-python_b200"))
import numpy as np
import bench
from optical_flow import _lib, estimate_flow_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "mixed"
i1, i2, _ = bench.make_batch(B, 3)
ctx = _lib.default_context(0)
os.environ.pop("B200FLOW_TRACE")
estimate_flow_batch(i1, i2, "classic+nl-fast", params={"solver_precision": prec})
os.environ["B200FLOW_TRACE"] = "1"
ctx.set_timing(True)
uv, st = estimate_flow_batch(i1, i2, "classic+nl-fast", params={"solver_precision": prec}, return_stats=True)
print(st)
