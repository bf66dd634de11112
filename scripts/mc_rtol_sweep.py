"""Tuning / evidence helper: two-channel classic+nl-fast (tests/golden multichannel fixture, ill-conditioned in the
reference itself) final-flow distance to the reference as the exact solve is tightened, per preconditioner."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-python_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from optical_flow import estimate_flow
mc = np.load(os.path.join(ROOT, "tests", "golden", "multichannel.npz"))
for prec in ("mixed", "mixed-jacobi", "fp64"):
    for rtol in (1e-10, 1e-12, 1e-13, 1e-14):
        try:
            uv = estimate_flow(mc["c1"], mc["c2"], "classic+nl-fast", {"exact_rtol": rtol, "solver_precision": prec})
            d = np.abs(uv - mc["e2e_classic+nl-fast"])
            print("%-13s rtol %.0e: max %.3e px, %d entries > 1e-3" % (prec, rtol, d.max(), int((d > 1e-3).sum())), flush=True)
        except Exception as e:  # noqa
            print(prec, rtol, "failed:", str(e)[:200], flush=True)
