"""Transcripts of what the UNMODIFIED reference prints under display=True (classic_nl.py:141-152,255-256,186-198;
ba.py:100-114,189-190,132-133; hs.py:80-81,123-124) on the 64 x 80 RubberWhale crop of gen_golden.py -> display.json.
Run in the build container (needs /root/reference):  python tests/golden/gen_golden_display.py"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_stubs"))     # matplotlib stand-in (the reference's viz module imports it)
sys.path.insert(1, "/root/reference")
sys.dont_write_bytecode = True
import optical_flow  # noqa: E402  (the REFERENCE package)
from optical_flow import estimate_flow  # noqa: E402

assert optical_flow.__file__.startswith("/root/reference"), optical_flow.__file__

CROP = (slice(150, 214), slice(230, 310))
CASES = [("classic+nl-fast", None), ("hs", None), ("hs-brightness", {"display": True}),
         ("ba", {"display": True, "max_iters": 2, "gnc_iters": 2}),
         ("classic++", {"display": True, "max_iters": 2, "max_linear": 2}),
         # display off: HS is silent, the GNC drivers still print their "GNC stage k finished" lines (ba.py:132-133,
         # classic_nl.py:186-198)
         ("hs-brightness", None), ("ba", {"max_iters": 1, "gnc_iters": 2}), ("classic++", {"max_iters": 1}),
         ("classic+nl-fast", {"display": False, "max_iters": 1})]


def main():
    d = np.load(os.path.join(HERE, "rubberwhale_10_11.npz"))
    c1, c2 = d["im1"][CROP].copy(), d["im2"][CROP].copy()
    out = {}
    for name, params in CASES:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            estimate_flow(c1, c2, name, params)
        key = name + ("" if params is None else "|" + json.dumps(params, sort_keys=True))
        out[key] = buf.getvalue().splitlines()
        print(key, len(out[key]), "lines", file=sys.stderr)
    with open(os.path.join(HERE, "display.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
