"""display=True: the per-stage / per-level / per-iteration lines of the reference's drivers (classic_nl.py:141-152,255-256;
ba.py:100-114,189-190; hs.py:80-81,123-124) against transcripts of the unmodified reference on the 64 x 80 RubberWhale crop
(tests/golden/display.json, written by tests/golden/gen_golden_display.py).  Same lines in the same order; the printed
norms agree to the flow's own parity (they are printed with six decimals), the wall-clock part of the "finished" lines is
not compared.  With display off the GNC drivers still print their "GNC stage k finished" lines, as the reference does
(ba.py:132-133, classic_nl.py:186-198), and Horn-Schunck prints nothing."""
import contextlib
import io
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu

CROP = (slice(150, 214), slice(230, 310))
with open(os.path.join(GOLDEN, "display.json")) as _f:
    TRANSCRIPTS = json.load(_f)
NUM = re.compile(r"^(.*\((?:delta|norm): )([0-9.eE+-]+)\)$")


@pytest.mark.parametrize("key", sorted(TRANSCRIPTS))
def test_display_lines_match_the_reference(key):
    from optical_flow import estimate_flow
    name, _, pj = key.partition("|")
    params = json.loads(pj) if pj else None
    d = load_golden("rubberwhale_10_11.npz")
    c1, c2 = d["im1"][CROP].copy(), d["im2"][CROP].copy()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        estimate_flow(c1, c2, name, params)
    got, want = buf.getvalue().splitlines(), TRANSCRIPTS[key]
    assert len(got) == len(want), "line count %d vs %d\n%s" % (len(got), len(want), "\n".join(got))
    worst = 0.0
    for g, w in zip(got, want):
        mg, mw = NUM.match(g), NUM.match(w)
        if mw:
            assert mg and mg.group(1) == mw.group(1), (g, w)
            a, b = float(mg.group(2)), float(mw.group(2))
            worst = max(worst, abs(a - b))
            assert abs(a - b) <= 2e-4 + 1e-5 * abs(b), (g, w)
        elif " finished, " in w:
            assert g.split(" finished, ")[0] == w.split(" finished, ")[0] and g.endswith("minutes passed"), (g, w)
        else:
            assert g == w
    print("display %s: %d lines, worst |norm difference| %.1e" % (key, len(want), worst))
