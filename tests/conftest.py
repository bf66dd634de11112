"""Shared test plumbing.  `-m gpu` tests call the CUDA path through the C ABI (via the ctypes shim in
optical-flow-python_b200/optical_flow) and check it against the oracle / committed goldens; everything else runs on CPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "optical-flow-python_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long-running")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def stages():
    return load_golden("stages.npz")


@pytest.fixture(scope="session")
def systems():
    return load_golden("systems.npz")


@pytest.fixture(scope="session")
def crop_rgb(stages):
    return stages["rgb1"].astype(float), stages["rgb2"].astype(float)


def maxabs(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, "shape %r != %r" % (a.shape, b.shape)
    return float(np.max(np.abs(a - b))) if a.size else 0.0


def _report(what, err, tol):
    """Measured parity errors are appended to gpurun_out/parity_report.txt (scratch; summarised in profiles/)."""
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_report.txt"), "a") as f:
            f.write("%-60s max|diff| %.3e  (tol %.1e)\n" % (what, err, tol))
    except OSError:
        pass


def assert_close(got, want, tol, what=""):
    err = maxabs(got, want)
    _report(what, err, tol)
    if not err <= tol:
        d = np.abs(np.asarray(got, dtype=float) - np.asarray(want, dtype=float))
        idx = np.unravel_index(np.nanargmax(d), d.shape)
        raise AssertionError("%s: max|diff| = %.3e > %.1e at %s (got %r, want %r); %d entries over tol" % (
            what, err, tol, idx, np.asarray(got)[idx], np.asarray(want)[idx], int((d > tol).sum())))
