"""The reference's OWN test-suite (/root/reference/tests, 82 tests) run against the B200 drop-in: the drop-in package is
first on sys.path, so every `from optical_flow ... import` of those tests binds to optical-flow-python_b200/optical_flow and
every numerical call goes through libb200flow.so.  The suite itself is not part of this repository: scripts/install_reference.sh
copies it (with the unmodified reference and the RubberWhale frames it loads) into baseline/_ref (package) and baseline/_ref_suite (tests + frames), which are git-ignored but
travels to the GPU box.  Skipped when that directory is absent."""
import os
import subprocess
import sys

import pytest

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu
REF_TESTS = os.path.join(ROOT, "baseline", "_ref_suite", "tests")


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="baseline/_ref_suite/tests not installed (scripts/install_reference.sh)")
def test_reference_suite_passes_against_the_dropin():
    env = dict(os.environ)
    env["PYTHONPATH"] = PKG                      # the drop-in, NOT baseline/_ref: the tests must not see the reference package
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    probe = subprocess.run([sys.executable, "-c", "import optical_flow, sys; sys.stdout.write(optical_flow.__file__)"],
                           env=env, capture_output=True, text=True, cwd=REF_TESTS)
    assert os.path.realpath(probe.stdout).startswith(os.path.realpath(PKG)), (probe.stdout, probe.stderr)
    r = subprocess.run([sys.executable, "-m", "pytest", REF_TESTS, "-q", "-p", "no:cacheprovider", "--rootdir", REF_TESTS],
                       env=env, capture_output=True, text=True, cwd=REF_TESTS, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    try:
        with open(os.path.join(ROOT, "gpurun_out", "reference_suite.log"), "w") as f:
            f.write(r.stdout + r.stderr)
    except OSError:
        pass
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and " failed" not in r.stdout, tail
