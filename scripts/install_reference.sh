#!/bin/bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to the GPU box), as the
# base contract describes.  Outcome in this image: dependency resolution fails offline (numpy>=1.21 is not in the wheelhouse
# index), so --no-deps; matplotlib is absent, a 2-file stub (tests/golden/_stubs) lets `import optical_flow` succeed.
# Also copies the reference's own test-suite and the one sequence it loads (RubberWhale) next to it, so that
# tests/test_gpu_reference_suite.py can run the reference's 82 tests against the B200 drop-in on the GPU box.
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/refcopy baseline/_ref
cp -r /root/reference /tmp/refcopy
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/refcopy
# the suite goes into a SEPARATE directory (baseline/_ref_suite): pytest puts a test package's parent on sys.path, and the
# parent must not contain the reference's optical_flow package when the suite is pointed at the drop-in
rm -rf baseline/_ref/tests baseline/_ref_suite
S=baseline/_ref_suite
mkdir -p $S/tests $S/data/other-data/RubberWhale $S/data/other-gt-flow/RubberWhale
cp /root/reference/tests/*.py $S/tests/
cp /root/reference/data/other-data/RubberWhale/frame10.png /root/reference/data/other-data/RubberWhale/frame11.png $S/data/other-data/RubberWhale/
cp /root/reference/data/other-gt-flow/RubberWhale/flow10.flo $S/data/other-gt-flow/RubberWhale/
ls baseline/_ref $S
