"""Evaluation / export edges on the device against the reference's own outputs
(tests/golden/eval.npz): AAE / std / AEPE with the unknown-flow mask and border crop, Middlebury colour coding (uint8:
bit-exact), and the .flo byte image (bit-exact)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ev():
    return load_golden("eval.npz")


@pytest.mark.parametrize("border", [0, 5])
def test_flow_error_batch(ev, border):
    from optical_flow import flow_error_batch, flow_angular_error
    want = ev["metrics_b%d" % border]
    uv = np.stack([ev["est"], ev["est"] * 0.5])
    gt = np.stack([ev["gt"], ev["gt"]])
    got = flow_error_batch(uv, gt, border)
    assert got.shape == (2, 4)
    np.testing.assert_allclose(got[0, :3], want, rtol=1e-10, atol=0)       # floating point: 1e-10 relative
    sl = slice(border, -border) if border else slice(None)
    known = (np.abs(ev["gt"][sl, sl]) < 1e9).all(axis=2)
    assert got[0, 3] == known.sum()
    other = flow_angular_error(ev["gt"][:, :, 0], ev["gt"][:, :, 1], 0.5 * ev["est"][:, :, 0], 0.5 * ev["est"][:, :, 1], border)
    np.testing.assert_allclose(got[1, :3], other, rtol=1e-10, atol=0)


def test_flow_to_color_bit_exact(ev):
    from optical_flow import flow_to_color
    np.testing.assert_array_equal(flow_to_color(ev["est"]), ev["color_auto"])
    np.testing.assert_array_equal(flow_to_color(ev["big"], max_flow=2.0), ev["color_max2"])
    np.testing.assert_array_equal(flow_to_color(ev["big"]), ev["color_auto_unknown"])
    both = flow_to_color(np.stack([ev["est"], ev["big"]]))
    np.testing.assert_array_equal(both[0], ev["color_auto"])
    np.testing.assert_array_equal(both[1], ev["color_auto_unknown"])


def test_flo_bytes_bit_exact(ev, tmp_path):
    from optical_flow import flo_bytes, read_flo
    b = flo_bytes(ev["est"])
    assert b == ev["flo_bytes"].tobytes()
    p = tmp_path / "x.flo"
    p.write_bytes(b)
    np.testing.assert_array_equal(read_flo(str(p)), ev["est"].astype(np.float32))
    assert flo_bytes(np.stack([ev["est"], ev["est"]]))[1] == b
