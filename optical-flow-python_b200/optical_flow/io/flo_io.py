"""Middlebury .flo files (host I/O edge of the path; reference: io/flo_io.py).

Layout: float32 tag 202021.25, int32 width, int32 height, then height*width*(u,v) float32, row-major."""
import os

import numpy as np

TAG_FLOAT = 202021.25


def read_flo(filename):
    with open(filename, 'rb') as f:
        head = np.frombuffer(f.read(12), dtype=[('tag', '<f4'), ('w', '<i4'), ('h', '<i4')])
        if head.size != 1 or head['tag'][0] != np.float32(TAG_FLOAT):
            tag = head['tag'][0] if head.size else None
            raise ValueError(f'Invalid .flo file tag: {tag} (expected {TAG_FLOAT})')
        w, h = int(head['w'][0]), int(head['h'][0])
        data = np.fromfile(f, '<f4', count=2 * w * h)
    return data.reshape(h, w, 2)


def write_flo(flow, filename):
    flow = np.asarray(flow, dtype=np.float32)
    if flow.ndim != 3 or flow.shape[2] != 2:
        raise ValueError(f"Flow must be (H, W, 2) array, got shape {flow.shape}")
    h, w = flow.shape[:2]
    with open(filename, 'wb') as f:
        f.write(np.float32(TAG_FLOAT).tobytes())
        f.write(np.array([w, h], dtype='<i4').tobytes())
        f.write(np.ascontiguousarray(flow, dtype='<f4').tobytes())


def flo_bytes(flow):
    """The byte image of write_flo(flow, ...) built on the device (b200flow_flow_to_flo): (H, W, 2) -> bytes, or
    (B, H, W, 2) -> list of bytes.  Lets a batch benchmarking loop export float32 .flo payloads without first bringing
    the float64 fields to the host."""
    from optical_flow import _lib
    flow = _lib.f64(flow)
    single = flow.ndim == 3
    if single:
        flow = flow[None]
    if flow.ndim != 4 or flow.shape[3] != 2:
        raise ValueError(f"Flow must be (H, W, 2) array, got shape {flow.shape}")
    B, H, W = flow.shape[:3]
    out = np.empty((B, 12 + 8 * H * W), dtype=np.uint8)
    _lib.default_context().call("b200flow_flow_to_flo", _lib.ptr(flow), B, H, W, _lib.ptr(out))
    res = [out[b].tobytes() for b in range(B)]
    return res[0] if single else res


def read_flow_file(seq_name, i_seq, data_dir=None):
    """(im1, im2, tu, tv) of a Middlebury sequence under data_dir/other-data and data_dir/other-gt-flow."""
    from PIL import Image
    if data_dir is None:
        data_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'data')
    base = os.path.join(data_dir, 'other-data', seq_name)
    im1 = np.array(Image.open(os.path.join(base, f'frame{i_seq:02d}.png'))).astype(np.float64)
    im2 = np.array(Image.open(os.path.join(base, f'frame{i_seq + 1:02d}.png'))).astype(np.float64)
    gt = os.path.join(data_dir, 'other-gt-flow', seq_name, f'flow{i_seq:02d}.flo')
    if os.path.exists(gt):
        fl = read_flo(gt)
        return im1, im2, fl[:, :, 0], fl[:, :, 1]
    return im1, im2, None, None
