"""High-level entry point (reference: optical_flow/interface.py:11-141).

estimate_flow keeps the reference's signature and semantics.  estimate_flow_batch / estimate_flow_sharded are
additions: B same-size pairs per device call, and pairs dealt round-robin over the ranks of a torch.distributed
job (one process per GPU, no data-path collective -- frame pairs are independent)."""
import numpy as np

from optical_flow import _lib
from optical_flow.methods.config import load_of_method


def _rgb2gray(im):
    """double(rgb2gray(uint8(im))) with round-half-up at both steps (interface.py:74-88), on the device."""
    im = _lib.f64(im)
    if im.ndim == 2:
        return im
    H, W = im.shape[:2]
    out = np.empty((H, W))
    _lib.default_context().call("b200flow_rgb2gray", _lib.ptr(np.ascontiguousarray(im[:, :, :3])), H, W, _lib.ptr(out))
    return out


def _rgb2lab(im, scale_channels=False):
    """BT.709 / D65 CIE Lab (interface.py:91-141); scale_channels additionally maps each channel to [0,255]."""
    im = _lib.f64(im)
    H, W = im.shape[:2]
    out = np.empty((H, W, 3))
    _lib.default_context().call("b200flow_rgb2lab", _lib.ptr(np.ascontiguousarray(im[:, :, :3])), H, W,
                                int(bool(scale_channels)), _lib.ptr(out))
    return out


def _prepare(im1, im2, method, params):
    im1 = np.asarray(im1, dtype=float)
    im2 = np.asarray(im2, dtype=float)
    ope = load_of_method(method)
    if params is not None:
        ope.parse_input_parameter(params)
    return im1, im2, ope


def estimate_flow(im1, im2, method='classic+nl-fast', params=None):
    """Optical flow (H, W, 2) between two images ((H, W) gray or (H, W, 3) RGB, float or uint8)."""
    im1, im2, ope = _prepare(im1, im2, method, params)
    if im1.ndim == 3 and im1.shape[2] >= 3:
        ope.images = np.stack([_rgb2gray(im1), _rgb2gray(im2)], axis=2)
    else:
        ope.images = np.stack([im1, im2], axis=2) if im1.ndim == 2 else np.concatenate([im1, im2], axis=2)
    if ope.color_images is not None:
        if im1.ndim == 3 and im1.shape[2] >= 3:
            ope.color_images = _rgb2lab(im1, scale_channels=True)
        else:
            ope.color_images = im1.copy()
    H, W = im1.shape[:2]
    return ope.compute_flow(np.zeros((H, W, 2)))


def estimate_flow_batch(ims1, ims2, method='classic+nl-fast', params=None, device=None, return_stats=False):
    """Flow for B same-size uint8 RGB pairs in ONE device call: ims1, ims2 (B, H, W, 3) uint8 -> (B, H, W, 2).
    Colour conversion (gray, Lab) runs on the device too (b200flow_estimate_rgb8)."""
    ims1 = np.ascontiguousarray(ims1)
    ims2 = np.ascontiguousarray(ims2)
    if ims1.dtype != np.uint8 or ims1.ndim != 4 or ims1.shape[3] != 3 or ims1.shape != ims2.shape:
        raise ValueError("estimate_flow_batch expects two (B, H, W, 3) uint8 arrays of equal shape")
    ope = load_of_method(method)
    if params is not None:
        ope.parse_input_parameter(params)
    B, H, W = ims1.shape[:3]
    probe = np.empty((H, W, 2))
    if getattr(ope, 'auto_level', True) or ope._method_code == 0:
        ope.pyramid_levels = ope._auto_pyramid_levels(probe)
    P = ope._c_params(levels=ope.pyramid_levels)
    if ope.pyramid_levels < 1:
        P.pyramid_levels, P.auto_level = 0, 1
    ope._apply_solver(P)
    ctx = _lib.default_context(device)
    uv = ctx.pinned_empty((B, H, W, 2))       # page-locked: the result comes back in one DMA
    use_color = int(ope.color_images is not None)
    MAXB = 128                                # systems one persistent solve tracks (one scalar-update thread per system)
    stats = None
    for b0 in range(0, B, MAXB):
        b1 = min(B, b0 + MAXB)
        st = _lib.Stats()
        ctx.call("b200flow_estimate_rgb8", P, b1 - b0, H, W, _lib.ptr(ims1[b0:b1]), _lib.ptr(ims2[b0:b1]), use_color,
                 _lib.ptr(uv[b0:b1]), _lib.C.byref(st))
        d = st.as_dict()
        if stats is None:
            stats = d
        else:                                   # larger batches run as chunks of 128 pairs: counters and times add up
            for k in ("pcg_iters", "pcg_pixel_iters", "kernel_launches", "not_converged", "solver_ms", "warp_ms", "filter_ms",
                      "pre_ms", "total_ms", "solves"):
                stats[k] += d[k]
            for name, kd in d["kernels"].items():
                for k in kd:
                    stats["kernels"][name][k] += kd[k]
    return (uv, stats) if return_stats else uv


def shard_indices(n_items, rank, world_size):
    """Frame pairs are independent: pair k belongs to rank k mod world_size."""
    return list(range(rank, n_items, world_size))


def estimate_flow_sharded(ims1, ims2, method='classic+nl-fast', params=None, batch=8, gather=True):
    """Data-parallel flow over the ranks of an initialised torch.distributed job (or a single process).
    Every rank passes the same (N, H, W, 3) uint8 stacks; rank r computes pairs r, r+ws, ... on its own GPU in
    batches of `batch`; with gather=True the per-rank results are all-gathered so every rank returns (N, H, W, 2)."""
    import torch.distributed as dist
    ws = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if ws > 1 else 0
    N = len(ims1)
    mine = shard_indices(N, rank, ws)
    H, W = np.shape(ims1)[1:3]
    out = np.zeros((N, H, W, 2))
    for s in range(0, len(mine), batch):
        idx = mine[s:s + batch]
        out[idx] = estimate_flow_batch(np.asarray(ims1)[idx], np.asarray(ims2)[idx], method, params)
    if ws > 1 and gather:
        import torch
        t = torch.from_numpy(out)
        if dist.get_backend() == 'nccl':
            t = t.cuda()
        dist.all_reduce(t)              # disjoint supports: the sum is the concatenation
        out = t.cpu().numpy()
    return out
