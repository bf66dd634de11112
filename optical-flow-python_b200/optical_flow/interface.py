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


def _frames_and_guidance(im1, im2, want_color):
    """The image stack and the guidance image estimate_flow hands to the driver (interface.py:40-66): RGB -> gray frames
    + scaled Lab guidance; gray or 1-2 channel inputs -> the frames themselves (channels of frame 1, then of frame 2)."""
    if im1.ndim == 3 and im1.shape[2] >= 3:
        images = np.stack([_rgb2gray(im1), _rgb2gray(im2)], axis=2)
    else:
        images = np.stack([im1, im2], axis=2) if im1.ndim == 2 else np.concatenate([im1, im2], axis=2)
    color = None
    if want_color:
        color = _rgb2lab(im1, scale_channels=True) if im1.ndim == 3 and im1.shape[2] >= 3 else im1.copy()
    return images, color


def estimate_flow(im1, im2, method='classic+nl-fast', params=None):
    """Optical flow (H, W, 2) between two images ((H, W) gray or (H, W, 3) RGB, float or uint8)."""
    im1, im2, ope = _prepare(im1, im2, method, params)
    ope.images, color = _frames_and_guidance(im1, im2, ope.color_images is not None)
    if ope.color_images is not None:
        ope.color_images = color
    H, W = im1.shape[:2]
    return ope.compute_flow(np.zeros((H, W, 2)))


def estimate_flow_batch(ims1, ims2, method='classic+nl-fast', params=None, device=None, return_stats=False):
    """Flow for B same-size pairs in ONE device call per 128 pairs: ims1, ims2 (B, H, W, 3) uint8 -> (B, H, W, 2), colour
    conversion (gray, Lab) on the device too (b200flow_estimate_rgb8) -- the fast path.  Any other input estimate_flow accepts
    ((B, H, W) gray, (B, H, W, C) float or uint8 stacks) is prepared pair by pair exactly as estimate_flow does and solved as
    one batch through b200flow_estimate_mc."""
    ims1 = np.ascontiguousarray(ims1)
    ims2 = np.ascontiguousarray(ims2)
    if ims1.ndim not in (3, 4) or ims1.shape != ims2.shape:
        raise ValueError("estimate_flow_batch expects two (B, H, W) or (B, H, W, C) arrays of equal shape")
    ope = load_of_method(method)
    if params is not None:
        ope.parse_input_parameter(params)
    B, H, W = ims1.shape[:3]
    if not (ims1.dtype == np.uint8 and ims2.dtype == np.uint8 and ims1.ndim == 4 and ims1.shape[3] == 3):
        return _estimate_flow_batch_general(ope, ims1, ims2, device, return_stats)
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
        stats = _sum_stats(stats, st.as_dict())    # larger batches run as chunks of 128 pairs: counters and times add up
    return (uv, stats) if return_stats else uv


def _sum_stats(stats, d):
    if stats is None:
        return d
    for k in ("pcg_iters", "pcg_pixel_iters", "kernel_launches", "not_converged", "solver_ms", "warp_ms", "filter_ms",
              "pre_ms", "total_ms", "solves"):
        stats[k] += d[k]
    for name, kd in d["kernels"].items():
        for k in kd:
            stats["kernels"][name][k] += kd[k]
    return stats


def _estimate_flow_batch_general(ope, ims1, ims2, device, return_stats):
    """Batched estimate for inputs other than uint8 RGB: frames / guidance image built per pair as estimate_flow builds them
    (interface.py:40-66), the coarse-to-fine loop of up to 128 pairs in one b200flow_estimate_mc call."""
    B, H, W = ims1.shape[:3]
    want_color = ope.color_images is not None
    frames, guides = [], []
    for b in range(B):
        f, g = _frames_and_guidance(np.asarray(ims1[b], dtype=float), np.asarray(ims2[b], dtype=float), want_color)
        frames.append(f)
        guides.append(g)
    images = _lib.f64(np.stack(frames))
    color = None
    if want_color:
        color = _lib.f64(np.stack(guides))
        if color.size < B * H * W:              # weighted_median.py:42-47: no usable guidance image -> plain median
            color = None
    nc = images.shape[3] // 2
    Cn = 0 if color is None else (1 if color.ndim == 3 else color.shape[3])
    if getattr(ope, 'auto_level', True) or ope._method_code == 0:
        ope.pyramid_levels = ope._auto_pyramid_levels(images[0])
    P = ope._c_params(levels=ope.pyramid_levels)
    if ope.pyramid_levels < 1:
        P.pyramid_levels, P.auto_level = 0, 1
    ope._apply_solver(P)
    ctx = _lib.default_context(device)
    uv = ctx.pinned_empty((B, H, W, 2))
    stats = None
    for b0 in range(0, B, 128):
        b1 = min(B, b0 + 128)
        st = _lib.Stats()
        ctx.call("b200flow_estimate_mc", P, b1 - b0, H, W, nc, Cn, _lib.ptr(images[b0:b1]),
                 _lib.ptr(color[b0:b1]) if color is not None else None, None, _lib.ptr(uv[b0:b1]), _lib.C.byref(st))
        stats = _sum_stats(stats, st.as_dict())
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
