"""Row-band split of ONE frame pair over the GPUs of a box (BASELINE config 5).

Not in the reference (which is single threaded); the loop it shortens is the linear solve inside every warp iteration
(ba.py:140-206, classic_nl.py:200-277).  One process per GPU.  Every rank calls the ordinary drop-in API -- `estimate_flow`,
the method classes -- with the SAME arguments; while a `RowBand` is active on the rank's context, every linear solve of a
level with >= 2^18 pixels is iterated band-wise (include/b200flow.h, b200flow_band_*): rank g owns a band of rows, reads one
halo row of two Krylov vectors from each neighbour's memory over NVLink P2P per iteration, and the dot products travel
through peer-mapped flags -- no NCCL and no host on the data path.  torch.distributed is only used here to all-gather the
64-byte CUDA IPC handles of the ranks' arena blocks.

    with RowBand.from_torch_distributed(H, W):            # inside a torchrun job, one rank per GPU
        uv = estimate_flow(im1, im2, 'classic++')          # identical result on every rank
"""
import ctypes as C

from optical_flow import _lib

BYTES_PER_PIXEL = 900          # device arena per pixel of the full-resolution frame (measured high-water mark ~620 B/pixel for
                               # classic++ with the pyramids, the system, the Krylov vectors and the staging of one call)


def arena_bytes_for(H, W):
    return int(H) * int(W) * BYTES_PER_PIXEL + (256 << 20)


class RowBand:
    def __init__(self, ctx, rank, world, arena_bytes, same_device=False):
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        ctx.call("b200flow_band_init", self.rank, self.world, C.c_ulonglong(int(arena_bytes)), int(bool(same_device)))
        self.active = self.world > 1

    def export(self):
        """(64-byte IPC handle, device address) of this rank's arena block"""
        h = (C.c_char * 64)()
        base = C.c_void_p()
        self.ctx.call("b200flow_band_export", C.cast(h, C.c_void_p), C.byref(base))
        return bytes(h.raw), int(base.value)

    def connect(self, peers, same_process=False):
        """peers: list over ranks of export() results"""
        for r, (handle, base) in enumerate(peers):
            if r == self.rank:
                continue
            if same_process:
                self.ctx.call("b200flow_band_connect", r, None, C.c_void_p(base))
            else:
                buf = C.create_string_buffer(handle, 64)
                self.ctx.call("b200flow_band_connect", r, C.cast(buf, C.c_void_p), None)

    def close(self):
        if self.active:
            self.active = False
            self.ctx.call("b200flow_band_close")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    @classmethod
    def from_torch_distributed(cls, H, W, device=None):
        """Collective over the ranks of an initialised torch.distributed job (one rank per GPU of one box)."""
        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        ctx = _lib.default_context(device)
        rb = cls(ctx, rank, world, arena_bytes_for(H, W))
        if world > 1:
            mine = rb.export()
            peers = [None] * world
            dist.all_gather_object(peers, mine)
            rb.connect(peers)
            dist.barrier()              # nobody starts before every block is mapped everywhere
        return rb
