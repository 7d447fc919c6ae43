"""Image-tile sharding across GPUs: one process per GPU, scene replicated, framebuffer gathered over NCCL.

The reference renders every pixel on one thread (pg1/simpleguidx11.cpp:102-118); pixels are independent, so the
frame is cut into 32x8-pixel tiles dealt round-robin to ranks (load balance: sky tiles are cheap, canopy tiles are
not) and the only exchange is the gather of the per-rank compact tile buffers into rank 0, followed by the un-tile
kernel.  The same tile arithmetic is restated here in numpy so the host logic is testable on CPU (gloo).
"""
from __future__ import annotations

import numpy as np

TILE_W, TILE_H = 32, 8
TILE_PIXELS = TILE_W * TILE_H


def tiles_xy(width: int, height: int):
    return (width + TILE_W - 1) // TILE_W, (height + TILE_H - 1) // TILE_H


def shard_pixels(width: int, height: int, n_ranks: int) -> int:
    """Padded pixel slots per rank (``pgrt_shard_pixels``)."""
    tx, ty = tiles_xy(width, height)
    return (tx * ty + n_ranks - 1) // n_ranks * TILE_PIXELS


def slot_pixels(width: int, height: int, rank: int, n_ranks: int):
    """Slot -> (x, y, valid) of one rank: numpy mirror of ``slot_to_pixel`` in csrc/render.cuh."""
    tx, ty = tiles_xy(width, height)
    n = shard_pixels(width, height, n_ranks)
    slot = np.arange(n, dtype=np.int64)
    k, q = slot // TILE_PIXELS, slot % TILE_PIXELS
    t = k * n_ranks + rank
    b, l = q >> 5, q & 31
    x = (t % tx) * TILE_W + (b & 3) * 8 + (l & 7)
    y = (t // tx) * TILE_H + (b >> 2) * 4 + (l >> 3)
    valid = (t < tx * ty) & (x < width) & (y < height)
    return x, y, valid


def untile_numpy(gathered: np.ndarray, width: int, height: int, n_ranks: int) -> np.ndarray:
    """[n_ranks * shard_pixels, C] rank-major compact buffers -> [H, W, C] frame (mirror of ``k_untile``)."""
    spr = shard_pixels(width, height, n_ranks)
    out = np.zeros((height, width, gathered.shape[-1]), gathered.dtype)
    for r in range(n_ranks):
        x, y, valid = slot_pixels(width, height, r, n_ranks)
        out[y[valid], x[valid]] = gathered[r * spr:(r + 1) * spr][valid]
    return out


def tile_numpy(frame: np.ndarray, rank: int, n_ranks: int) -> np.ndarray:
    """[H, W, C] -> this rank's compact [shard_pixels, C] buffer (unused slots zero)."""
    h, w = frame.shape[:2]
    x, y, valid = slot_pixels(w, h, rank, n_ranks)
    out = np.zeros((x.shape[0], frame.shape[-1]), frame.dtype)
    out[valid] = frame[y[valid], x[valid]]
    return out


def gather_shards(shard, group=None, dst: int = 0):
    """Gather equal-sized per-rank shard tensors (torch, any device) to ``dst``; returns the rank-major
    concatenation on ``dst`` and ``None`` elsewhere.  NCCL on GPUs, gloo on CPU; no other collective is used."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return shard
    rank = dist.get_rank(group)
    if rank == dst:
        out = torch.empty((world,) + tuple(shard.shape), dtype=shard.dtype, device=shard.device)
        dist.gather(shard, list(out.unbind(0)), dst=dst, group=group)
        return out.reshape((world * shard.shape[0],) + tuple(shard.shape[1:]))
    dist.gather(shard, None, dst=dst, group=group)
    return None


class ShardedRenderer:
    """One rank of a tile-sharded render.  ``begin(k)`` enqueues this rank's tiles of frame k in slot k % depth and,
    on the communication stream (torch's current stream), the gather to rank 0 and the un-tile; ``end(k)`` collects
    the stats.  With depth > 1 the gather of one frame overlaps the rendering of the next."""

    def __init__(self, raytracer, rank: int, world: int, device, depth: int = 1):
        import torch

        self.rt, self.rank, self.world, self.device, self.depth = raytracer, rank, world, device, depth
        raytracer.set_shard(rank, world)
        self.n_slots = raytracer.shard_pixels()
        self.frame = torch.zeros((raytracer.height, raytracer.width, 4), dtype=torch.float32, device=device) if rank == 0 else None
        self.shards = [torch.zeros((self.n_slots, 4), dtype=torch.float32, device=device) for _ in range(depth)]
        self.gathered = [torch.empty((world, self.n_slots, 4), dtype=torch.float32, device=device) if (rank == 0 and world > 1) else None
                         for _ in range(depth)]
        self.slot_streams = [torch.cuda.ExternalStream(raytracer.slot_stream(i), device=device) for i in range(depth)]
        self.comm_done = [None] * depth
        torch.cuda.synchronize(device)

    def begin(self, k: int, params=None, profile: int = 0, before=None):
        """``before(stream)``: optional work to enqueue on the slot's stream ahead of the frame (bench: the L2 flush)."""
        import torch
        import torch.distributed as dist

        s = k % self.depth
        comm = torch.cuda.current_stream(self.device)
        if self.comm_done[s] is not None:          # the slot's shard buffer is free once its last gather has run
            self.slot_streams[s].wait_event(self.comm_done[s])
        if before is not None:
            before(self.slot_streams[s])
        if self.world == 1:                          # nothing to gather: resolve straight into the frame
            self.rt.render_begin(s, params, device_ptr=self.frame.data_ptr(), profile=profile)
            return
        self.rt.render_begin(s, params, shard_ptr=self.shards[s].data_ptr(), profile=profile)
        self.rt.stream_wait_slot(s, comm.cuda_stream)
        if self.rank == 0:
            dist.gather(self.shards[s], list(self.gathered[s].unbind(0)), dst=0)
            self.rt.untile(self.gathered[s].data_ptr(), self.world, self.frame.data_ptr(), comm.cuda_stream)
        else:
            dist.gather(self.shards[s], None, dst=0)
        ev = torch.cuda.Event(); ev.record(comm)
        self.comm_done[s] = ev

    def end(self, k: int) -> dict:
        return self.rt.render_end(k % self.depth)

    def render(self, params=None, profile=False):
        self.begin(0, params, int(profile))
        return self.end(0)
