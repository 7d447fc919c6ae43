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


class _DevicePtr:
    """A raw device pointer with the CUDA array interface, so torch can view it without copying."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2, "strides": None}


def share_frames(raytracer, n_frames: int, rank: int, device, group=None):
    """``n_frames`` full frames in rank 0's memory, mapped into every rank (``pgrt_frame_alloc / _export / _import``:
    CUDA IPC, opened from each rank's own device, so its resolve kernel stores into them over NVLink).
    Returns (pointers valid on this rank, rank-0 torch views or None); (None, None) when any rank cannot map them."""
    import torch
    import torch.distributed as dist

    nbytes = raytracer.width * raytracer.height * 16
    ok, ptrs, handles = 1, None, [None]
    try:
        if rank == 0:
            ptrs = [raytracer.frame_alloc(nbytes) for _ in range(n_frames)]
            handles = [[raytracer.frame_export(p) for p in ptrs]]
    except Exception:
        ok = 0
    dist.broadcast_object_list(handles, src=0, group=group)
    try:
        if rank != 0:
            if handles[0] is None:
                ok = 0
            else:
                ptrs = [raytracer.frame_import(h) for h in handles[0]]
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) != 1:
        return None, None
    views = None
    if rank == 0:
        views = [torch.as_tensor(_DevicePtr(p, (raytracer.height, raytracer.width, 4)), device=device) for p in ptrs]
    return ptrs, views


def share_host_frames(raytracer, n_frames: int, rank: int, device, group=None):
    """``n_frames`` full frames in HOST memory shared by every rank: rank 0 creates a memfd, the others open it through
    ``/proc/<pid>/fd``, every rank maps it and registers the mapping with its own device (``pgrt_host_frame_register``).
    Each rank's resolve kernel then stores its tiles straight into the host frame through its own PCIe link, so a frame
    that has to end in host memory needs no device-to-host copy on rank 0 (whose single link would carry all of it).
    Returns (device-side pointers valid on this rank, rank-0 CPU torch views or None, keep-alive); (None, None, None)
    when any rank cannot set it up."""
    import ctypes
    import mmap
    import os
    import torch
    import torch.distributed as dist

    stride = (raytracer.width * raytracer.height * 16 + 4095) // 4096 * 4096
    total = stride * n_frames
    ok, fd, mm, base, info = 1, -1, None, 0, [None]
    try:
        if rank == 0:
            fd = os.memfd_create("pgrt_host_frames")
            os.ftruncate(fd, total)
            info = [(os.getpid(), fd)]
    except Exception:
        ok = 0
    dist.broadcast_object_list(info, src=0, group=group)
    dev_base = 0
    try:
        if info[0] is None:
            ok = 0
        else:
            if rank != 0:
                fd = os.open(f"/proc/{info[0][0]}/fd/{info[0][1]}", os.O_RDWR)
            mm = mmap.mmap(fd, total)
            base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
            dev_base = raytracer.host_frame_register(base, total)
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)     # also: everybody has opened the memfd before rank 0 may close it
    if int(flag.item()) != 1:
        # some rank could not map or register the frames: undo what this one did and let the caller fall back
        if dev_base:
            try:
                raytracer.host_frame_unregister(base)
            except Exception:
                pass
        if mm is not None:
            mm.close()
        if fd >= 0:
            os.close(fd)
        return None, None, None
    views = None
    if rank == 0:
        flat = torch.frombuffer(mm, dtype=torch.float32).view(n_frames, stride // 4)
        views = [flat[i, : raytracer.height * raytracer.width * 4].view(raytracer.height, raytracer.width, 4) for i in range(n_frames)]
    return [dev_base + i * stride for i in range(n_frames)], views, (mm, fd, base)


class ShardedRenderer:
    """One rank of a tile-sharded render, ``depth`` frames in flight.

    mode "p2p" (default on GPUs when the frames can be peer-mapped): every rank's resolve kernel stores its tiles
    straight into rank 0's frame over NVLink (``pgrt_render_shard_to_frame_begin``); the only collective is a 4-byte
    NCCL all-reduce per frame that serves as the completion barrier.  mode "host": the same, but the frames live in host
    memory shared by all ranks (``share_host_frames``): every rank writes its tiles through its own PCIe link and
    ``frames`` are CPU tensors on rank 0.  mode "nccl": compact per-rank tile buffers,
    ``gather`` to rank 0, un-tile kernel.  ``begin(k)`` enqueues frame k; ``end(k)`` collects this rank's stats;
    on rank 0, ``frames[k % depth]`` holds frame k once the communication stream has passed ``ready[k % depth]``."""

    def __init__(self, raytracer, rank: int, world: int, device, depth: int = 1, mode: str = "auto"):
        import torch

        self.rt, self.rank, self.world, self.device, self.depth = raytracer, rank, world, device, depth
        raytracer.set_shard(rank, world)
        self.n_slots = raytracer.shard_pixels()
        self.slot_streams = [torch.cuda.ExternalStream(raytracer.slot_stream(i), device=device) for i in range(depth)]
        self.frames = None
        self.mode = "local" if world == 1 else mode
        self.frame_ptrs = None
        if self.mode in ("auto", "p2p"):
            self.frame_ptrs, views = share_frames(raytracer, depth, rank, device)
            if self.frame_ptrs is None and self.mode == "p2p":
                raise RuntimeError("ShardedRenderer: the frames of rank 0 cannot be peer-mapped (CUDA IPC)")
            self.mode = "p2p" if self.frame_ptrs is not None else "nccl"
            if self.mode == "p2p":
                self.frames = views
        if self.mode == "host":
            self.frame_ptrs, views, self._host_keep = share_host_frames(raytracer, depth, rank, device)
            if self.frame_ptrs is None:
                raise RuntimeError("ShardedRenderer: cannot set up frames in shared host memory")
            self.frames = views
        if self.mode not in ("p2p", "host") and rank == 0:
            self.frames = [torch.zeros((raytracer.height, raytracer.width, 4), dtype=torch.float32, device=device) for _ in range(depth)]
        if self.mode == "nccl":
            self.shards = [torch.zeros((self.n_slots, 4), dtype=torch.float32, device=device) for _ in range(depth)]
            self.gathered = [torch.empty((world, self.n_slots, 4), dtype=torch.float32, device=device) if rank == 0 else None for _ in range(depth)]
        self.token = torch.zeros(1, dtype=torch.float32, device=device)
        self._events = [torch.cuda.Event() for _ in range(4 * depth + 8)]   # recycled: an event is dead 2*depth+2 steps later
        self.host_s, self.host_n = [0.0, 0.0, 0.0, 0.0], 0                 # host time per part of begin() (bench reports it)
        self.ready = [None] * depth          # event on the communication stream: frame of this slot complete on rank 0
        self.history = {}                    # step -> ready event (kept for the last `depth` steps)
        torch.cuda.synchronize(device)

    def close(self):
        """Release the shared host frames of mode "host" (device frames go with the context)."""
        keep = getattr(self, "_host_keep", None)
        if keep is not None:
            import os
            import torch
            torch.cuda.synchronize(self.device)
            mm, fd, base = keep
            self.rt.host_frame_unregister(base)
            self.frames = None
            self._host_keep = None
            try:
                mm.close()
            except BufferError:      # a caller still holds a view of a frame: the mapping goes when that does
                pass
            os.close(fd)

    @property
    def frame(self):
        return self.frames[0] if self.frames else None

    def begin(self, k: int, params=None, profile: int = 0, before=None):
        """``before(stream)``: optional work to enqueue on the slot's stream ahead of the frame (bench: the L2 flush)."""
        import time
        import torch
        import torch.distributed as dist

        t = [time.perf_counter()]
        s = k % self.depth
        comm = torch.cuda.current_stream(self.device)
        # the slot's destination is free once what consumed its last frame has run: that work sits on rank 0's
        # communication stream before the barrier of the NEXT step, so waiting for that barrier (here, on every rank) suffices
        prev = self.history.get(k - self.depth + 1 if self.depth > 1 else k - 1)
        if prev is not None and k >= self.depth and self.mode != "local":
            self.slot_streams[s].wait_event(prev)
        if before is not None:
            before(self.slot_streams[s])
        t.append(time.perf_counter())
        if self.mode == "local":
            self.rt.render_begin(s, params, device_ptr=self.frames[s].data_ptr(), profile=profile)
            t.append(time.perf_counter())
            self.rt.stream_wait_slot(s, comm.cuda_stream)
        elif self.mode in ("p2p", "host"):
            self.rt.render_begin(s, params, frame_ptr=self.frame_ptrs[s], profile=profile)
            t.append(time.perf_counter())
            self.rt.stream_wait_slot(s, comm.cuda_stream)
            dist.all_reduce(self.token)          # completion barrier: 4 bytes; the pixels travelled inside the resolve kernel
        else:
            self.rt.render_begin(s, params, shard_ptr=self.shards[s].data_ptr(), profile=profile)
            t.append(time.perf_counter())
            self.rt.stream_wait_slot(s, comm.cuda_stream)
            if self.rank == 0:
                dist.gather(self.shards[s], list(self.gathered[s].unbind(0)), dst=0)
                self.rt.untile(self.gathered[s].data_ptr(), self.world, self.frames[s].data_ptr(), comm.cuda_stream)
            else:
                dist.gather(self.shards[s], None, dst=0)
        t.append(time.perf_counter())
        ev = self._events[k % len(self._events)]
        ev.record(comm)
        self.ready[s] = ev
        self.history[k] = ev
        self.history.pop(k - 2 * self.depth - 2, None)
        t.append(time.perf_counter())
        for i in range(4):                       # host seconds spent in: wait+before, render_begin, barrier/gather, event
            self.host_s[i] += t[i + 1] - t[i]
        self.host_n += 1

    def end(self, k: int) -> dict:
        return self.rt.render_end(k % self.depth)

    def render(self, params=None, profile=False):
        self.begin(0, params, int(profile))
        return self.end(0)
