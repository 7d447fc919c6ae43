"""Image-tile sharding across GPUs: one process per GPU, scene replicated, framebuffer gathered over NCCL.

The reference renders every pixel on one thread (pg1/simpleguidx11.cpp:102-118); pixels are independent, so the
frame is cut into 32x8-pixel tiles dealt round-robin to ranks (load balance: sky tiles are cheap, canopy tiles are
not).  Every rank's frame kernel stores its tiles at their final place of rank 0's frame (peer-mapped device memory
over NVLink, or host memory shared by all ranks); completion travels as one 32-bit flag per rank and frame slot in a
page of shared host memory (``FlagProtocol``), waited for with stream memory operations: no collective runs per frame.
Fallback: compact per-rank tile buffers, NCCL ``gather``, un-tile kernel.  The tile arithmetic and the flag protocol
are plain Python / numpy so the host logic is testable on CPU (gloo).
"""
from __future__ import annotations

import os

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


def share_frames(raytracer, n_frames: int, rank: int, device, group=None, nbytes: int | None = None):
    """``n_frames`` full frames (or buffers of ``nbytes``) in rank 0's memory, zeroed, mapped into every rank
    (``pgrt_frame_alloc / _export / _import``: CUDA IPC, opened from each rank's own device, so its kernels reach them over NVLink).
    Returns (pointers valid on this rank, rank-0 torch views or None); (None, None) when any rank cannot map them."""
    import torch
    import torch.distributed as dist

    as_frames = nbytes is None
    if nbytes is None:
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
    if rank == 0 and as_frames:
        views = [torch.as_tensor(_DevicePtr(p, (raytracer.height, raytracer.width, 4)), device=device) for p in ptrs]
    return ptrs, views


def share_host_region(raytracer, nbytes: int, rank: int, device, group=None, name: str = "pgrt_host_region"):
    """``nbytes`` (a multiple of 4096) of HOST memory shared by every rank of the box: rank 0 creates a memfd, the others open it
    through ``/proc/<pid>/fd``, every rank maps it and registers the mapping with its own device
    (``pgrt_host_frame_register``).  Returns (device-side base address valid on this rank, the mmap, keep-alive tuple);
    (None, None, None) on EVERY rank when any rank cannot set it up (nothing stays mapped or registered then)."""
    import ctypes
    import mmap
    import os
    import torch
    import torch.distributed as dist

    ok, fd, mm, base, info = 1, -1, None, 0, [None]
    try:
        if rank == 0:
            fd = os.memfd_create(name)
            os.ftruncate(fd, nbytes)
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
            mm = mmap.mmap(fd, nbytes)
            base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
            dev_base = raytracer.host_frame_register(base, nbytes)
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)     # also: everybody has opened the memfd before rank 0 may close it
    if int(flag.item()) != 1:
        # some rank could not map or register the region: undo what this one did and let the caller fall back
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
    return dev_base, mm, (mm, fd, base)


def release_host_region(raytracer, keep) -> None:
    import os
    mm, fd, base = keep
    raytracer.host_frame_unregister(base)
    try:
        mm.close()
    except BufferError:      # a caller still holds a view: the mapping goes when that does
        pass
    os.close(fd)


def share_host_frames(raytracer, n_frames: int, rank: int, device, group=None, bytes_per_pixel: int = 16):
    """``n_frames`` full frames in HOST memory shared by every rank (``share_host_region``).  Each rank's frame kernel then
    stores its tiles straight into the host frame through its own PCIe link, so a frame that has to end in host memory needs
    no device-to-host copy on rank 0 (whose single link would carry all of it).  ``bytes_per_pixel`` 16 = float RGBA, 4 = RGBA8.
    Returns (device-side pointers valid on this rank, rank-0 CPU torch views or None, keep-alive); (None, None, None)
    when any rank cannot set it up."""
    import torch

    stride = (raytracer.width * raytracer.height * bytes_per_pixel + 4095) // 4096 * 4096
    dev_base, mm, keep = share_host_region(raytracer, stride * n_frames, rank, device, group, "pgrt_host_frames")
    if dev_base is None:
        return None, None, None
    views = None
    if rank == 0:
        if bytes_per_pixel == 16:
            flat = torch.frombuffer(mm, dtype=torch.float32).view(n_frames, stride // 4)
            views = [flat[i, : raytracer.height * raytracer.width * 4].view(raytracer.height, raytracer.width, 4) for i in range(n_frames)]
        else:
            flat = torch.frombuffer(mm, dtype=torch.uint8).view(n_frames, stride)
            views = [flat[i, : raytracer.height * raytracer.width * 4].view(raytracer.height, raytracer.width, 4) for i in range(n_frames)]
    return [dev_base + i * stride for i in range(n_frames)], views, keep


class FlagProtocol:
    """Completion flags of a tile-sharded frame pipeline, ``depth`` frame slots, ``world`` ranks, in one page of shared memory:
    ``done[slot][rank]`` (stored by the last kernel of that rank's frame once its tiles of the slot's frame are in place) and
    ``consumed[slot]`` (stored by rank 0 behind whatever consumed the slot's frame).  The n-th frame rendered into a slot
    carries tag n in both (words start at 0; every wait is "word >= tag").  Every rank calls ``next_frame`` for the same
    slots in the same order, so all ranks agree on the tags without talking to each other.  No CUDA here:
    ``ShardedRenderer`` turns the answers into stream memory operations (``pgrt_slot_signal`` /
    ``pgrt_stream_wait_value32`` / ``pgrt_stream_write_value32``), tests/test_dist_gloo.py into host stores and polls."""

    def __init__(self, depth: int, world: int):
        self.depth, self.world = depth, world
        self.nbytes = (4 * (depth * world + depth) + 4095) // 4096 * 4096
        self.count = [0] * depth       # frames begun per slot = tag of the slot's latest frame
        self.last_slot = None

    def done_offset(self, slot: int, rank: int) -> int:
        return 4 * (slot * self.world + rank)

    def consumed_offset(self, slot: int) -> int:
        return 4 * (self.depth * self.world + slot)

    def next_frame(self, slot: int, rank: int) -> dict:
        """The memory operations around the next frame rendered into ``slot``, each an (offset, value) pair:
        ``mark_consumed``: rank 0 stores it on the consumer stream FIRST -- everything enqueued there since the previous
        ``next_frame`` has had the previous frame (None for the very first frame);
        ``before``: every rank's slot stream waits for these before the frame may overwrite the slot;
        ``signal``: what this rank's frame stores when its tiles are in place;
        ``done``: what rank 0's consumer stream waits for (all ranks' tiles of this frame are in place);
        ``done_count``: the same as ONE wait, when every rank adds 1 to a counter of the slot instead of storing its own word."""
        ops = {"mark_consumed": None, "before": []}
        if self.last_slot is not None:
            ops["mark_consumed"] = (self.consumed_offset(self.last_slot), self.count[self.last_slot])
        if self.count[slot] > 0:
            ops["before"].append((self.consumed_offset(slot), self.count[slot]))
        self.count[slot] = (self.count[slot] + 1) & 0xFFFFFFFF
        tag = self.count[slot]
        ops["signal"] = (self.done_offset(slot, rank), tag)
        ops["done"] = [(self.done_offset(slot, r), tag) for r in range(self.world)]
        # with ONE counter per slot that every rank's finished frame adds 1 to: byte offset of the counter, value to wait for
        ops["done_count"] = (4 * slot, (tag * self.world) & 0xFFFFFFFF)
        self.last_slot = slot
        return ops


class ShardedRenderer:
    """One rank of a tile-sharded render, ``depth`` frames in flight.

    mode "p2p" (default on GPUs when the frames can be peer-mapped): every rank's frame kernel stores its tiles straight
    into rank 0's frame over NVLink (``pgrt_render_shard_to_frame_begin``).  mode "host": the same, but the frames live in
    host memory shared by all ranks (``share_host_frames``): every rank writes its tiles through its own PCIe link and
    ``frames`` are CPU tensors on rank 0 (``rgba8=True``: R8G8B8A8_UNORM frames, a quarter of the bytes).
    Completion in both: ``FlagProtocol`` -- the last kernel of a rank's frame stores the frame's tag in that rank's word of a
    shared host page (only if no secondary-ray queue overflowed: a retried frame signals when the retry is done), rank 0's
    consumer stream waits for all ``world`` words, and stores the slot's "consumed" word behind whatever consumed the frame,
    which every rank's slot stream waits for before it overwrites the slot.  No collective runs per frame; when the driver
    lacks stream memory operations a 4-byte NCCL all-reduce per frame takes the flags' place.
    mode "nccl": compact per-rank tile buffers, ``gather`` to rank 0, un-tile kernel.

    ``begin(k)`` enqueues a frame into slot k % depth (every rank makes the same calls in the same order); ``end(k)`` collects
    this rank's stats (after a queue overflow it re-renders the frame, and only then is the frame signalled complete).  On
    rank 0, ``frames[k % depth]`` holds the frame for work enqueued on the current stream after ``begin(k)`` returned and before
    the next ``begin`` is called: that next call is what marks the frame as consumed (with ``depth`` = 1 it is also what lets
    every rank start the next frame).
    Mode "local" (one rank) orders nothing for the caller beyond ``stream_wait_slot``: sync before reusing a frame."""

    def __init__(self, raytracer, rank: int, world: int, device, depth: int = 1, mode: str = "auto", rgba8: bool = False, flags: bool = True,
                 counters: bool = True):
        import torch

        self.rt, self.rank, self.world, self.device, self.depth = raytracer, rank, world, device, depth
        raytracer.set_shard(rank, world)
        self.n_slots = raytracer.shard_pixels()
        self.slot_streams = [torch.cuda.ExternalStream(raytracer.slot_stream(i), device=device) for i in range(depth)]
        self.frames = None
        self.mode = "local" if world == 1 else mode
        self.rgba8 = bool(rgba8)
        if self.rgba8 and self.mode not in ("host", "local"):
            raise ValueError("ShardedRenderer: rgba8 frames are implemented for mode 'host' (and one rank)")
        self.frame_ptrs = None
        self._host_keep = None; self._ctl_keep = None; self._owned_frames = []
        if self.mode in ("auto", "p2p"):
            self.frame_ptrs, views = share_frames(raytracer, depth, rank, device)
            if self.frame_ptrs is None and self.mode == "p2p":
                raise RuntimeError("ShardedRenderer: the frames of rank 0 cannot be peer-mapped (CUDA IPC)")
            self.mode = "p2p" if self.frame_ptrs is not None else "nccl"
            if self.mode == "p2p":
                self.frames = views
                self._owned_frames = list(self.frame_ptrs)
        if self.mode == "host":
            self.frame_ptrs, views, self._host_keep = share_host_frames(raytracer, depth, rank, device, bytes_per_pixel=4 if self.rgba8 else 16)
            if self.frame_ptrs is None:
                raise RuntimeError("ShardedRenderer: cannot set up frames in shared host memory")
            self.frames = views
        if self.mode not in ("p2p", "host") and rank == 0:
            self.frames = [torch.zeros((raytracer.height, raytracer.width, 4), dtype=torch.uint8 if self.rgba8 else torch.float32, device=device) for _ in range(depth)]
        if self.mode == "nccl":
            self.shards = [torch.zeros((self.n_slots, 4), dtype=torch.float32, device=device) for _ in range(depth)]
            self.gathered = [torch.empty((world, self.n_slots, 4), dtype=torch.float32, device=device) if rank == 0 else None for _ in range(depth)]
        # completion flags in a page of shared host memory (all ranks or none); on top of that, when rank 0's memory can be
        # peer-mapped, one "frames done" counter per slot THERE: every rank's finished frame adds 1 to it over NVLink, and rank 0's
        # consumer stream needs one wait per frame on its own memory instead of one wait per rank on host memory
        self.proto, self.ctl_dev, self.ctl, self.cnt_ptr = None, 0, None, 0
        if self.mode in ("p2p", "host") and flags:
            proto = FlagProtocol(depth, world)
            dev_base, mm, keep = share_host_region(raytracer, proto.nbytes, rank, device, name="pgrt_flags")
            ok = 0
            if dev_base is not None:
                try:      # the driver must offer stream memory operations on every rank
                    raytracer.stream_wait_value32(torch.cuda.current_stream(device).cuda_stream, dev_base, 0)
                    ok = 1
                except Exception:
                    ok = 0
                import torch.distributed as dist
                t = torch.tensor([ok], dtype=torch.int32, device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                ok = int(t.item())
                if ok:
                    self.proto, self.ctl_dev, self._ctl_keep = proto, dev_base, keep
                    self.ctl = np.frombuffer(mm, dtype=np.uint32)
                    if counters:
                        ptrs, _ = share_frames(raytracer, 1, rank, device, nbytes=4096)
                        if ptrs is not None:
                            self.cnt_ptr = ptrs[0]; self._owned_frames.append(ptrs[0])
                else:
                    release_host_region(raytracer, keep)
        self.token = torch.zeros(1, dtype=torch.float32, device=device)
        self._probe = set(os.environ.get("PGRT_DIST_PROBE", "").split(","))     # measurement only (tools/scale_probe.py): parts of the protocol left out
        self._events = [torch.cuda.Event() for _ in range(4 * depth + 8)]   # recycled: an event is dead 2*depth+2 steps later
        self.host_s, self.host_n = [0.0, 0.0, 0.0, 0.0], 0                 # host time per part of begin() (bench reports it)
        self.ready = [None] * depth          # event on the communication stream: frame of this slot complete on rank 0
        self.history = {}                    # step -> ready event (kept for the last `depth` steps)
        torch.cuda.synchronize(device)

    def _host_sees(self, off: int, val: int, spin_s: float = 150e-6) -> bool:
        """Has the flag word at byte ``off`` reached ``val`` (cyclic >=, the comparison of cuStreamWaitValue32)?  Polls for at most ``spin_s``."""
        import time
        w = self.ctl
        i = off // 4
        t_end = None
        while True:
            if ((int(w[i]) - val) & 0xFFFFFFFF) < 0x80000000:
                return True
            now = time.perf_counter()
            if t_end is None:
                t_end = now + spin_s
            elif now >= t_end:
                return False

    @property
    def completion(self) -> str:
        if self.mode in ("local", "nccl"):
            return self.mode
        if self.proto is None:
            return "nccl-allreduce"
        return "flags+counter" if self.cnt_ptr else "flags"

    def close(self):
        """Release what this renderer shared between the ranks: host frames, the flag page, peer-mapped device frames."""
        import torch
        torch.cuda.synchronize(self.device)
        for s in range(self.depth):
            try:
                self.rt.slot_signal(s, 0, 0)
            except Exception:
                pass
        self.ctl = None
        if self._ctl_keep is not None:
            release_host_region(self.rt, self._ctl_keep); self._ctl_keep = None; self.proto = None
        if self._host_keep is not None:
            self.frames = None
            release_host_region(self.rt, self._host_keep); self._host_keep = None
        if self._owned_frames:
            import torch.distributed as dist
            self.frames = None
            if self.rank != 0:
                for p in self._owned_frames:
                    self.rt.frame_unmap(p)
            dist.barrier()                  # nobody holds a mapping any more
            if self.rank == 0:
                for p in self._owned_frames:
                    self.rt.frame_free(p)
            self._owned_frames = []

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
        ops = None
        if self.proto is not None:
            ops = self.proto.next_frame(s, self.rank)
            if self.rank == 0 and ops["mark_consumed"] is not None and "nomark" not in self._probe:   # everything enqueued on the consumer stream since the last begin() has had that frame
                off, val = ops["mark_consumed"]
                self.rt.stream_write_value32(comm.cuda_stream, self.ctl_dev + off, val)
            for off, val in ([] if "nobefore" in self._probe else ops["before"]):
                # The host looks first (the word is in host memory; it is usually set within microseconds of end(k - depth)
                # returning): a stream wait that is NOT yet satisfied parks the hardware channel of the slot's stream, and
                # with it every other stream the driver mapped to that channel -- 0.29 -> 0.38 ms per frame at two GPUs
                # with the default 8 channels (profiles/r2_scale_probe_n2.txt).  Only a consumer that really lags gets
                # the stream wait, so begin() still never blocks for longer than `spin_s`.
                if "nospin" in self._probe or not self._host_sees(off, val):
                    self.rt.stream_wait_value32(self.slot_streams[s].cuda_stream, self.ctl_dev + off, val)
        elif self.mode != "local":
            # the slot's destination is free once what consumed its last frame has run: that work sits on rank 0's
            # communication stream before the barrier of the NEXT step, so waiting for that barrier (here, on every rank) suffices
            prev = self.history.get(k - self.depth + 1 if self.depth > 1 else k - 1)
            if prev is not None and k >= self.depth:
                self.slot_streams[s].wait_event(prev)
        if before is not None:
            before(self.slot_streams[s])
        t.append(time.perf_counter())
        if self.mode == "local":
            if self.rgba8:
                raise ValueError("ShardedRenderer: one rank renders rgba8 frames through Raytracer.render_begin(host_ptr=..., rgba8=True)")
            self.rt.render_begin(s, params, device_ptr=self.frames[s].data_ptr(), profile=profile)
            t.append(time.perf_counter())
            self.rt.stream_wait_slot(s, comm.cuda_stream)
        elif self.mode in ("p2p", "host"):
            if ops is not None and "nosignal" in self._probe:
                self.rt.slot_signal(s, 0, 0)
            elif ops is not None:
                if self.cnt_ptr:
                    self.rt.slot_signal_add(s, self.cnt_ptr + ops["done_count"][0])
                else:
                    self.rt.slot_signal(s, self.ctl_dev + ops["signal"][0], ops["signal"][1])
            self.rt.render_begin(s, params, frame_ptr=self.frame_ptrs[s], profile=profile, rgba8=self.rgba8)
            t.append(time.perf_counter())
            if ops is not None and "nosignal" in self._probe:
                self.rt.stream_wait_slot(s, comm.cuda_stream)
            elif ops is not None:
                if self.rank == 0:
                    if self.cnt_ptr:
                        self.rt.stream_wait_value32(comm.cuda_stream, self.cnt_ptr + ops["done_count"][0], ops["done_count"][1])
                    else:
                        for off, val in ops["done"]:
                            self.rt.stream_wait_value32(comm.cuda_stream, self.ctl_dev + off, val)
            else:
                self.rt.stream_wait_slot(s, comm.cuda_stream)
                dist.all_reduce(self.token)          # completion barrier: 4 bytes; the pixels travelled inside the frame kernel
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
        for i in range(4):                       # host seconds spent in: wait+before, render_begin, completion, event
            self.host_s[i] += t[i + 1] - t[i]
        self.host_n += 1

    def end(self, k: int) -> dict:
        return self.rt.render_end(k % self.depth)

    def render(self, params=None, profile=False):
        self.begin(0, params, int(profile))
        return self.end(0)
