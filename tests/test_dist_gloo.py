"""World-size-2 gloo run of the multi-GPU host logic on CPU: every rank tiles its share of a known frame, the
shards are gathered to rank 0 (the only collective of the path) and un-tiled; the result must equal the frame."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pgi_raytracing_b200 import dist as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, w, h, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frame = np.random.default_rng(123).random((h, w, 4)).astype(np.float32)     # same on every rank (scene replicated)
    shard = torch.from_numpy(D.tile_numpy(frame, rank, world))
    assert shard.shape[0] == D.shard_pixels(w, h, world)
    g = D.gather_shards(shard, dst=0)
    if rank == 0:
        out = D.untile_numpy(g.numpy(), w, h, world)
        q.put(bool(np.array_equal(out, frame)))
    else:
        assert g is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("wh", [(96, 40), (70, 21)])
def test_gather_and_untile_world2(wh):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, wh[0], wh[1], q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


class _HostOnlyRaytracer:
    """Stands in for api.Raytracer where share_host_frames needs it: on a CPU box 'registering' a host mapping with the
    device is the identity (the kernels of a real rank store through the device alias of the same pages)."""
    def __init__(self, w, h):
        self.width, self.height, self.registered = w, h, []

    def host_frame_register(self, host_ptr, nbytes):
        self.registered.append((host_ptr, nbytes))
        return host_ptr

    def host_frame_unregister(self, host_ptr):
        self.registered = [r for r in self.registered if r[0] != host_ptr]


def _host_frames_worker(rank, world, port, w, h, q):
    import ctypes
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rt = _HostOnlyRaytracer(w, h)
    n_frames = 3
    ptrs, views, keep = D.share_host_frames(rt, n_frames, rank, torch.device("cpu"))
    assert ptrs is not None and len(ptrs) == n_frames and ptrs[0] % 4096 == 0 and (ptrs[1] - ptrs[0]) % 4096 == 0
    assert (views is not None) == (rank == 0)
    # every rank writes "its tiles" of every frame through its own mapping of the one memfd: rank r fills rows r, r+world, ...
    for f in range(n_frames):
        row = np.ctypeslib.as_array((ctypes.c_float * (w * h * 4)).from_address(ptrs[f])).reshape(h, w, 4)
        row[rank::world] = 100.0 * f + rank + 1
    dist.barrier()
    if rank == 0:
        ok = True
        for f in range(n_frames):
            got = views[f].numpy()
            for r in range(world):
                ok &= bool(np.all(got[r::world] == 100.0 * f + r + 1))
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_shared_host_frames_world2():
    """dist.share_host_frames (mode "host" of ShardedRenderer): rank 0 creates a memfd, rank 1 opens it through
    /proc/<pid>/fd, both map it; what each rank writes through its mapping is what rank 0 reads from its CPU views."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_host_frames_worker, args=(r, 2, port, 70, 21, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


class _FailingRaytracer(_HostOnlyRaytracer):
    def host_frame_register(self, host_ptr, nbytes):
        raise RuntimeError("cudaHostRegister failed (simulated)")


def _host_frames_failure_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rt = _FailingRaytracer(64, 16) if rank == 1 else _HostOnlyRaytracer(64, 16)
    fds_before = len(os.listdir("/proc/self/fd"))
    out = D.share_host_frames(rt, 2, rank, torch.device("cpu"))
    q.put((rank, out == (None, None, None), len(os.listdir("/proc/self/fd")) <= fds_before, len(rt.registered)))
    dist.barrier()
    dist.destroy_process_group()


def test_shared_host_frames_fall_back_together_when_one_rank_fails():
    """One rank cannot register the mapping: EVERY rank gets (None, None, None) (the caller then takes the NVLink / NCCL
    path) and the rank that succeeded has unregistered its mapping and closed its descriptor."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_host_frames_failure_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, True, True, 0), (1, True, True, 0)], got


def _flag_pipeline_worker(rank, world, port, depth, n_frames, q, counter_lock=None):
    """The flag protocol of dist.ShardedRenderer with host stores and polls in place of stream memory operations: every
    rank 'renders' its rows of a shared frame, rank 0 consumes the frame and releases the slot."""
    import ctypes
    import time
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h = 16, 8
    rt = _HostOnlyRaytracer(w, h)
    proto = D.FlagProtocol(depth, world)
    ctl_base, ctl_mm, ctl_keep = D.share_host_region(rt, proto.nbytes, rank, torch.device("cpu"), name="flags_test")
    ptrs, views, keep = D.share_host_frames(rt, depth, rank, torch.device("cpu"))
    ctl = np.frombuffer(ctl_mm, dtype=np.uint32)
    assert ctl_base is not None and (ctl == 0).all()
    # counter_lock: the "flags+counter" completion -- every rank's finished frame ADDS 1 to one word of the slot (on the GPU a
    # system-scope atomic in rank 0's memory; here a locked read-modify-write on a second shared page), rank 0 waits once
    cnt = None
    if counter_lock is not None:
        cnt_base, cnt_mm, cnt_keep = D.share_host_region(rt, 4096, rank, torch.device("cpu"), name="counters_test")
        cnt = np.frombuffer(cnt_mm, dtype=np.uint32)
        assert cnt_base is not None and (cnt == 0).all()
    frames = [np.ctypeslib.as_array((ctypes.c_float * (w * h * 4)).from_address(p)).reshape(h, w, 4) for p in ptrs]
    rng = np.random.default_rng(rank)

    def wait(off, val):
        t0 = time.time()
        while ctl[off // 4] < val:
            assert time.time() - t0 < 60, "flag wait timed out"
            time.sleep(0)
    ok, consumed_log = True, []
    for k in range(n_frames):
        s = k % depth
        ops = proto.next_frame(s, rank)
        if rank == 0 and ops["mark_consumed"] is not None:
            off, val = ops["mark_consumed"]; ctl[off // 4] = val
        for off, val in ops["before"]:
            wait(off, val)                          # the slot's previous frame has been consumed by rank 0
        if rng.random() < 0.5:
            time.sleep(0.002 * rng.random())        # ranks drift against each other
        frames[s][rank::world] = float(k + 1)       # "render": this rank's rows of frame k
        if cnt is not None:
            with counter_lock:
                cnt[ops["done_count"][0] // 4] += 1
        else:
            off, val = ops["signal"]; ctl[off // 4] = val
        if rank == 0:
            if cnt is not None:
                off, val = ops["done_count"]
                t0 = time.time()
                while ((int(cnt[off // 4]) - val) & 0xFFFFFFFF) >= 0x80000000:      # cyclic >=, as cuStreamWaitValue32 compares
                    assert time.time() - t0 < 60, "counter wait timed out"
                    time.sleep(0)
            else:
                for off, val in ops["done"]:
                    wait(off, val)
            ok &= bool((frames[s] == float(k + 1)).all())     # every rank's rows hold frame k: nobody ran ahead into this slot
            consumed_log.append(k)
    if rank == 0:
        q.put((ok, consumed_log == list(range(n_frames)), [int(x) for x in proto.count]))
    dist.barrier()
    D.release_host_region(rt, ctl_keep)
    if cnt is not None:
        del cnt
        D.release_host_region(rt, cnt_keep)
    dist.destroy_process_group()


@pytest.mark.parametrize("depth,counter", [(1, False), (3, False), (3, True)])
def test_flag_protocol_world2(depth, counter):
    """FlagProtocol at world size 2: tags agree without communication, no rank overwrites a slot before rank 0 has consumed
    it (depth 1: strict alternation), rank 0 never sees a frame before both ranks' rows are in place."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_frames = 40
    lock = ctx.Lock() if counter else None
    procs = [ctx.Process(target=_flag_pipeline_worker, args=(r, 2, port, depth, n_frames, q, lock)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    ok, in_order, counts = q.get(timeout=5)
    assert ok and in_order and sum(counts) == n_frames


def test_flag_protocol_offsets_and_tags():
    p = D.FlagProtocol(3, 4)
    assert p.nbytes == 4096 and len({p.done_offset(s, r) for s in range(3) for r in range(4)} | {p.consumed_offset(s) for s in range(3)}) == 15
    a = p.next_frame(0, 2)
    assert a["mark_consumed"] is None and a["before"] == [] and a["signal"] == (p.done_offset(0, 2), 1) and len(a["done"]) == 4
    b = p.next_frame(1, 2)
    assert b["mark_consumed"] == (p.consumed_offset(0), 1) and b["before"] == []
    p.next_frame(2, 2)
    c = p.next_frame(0, 2)
    assert c["mark_consumed"] == (p.consumed_offset(2), 1) and c["before"] == [(p.consumed_offset(0), 1)] and c["signal"][1] == 2
    # the counter form: one word per slot, every rank adds 1 per frame, rank 0 waits for frames * world
    assert a["done_count"] == (0, 4) and b["done_count"] == (4, 4) and c["done_count"] == (0, 8)
