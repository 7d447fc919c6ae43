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
