"""Two or more ranks (torchrun): the peer-mapped direct-write path, the NCCL gather path, the shared-host-frame path
(every rank stores its tiles into one memfd-backed host frame) and a single-GPU render must give the same frame, bit for bit.  Run: torchrun --nproc-per-node 2 tools/p2p_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgi_raytracing_b200 import raytracer_for, scenes, default_params, to_srgb8
from pgi_raytracing_b200.dist import ShardedRenderer

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sc = scenes.avenger_proxy()
p = dict(sampling_width=2, seed=3, max_depth=6)
rt = raytracer_for(sc, device=local)
ref, st_ref = rt.render(p)                       # every rank renders the whole frame alone first
out = {}
# (mode, completion by flags in shared host memory?, 8-bit frames?, squeeze the ray pool so that frames overflow and are re-rendered?)
# flags: True = flags + one counter per slot in rank 0's memory (default), "words" = one flag word per rank in host memory, False = NCCL all-reduce
cases = [("p2p", True, False, False), ("p2p", "words", False, False), ("p2p", False, False, False), ("nccl", True, False, False), ("host", True, False, False),
         ("host", True, True, False), ("p2p", True, False, True), ("host", "words", False, True)]
for mode, flags, rgba8, squeeze in cases:
    if squeeze:      # every rank's first frames overflow their queues: the retry, not the overflowed attempt, must be what rank 0 sees
        os.environ["PGRT_MIN_LEVEL_CAP"] = "3000"; os.environ["PGRT_LEVEL_CAP_FACTOR"] = "0.002"
        rt2 = raytracer_for(sc, device=local)
        del os.environ["PGRT_MIN_LEVEL_CAP"]; del os.environ["PGRT_LEVEL_CAP_FACTOR"]
    else:
        rt2 = rt
    sr = ShardedRenderer(rt2, rank, world, dev, depth=3, mode=mode, flags=bool(flags), rgba8=rgba8, counters=flags is True)
    rays = 0; retries = 0
    snap = []                                     # rank 0: a copy of every frame, enqueued on the consumer stream right behind begin()
    for k in range(7):                            # more frames than slots: buffers are reused
        if k >= 3:
            st = sr.end(k - 3); rays = st["total"]; retries += st["overflow_retries"]
        sr.begin(k, p)
        if rank == 0:
            snap.append(sr.frames[k % 3].to(dev, non_blocking=True).clone() if mode != "host" else None)
    for k in range(4, 7):
        st = sr.end(k); rays = st["total"]; retries += st["overflow_retries"]
    torch.cuda.synchronize(); dist.barrier()
    tot = torch.tensor([float(rays), float(retries)], device=dev); dist.all_reduce(tot)
    if rank == 0:
        want = np.concatenate([to_srgb8(ref), np.full(ref.shape[:2] + (1,), 255, np.uint8)], -1) if rgba8 else ref
        for s in range(3):
            f = sr.frames[s].cpu().numpy().copy()
            assert np.array_equal(f, want, equal_nan=True), (mode, s)
        for k, t in enumerate(snap):              # what a consumer ordered behind begin(k) saw: never a half-finished or overflowed frame
            if t is not None:
                assert np.array_equal(t.cpu().numpy(), want, equal_nan=True), (mode, "snapshot", k)
        assert int(tot[0].item()) == st_ref["total"], (mode, tot[0].item(), st_ref["total"])
        assert (tot[1].item() > 0) == squeeze, (mode, squeeze, tot[1].item())
        print(f"{mode}{' rgba8' if rgba8 else ''}{' squeezed-pool' if squeeze else ''}: {world} ranks, 3 slots x 7 frames bit-identical to the single-GPU frame; "
              f"rays {int(tot[0].item())} == {st_ref['total']}; overflow retries {int(tot[1].item())}; mode used = {sr.mode}; completion = {sr.completion}")
    sr.close()
    rt2.set_shard(0, 1)
    if rt2 is not rt:
        rt2.close()
    dist.barrier()
dist.destroy_process_group()
