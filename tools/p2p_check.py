"""Two or more ranks (torchrun): the peer-mapped direct-write path, the NCCL gather path, the shared-host-frame path
(every rank stores its tiles into one memfd-backed host frame) and a single-GPU render must give the same frame, bit for bit.  Run: torchrun --nproc-per-node 2 tools/p2p_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgi_raytracing_b200 import raytracer_for, scenes, default_params
from pgi_raytracing_b200.dist import ShardedRenderer

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sc = scenes.avenger_proxy()
p = dict(sampling_width=2, seed=3, max_depth=6)
rt = raytracer_for(sc, device=local)
ref, st_ref = rt.render(p)                       # every rank renders the whole frame alone first
out = {}
for mode in ("p2p", "nccl", "host"):
    sr = ShardedRenderer(rt, rank, world, dev, depth=3, mode=mode)
    rays = 0
    for k in range(7):                            # more frames than slots: buffers are reused
        if k >= 3:
            rays = sr.end(k - 3)["total"]
        sr.begin(k, p)
    for k in range(4, 7):
        rays = sr.end(k)["total"]
    torch.cuda.synchronize(); dist.barrier()
    tot = torch.tensor([float(rays)], device=dev); dist.all_reduce(tot)
    if rank == 0:
        for s in range(3):
            f = sr.frames[s].cpu().numpy().copy()
            assert np.array_equal(f, ref, equal_nan=True), (mode, s, float(np.nanmax(np.abs(f - ref))))
        assert int(tot.item()) == st_ref["total"], (mode, tot.item(), st_ref["total"])
        print(f"{mode}: {world} ranks, 3 slots x 7 frames bit-identical to the single-GPU frame; rays {int(tot.item())} == {st_ref['total']}; mode used = {sr.mode}")
    sr.close()
    rt.set_shard(0, 1)
    dist.barrier()
dist.destroy_process_group()
