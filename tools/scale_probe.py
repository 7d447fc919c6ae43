"""Where does a sharded frame's time go?  Under torchrun with N ranks, the C2 frame pipelined `depth` deep, one variant after another:
  solo        every rank renders its shard into a frame of its OWN memory, nothing crosses ranks (the kernels alone, all ranks at once)
  p2p-nosync  the tiles go to rank 0's frame over NVLink, completion is each rank's own (NOT a protocol: it isolates the cost of the stores)
  counter / words / allreduce   dist.ShardedRenderer with its three completion protocols
Device time between two events on the consumer stream, max over ranks.
    torchrun --nproc-per-node N tools/scale_probe.py [--depth 16] [--frames 600] [--variants solo,p2p-nosync,counter,words,allreduce] [--flush 1]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from pgi_raytracing_b200 import raytracer_for, default_params
from pgi_raytracing_b200.dist import ShardedRenderer, share_frames

ap = argparse.ArgumentParser()
ap.add_argument("--depth", default="16"); ap.add_argument("--frames", type=int, default=600); ap.add_argument("--flush", default="1")
ap.add_argument("--variants", default="solo,p2p-nosync,counter,words,allreduce"); ap.add_argument("--workload", default="c2"); ap.add_argument("--params", default="", help="JSON overrides of the render parameters")
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sc, p, desc = bench.workload(a.workload)
if a.params:
    import json
    p.update(json.loads(a.params))
rt = raytracer_for(sc, device=local)
params = default_params(**p)
FLUSH = int(torch.cuda.get_device_properties(dev).L2_cache_size * 1.125) // 4096 * 4096


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(run, n, depth):
    comm = torch.cuda.current_stream()
    run(3 * depth)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    e0.record(comm)
    rays = run(n, e0)
    e1.record(comm)
    barrier()
    h1 = time.perf_counter()
    t = torch.tensor([e0.elapsed_time(e1), (h1 - h0) * 1e3], dtype=torch.float64, device=dev)
    r = torch.tensor([float(rays)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(r)
    return float(t[0].item()) / n, float(t[1].item()) / n, float(r.item()) / n


for depth in [int(x) for x in a.depth.split(",")]:
    for flush in [int(x) for x in a.flush.split(",")]:
        for variant in a.variants.split(","):
            if variant in ("solo", "p2p-nosync"):
                rt.set_shard(rank, world)
                if variant == "solo":
                    frames = [torch.zeros((rt.height, rt.width, 4), dtype=torch.float32, device=dev) for _ in range(depth)]
                    ptrs = [f.data_ptr() for f in frames]
                else:
                    if world == 1:
                        continue
                    ptrs, _ = share_frames(rt, depth, rank, dev)
                    if ptrs is None:
                        print(variant, "frames cannot be peer-mapped"); continue
                streams = [torch.cuda.ExternalStream(rt.slot_stream(i), device=dev) for i in range(depth)]
                for s in range(depth):
                    rt.slot_signal(s, 0, 0)

                def run(n, e0=None):
                    comm = torch.cuda.current_stream(); rays = 0
                    if e0 is not None:
                        for s in streams:
                            s.wait_event(e0)
                    for k in range(n):
                        s = k % depth
                        if k >= depth:
                            rays += rt.render_end(s)["total"]
                        if flush:
                            rt.flush_l2(s, FLUSH, k & 0xFF)
                        rt.render_begin(s, params, frame_ptr=ptrs[s])
                        rt.stream_wait_slot(s, comm.cuda_stream)
                    for k in range(max(0, n - depth), n):
                        rays += rt.render_end(k % depth)["total"]
                    return rays
                ms, wall, rays = timed(run, a.frames, depth)
                if variant == "p2p-nosync":
                    barrier()
                    if rank != 0:
                        for q in ptrs:
                            rt.frame_unmap(q)
                    barrier()
                    if rank == 0:
                        for q in ptrs:
                            rt.frame_free(q)
                completion = "-"
            else:
                if world == 1:
                    continue
                sr = ShardedRenderer(rt, rank, world, dev, depth=depth, mode="p2p", flags=variant != "allreduce", counters=variant == "counter")

                def run(n, e0=None):
                    rays = 0
                    if e0 is not None:
                        for s in sr.slot_streams:
                            s.wait_event(e0)
                    for k in range(n):
                        if k >= depth:
                            rays += sr.end(k - depth)["total"]
                        sr.begin(k, params, before=(lambda st, k=k: rt.flush_l2(k % depth, FLUSH, k & 0xFF)) if flush else None)
                    for k in range(max(0, n - depth), n):
                        rays += sr.end(k)["total"]
                    return rays
                ms, wall, rays = timed(run, a.frames, depth)
                completion = sr.completion
                sr.close()
            if rank == 0:
                print(f"{a.params} N={world} depth {depth} flush {flush} {variant:11s} ({completion}): {ms:.4f} ms/frame device, {wall:.4f} wall, {rays / ms / 1e3:.0f} Mrays/s", flush=True)
            barrier()
if world > 1:
    dist.destroy_process_group()
