"""One GPU renders ONE rank's shard of the C2 frame (pgrt_set_shard(0, n)) with `depth` frames in flight: the per-rank cost of
a sharded frame without any other rank, flag or link in the picture.  Run under gpurun.
    python tools/quick_shard.py [--ranks 8] [--depth 16] [--frames 400] [--flush 1]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pgi_raytracing_b200 import raytracer_for, default_params

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, default=8); ap.add_argument("--depth", type=int, default=16); ap.add_argument("--frames", type=int, default=400)
ap.add_argument("--flush", type=int, default=1); ap.add_argument("--tag", default=""); ap.add_argument("--reps", type=int, default=1); ap.add_argument("--params", default="")
a = ap.parse_args()
sc, p, desc = bench.workload("c2")
if a.params:
    import json
    p.update(json.loads(a.params))
rt = raytracer_for(sc)
params = default_params(**p)
rt.set_shard(0, a.ranks)
frames = [torch.zeros((rt.height, rt.width, 4), dtype=torch.float32, device="cuda") for _ in range(a.depth)]
FLUSH = int(torch.cuda.get_device_properties(0).L2_cache_size * 1.125) // 4096 * 4096
torch.cuda.synchronize()
def run(n):
    rays = 0
    torch.cuda.synchronize(); t0 = time.perf_counter(); host = 0.0
    for k in range(n):
        s = k % a.depth
        if k >= a.depth:
            rays += rt.render_end(s)["total"]
        h0 = time.perf_counter()
        if a.flush:
            rt.flush_l2(s, FLUSH, k & 0xFF)
        rt.render_begin(s, params, frame_ptr=frames[s].data_ptr())
        host += time.perf_counter() - h0
    for k in range(max(0, n - a.depth), n):
        rays += rt.render_end(k % a.depth)["total"]
    torch.cuda.synchronize()
    return time.perf_counter() - t0, rays, host / n
run(3 * a.depth)
for rep in range(a.reps):
    t, rays, host = run(a.frames)
    print(f"{a.tag} {a.params} shard 1/{a.ranks} depth {a.depth} flush {a.flush} keep={os.environ.get('PGRT_KEEP_CTAS', '8')} ctas={os.environ.get('PGRT_FRAME_CTAS', 'auto')}: "
          f"{t / a.frames * 1e3:.4f} ms/frame, {rays / t / 1e6:.0f} Mrays/s of this rank's rays (x{a.ranks} = {rays / t / 1e6 * a.ranks:.0f}), host {host * 1e6:.1f} us/frame", flush=True)
