"""Quick C2 numbers for A/B runs of library variants (PGRT_LIB) and knobs: blocking frame latency with the frame kernel's
phases, then a short pipelined run.  Run under gpurun.    python tools/quick_c2.py [--workload c2] [--frames 200] [--depth 4]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pgi_raytracing_b200 import raytracer_for, default_params

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2"); ap.add_argument("--frames", type=int, default=200); ap.add_argument("--depth", type=int, default=4)
ap.add_argument("--tag", default=""); ap.add_argument("--params", default="")
a = ap.parse_args()
sc, p, desc = bench.workload(a.workload)
if a.params:
    import json
    p.update(json.loads(a.params))
rt = raytracer_for(sc)
params = default_params(**p)
dev = torch.zeros((rt.height, rt.width, 4), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
lat = []
for _ in range(6):
    st = rt.render_device(dev.data_ptr(), params, profile=True)
    lat.append((st["frame_ms"], st["kernel_us"], st["primary_phase_us"], st["pool_iters"], st["reflection"] + st["refraction"]))
lat = sorted(lat)[len(lat) // 2]
frames = [torch.zeros((rt.height, rt.width, 4), dtype=torch.float32, device="cuda") for _ in range(a.depth)]
torch.cuda.synchronize()
def run(n):
    rays = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        if k >= a.depth:
            rays += rt.render_end((k - a.depth) % a.depth)["total"]
        rt.render_begin(k % a.depth, params, device_ptr=frames[k % a.depth].data_ptr())
    for k in range(max(0, n - a.depth), n):
        rays += rt.render_end(k % a.depth)["total"]
    torch.cuda.synchronize()
    return time.perf_counter() - t0, rays
run(3 * a.depth)
t, rays = run(a.frames)
cyc = [0, 0, 0, 0]
for sl in range(a.depth):
    c = rt.frame_cycles(sl)
    cyc = [x + y for x, y in zip(cyc, c)]
if cyc[3]:
    print(f"   warp time per frame (pipelined, mean of the last {a.depth} frames): primary {cyc[0] / a.depth / 1.965e3:.0f} us-warp, secondary {cyc[1] / a.depth / 1.965e3:.0f}, "
          f"resident {cyc[2] / a.depth / 1.965e3:.0f} ({cyc[3] / a.depth:.0f} warps): {100 * cyc[0] / cyc[2]:.1f} % / {100 * cyc[1] / cyc[2]:.1f} % / other {100 * (cyc[2] - cyc[0] - cyc[1]) / cyc[2]:.1f} %", flush=True)
print(f"{a.tag or os.environ.get('PGRT_LIB', 'default')} keep={os.environ.get('PGRT_KEEP_CTAS', '8')} policy={os.environ.get('PGRT_POOL_POLICY', '1')} claim={os.environ.get('PGRT_MIN_CLAIM', '32')} "
      f"ctas={os.environ.get('PGRT_FRAME_CTAS_PER_SM', 'max')} | blocking frame {lat[0]:.3f} ms (kernel {lat[1]} us, primaries done at {lat[2]} us, {lat[4]} secondary rays in {lat[3]} warp iterations) | "
      f"pipelined x{a.depth}: {t / a.frames * 1e3:.3f} ms/frame = {rays / t / 1e6:.0f} Mrays/s", flush=True)
