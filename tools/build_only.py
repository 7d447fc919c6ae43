"""Commit one bench workload's scene a few times (for ncu launch lists of the BVH build).  Run under gpurun.
    python tools/build_only.py [--workload c5] [--reps 3]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pgi_raytracing_b200 import raytracer_for

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c5"); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
sc, p, desc = bench.workload(a.workload)
rt = raytracer_for(sc)
print(desc, rt.build_stats, flush=True)
for _ in range(a.reps - 1):
    t0 = time.perf_counter(); bs = rt.commit(); t1 = time.perf_counter()
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in bs.items()}, f"host wall {1e3 * (t1 - t0):.2f} ms", flush=True)
