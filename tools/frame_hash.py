"""sha256 of a workload's frame (A/B of library variants: PGRT_LIB).    python tools/frame_hash.py [--workload c2] [--params '{"scheduler": 2}']"""
import argparse, hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pgi_raytracing_b200 import raytracer_for

ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="c2"); ap.add_argument("--params", default="")
a = ap.parse_args()
sc, p, desc = bench.workload(a.workload)
if a.params:
    p.update(json.loads(a.params))
img, st = raytracer_for(sc).render(p)
print(os.environ.get("PGRT_LIB", "default").split("/")[-1], desc, a.params, hashlib.sha256(img.tobytes()).hexdigest()[:16], st["total"], "rays", st["launches"], "launches")
