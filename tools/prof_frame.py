"""Render a few frames of a bench workload (for ncu captures and per-level statistics).  Run under gpurun.

    python tools/prof_frame.py [--workload c2] [--frames 3] [--stats gpurun_out/level_stats.json]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pgi_raytracing_b200 import raytracer_for

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--stats", default="")
ap.add_argument("--params", default="", help="JSON object of render-parameter overrides, e.g. '{\"shader_mode\": 1}'")
a = ap.parse_args()
sc, p, desc = bench.workload(a.workload)
if a.params:
    p.update(json.loads(a.params)); desc += " " + a.params
rt = raytracer_for(sc)
for _ in range(a.frames):
    img, st = rt.render(p)
print(desc, {k: st[k] for k in ("primary", "shadow", "reflection", "refraction", "total", "frame_ms", "launches", "kernel_us", "primary_phase_us")})
rt.render(p, profile=3)
lv_t = rt.level_stats()
print("level timeline (ms after the frame kernel's first warp): " + "  ".join(f"L{l}:{x['shade_ms']:.3f}-{x['trace_ms']:.3f}" for l, x in enumerate(lv_t) if l >= 1 and x["rays"]))
if a.stats:
    img, st1 = rt.render(p, profile=1)
    lv_t = rt.level_stats()
    img, st3 = rt.render(p, profile=3)
    lv = rt.level_stats()
    for x, y in zip(lv, lv_t):
        x["trace_ms"], x["shade_ms"] = y["trace_ms"], y["shade_ms"]
    out = {"workload": desc, "build": rt.build_stats, "frame": st1, "instrumented": st3, "levels": lv}
    with open(a.stats, "w") as f:
        json.dump(out, f, indent=1)
    tot_q = st3["total"]
    print(f"nodes/ray={st3['nodes_visited'] / tot_q:.1f} tris/ray={st3['tris_tested'] / tot_q:.1f} max_nodes={st3['max_nodes_per_ray']}")
    for l, x in enumerate(lv):
        if x["rays"]:
            print(f"L{l}: rays={x['rays']} nodes/ray={x['nodes'] / max(x['rays'], 1):.1f} max={x['max_nodes']} shadow={x['shadow_rays']} "
                  f"sh_nodes/ray={x['shadow_nodes'] / max(x['shadow_rays'], 1):.1f} sh_max={x['shadow_max_nodes']} trace_ms={x['trace_ms']:.3f} shade_ms={x['shade_ms']:.3f}")
