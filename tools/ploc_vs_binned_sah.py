"""CPU only: the product's tree (Morton order refined by PLOC, collapsed to 8-wide nodes; the g++ build of csrc/bvh8.cuh +
traverse.cuh, tests/emul/) against a top-down binned-SAH tree put through the SAME collapse and the SAME traversal:
SAH cost, nodes fetched and triangles tested per ray.  This is the measurement behind DESIGN.md's deviation note
("LBVH refined by binned SAH" -> PLOC).      python tools/ploc_vs_binned_sah.py > profiles/r2_ploc_vs_binned_sah.txt"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from pgi_raytracing_b200 import scenes
from oracle import oracle
import test_bvh8_emul as T



def rays_for(sc, n_secondary, seed):
    """Camera rays of a small frame + rays that start inside the scene (secondary-like)."""
    orc = oracle.Oracle()          # camera only: no scene is loaded into the oracle
    c = sc.camera
    orc.set_camera(192, 108, c.fov_y, c.view_from, c.view_at)
    pr = orc.primary_rays(oracle.make_params(sampling_width=1, jitter=0, aperture=0.0))
    prim = np.zeros((pr.shape[0], 8), np.float32); prim[:, :4] = pr[:, :4]; prim[:, 4:7] = pr[:, 4:7]; prim[:, 7] = np.finfo(np.float32).max
    pos = T.scene_pos(sc)
    sec = T.random_rays(pos, n_secondary, seed)
    return prim, sec


def measure(lib, pos, rays, builder, layout=1):
    t0 = time.time()
    h = lib.emul_build(pos.ctypes.data, pos.shape[0], builder, layout)
    dt = time.time() - t0
    out = np.zeros((rays.shape[0], 4), np.float32); stats = np.zeros((rays.shape[0], 2), np.uint32)
    lib.emul_trace(h, rays.ctypes.data, rays.shape[0], out.ctypes.data, stats.ctypes.data, 0)
    r = dict(sah=lib.emul_sah(h), nodes=lib.emul_nodes(h), depth=lib.emul_depth(h), nodes_per_ray=float(stats[:, 0].mean()), tris_per_ray=float(stats[:, 1].mean()),
             worst=int(stats[:, 0].max()), build_s=dt, hits=out.copy())
    lib.emul_free(h)
    return r


if __name__ == "__main__":
    import pytest  # noqa: F401  (test_bvh8_emul imports it)
    # build the emulator library exactly as the test fixture does
    deps = [T.SRC]
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    import subprocess
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-mavx2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-o", T.LIB, T.SRC])
    lib = C.CDLL(T.LIB)
    lib.emul_build.restype = C.c_void_p; lib.emul_build.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int]
    lib.emul_free.argtypes = [C.c_void_p]; lib.emul_nodes.restype = C.c_uint32; lib.emul_nodes.argtypes = [C.c_void_p]
    lib.emul_depth.restype = C.c_uint32; lib.emul_depth.argtypes = [C.c_void_p]; lib.emul_sah.restype = C.c_double; lib.emul_sah.argtypes = [C.c_void_p]
    lib.emul_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
    cases = [("C2 avenger stand-in", scenes.avenger_proxy(with_images=False)), ("C4 palm grove, 150 palms", scenes.palm_grove(n_palms=150)),
             ("C5 soup, 300 k triangles", scenes.triangle_soup(300_000, seed=1))]
    print("# PLOC (product builder) vs top-down binned SAH (16 bins), both collapsed to 8-wide float-plane nodes and traversed by the product's")
    print("# traversal code compiled for the CPU (tests/emul/bvh8_emul.cpp).  nodes / tris = wide nodes fetched / triangles tested per ray.")
    for name, sc in cases:
        pos = T.scene_pos(sc)
        prim, sec = rays_for(sc, 20000, 5)
        print(f"\n## {name}: {pos.shape[0]} triangles; {prim.shape[0]} camera rays (192x108), {sec.shape[0]} rays from inside the scene")
        res = {}
        for label, b in (("PLOC", 0), ("binned SAH", 2)):
            rp = measure(lib, pos, prim, b); rs = measure(lib, pos, sec, b)
            res[label] = (rp, rs)
            print(f"  {label:11s} SAH cost {rp['sah']:8.2f}  wide nodes {rp['nodes']:7d} depth {rp['depth']:2d} | camera rays: {rp['nodes_per_ray']:6.2f} nodes {rp['tris_per_ray']:6.2f} tris (worst {rp['worst']}) | "
                  f"inside rays: {rs['nodes_per_ray']:6.2f} nodes {rs['tris_per_ray']:6.2f} tris (worst {rs['worst']}) | CPU build {rp['build_s']:.1f} s")
        a, b = res["PLOC"], res["binned SAH"]
        same = lambda x, y: np.array_equal(x.view(np.uint32), y.view(np.uint32))      # bit patterns: a miss carries the id 0xFFFFFFFF
        assert same(a[0]["hits"], b[0]["hits"]) and same(a[1]["hits"], b[1]["hits"]), "hit records must not depend on the tree"
        print(f"  PLOC / binned SAH: SAH {a[0]['sah'] / b[0]['sah']:.3f}, camera nodes {a[0]['nodes_per_ray'] / b[0]['nodes_per_ray']:.3f} tris {a[0]['tris_per_ray'] / max(b[0]['tris_per_ray'], 1e-9):.3f}, "
              f"inside nodes {a[1]['nodes_per_ray'] / b[1]['nodes_per_ray']:.3f} tris {a[1]['tris_per_ray'] / max(b[1]['tris_per_ray'], 1e-9):.3f}; identical hit records")
