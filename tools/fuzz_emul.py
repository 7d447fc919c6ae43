"""CPU fuzzing of the wide-node encoder + traversal (tests/emul) against brute force: python tools/fuzz_emul.py SEED0 SEED1
Scenes: soup / integer-grid sheets / size mixes far from the origin / duplicates + degenerates / slivers; rays aimed at edges and
vertices, zero components, axis-parallel, origins on the surface, empty intervals, 1e-12-scaled directions."""
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import test_bvh8_emul as T
lib = T.emul.__wrapped__() if hasattr(T.emul, '__wrapped__') else None
if lib is None:
    lib = C.CDLL(T.LIB)
    lib.emul_build.restype = C.c_void_p; lib.emul_build.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int]
    lib.emul_free.argtypes = [C.c_void_p]
    lib.emul_depth.restype = C.c_uint32; lib.emul_depth.argtypes = [C.c_void_p]
    lib.emul_check.restype = C.c_uint64; lib.emul_check.argtypes = [C.c_void_p]
    lib.emul_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
FMAX = np.finfo(np.float32).max
def gen_scene(rng, kind, n):
    if kind == 0:   # uniform soup
        c = rng.uniform(-100, 100, (n, 1, 3)); p = c + rng.uniform(-3, 3, (n, 3, 3))
    elif kind == 1: # axis-aligned grid sheets (flat boxes), exact integer coordinates
        g = int(np.sqrt(n / 2)) + 1
        xs, ys = np.meshgrid(np.arange(g), np.arange(g))
        a = np.stack([xs, ys, np.zeros_like(xs)], -1).reshape(-1, 3).astype(float)
        t1 = np.stack([a, a + [1, 0, 0], a + [1, 1, 0]], 1); t2 = np.stack([a, a + [1, 1, 0], a + [0, 1, 0]], 1)
        p = np.concatenate([t1, t2])[:n]
        if rng.random() < 0.5: p = p[..., [2, 0, 1]]
    elif kind == 2: # huge + tiny mix, far from origin
        c = rng.uniform(-1, 1, (n, 1, 3)) * 10.0 ** rng.uniform(-2, 4, (n, 1, 1)) + 5000.0
        p = c + rng.normal(size=(n, 3, 3)) * 10.0 ** rng.uniform(-3, 2, (n, 1, 1))
    elif kind == 3: # many duplicates + degenerate
        base = rng.uniform(-10, 10, (max(n // 8, 1), 3, 3))
        p = base[rng.integers(0, base.shape[0], n)]
        deg = rng.random(n) < 0.1; p[deg, 2] = p[deg, 1]
    else:           # long thin slivers along a diagonal
        t = rng.uniform(0, 1, (n, 1, 1)); c = t * np.array([100.0, 100.0, 100.0])
        p = c + rng.normal(size=(n, 3, 3)) * np.array([20.0, 0.01, 0.01])
    return np.ascontiguousarray(p.reshape(n, 9).astype(np.float32))
def gen_rays(rng, pos, m):
    P = pos.reshape(-1, 3); lo, hi = P.min(0), P.max(0); ext = np.maximum(hi - lo, 1e-3)
    org = rng.uniform(lo - ext, hi + ext, (m, 3)); tri = pos[rng.integers(0, pos.shape[0], m)].reshape(m, 3, 3)
    w = rng.dirichlet((1, 1, 1), m)[:, :, None]; tgt = (tri * w).sum(1)
    edge = rng.random(m) < 0.2; tgt[edge] = tri[edge, 0] * 0.5 + tri[edge, 1] * 0.5     # aim at edges
    vert = rng.random(m) < 0.1; tgt[vert] = tri[vert, 2]                                 # and vertices
    d = (tgt - org) * rng.uniform(0.01, 5, (m, 1))
    ax = rng.random(m) < 0.15; k = rng.integers(0, 3, m); d[ax, k[ax]] = 0.0            # zero components
    ax2 = rng.random(m) < 0.05; d[ax2] = 0; d[ax2, k[ax2]] = rng.choice([-1.0, 1.0], ax2.sum())
    inside = rng.random(m) < 0.2; org[inside] = tgt[inside]; d[inside] = rng.normal(size=(inside.sum(), 3))
    rays = np.zeros((m, 8), np.float32); rays[:, :3] = org; rays[:, 4:7] = d
    rays[:, 3] = rng.choice([0.0, 1e-3, 0.01], m); rays[:, 7] = FMAX
    b = rng.random(m) < 0.2; rays[b, 7] = rng.uniform(0.1, 2.0, b.sum())
    neg = rng.random(m) < 0.02; rays[neg, 3] = 5.0; rays[neg, 7] = 1.0                   # empty interval
    tiny = rng.random(m) < 0.03; rays[tiny, 4:7] *= 1e-12                                # near-denormal directions
    return rays
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(seed)
    kind = seed % 5; n = int(rng.choice([1, 2, 3, 7, 33, 200, 1500, 6000]))
    pos = gen_scene(rng, kind, n); rays = gen_rays(rng, pos, 600)
    ref = None
    for layout in (0, 1):
        for builder in (0, 1):
            h = lib.emul_build(pos.ctypes.data, pos.shape[0], builder, layout)
            try:
                if layout == 0 and lib.emul_check(h) != 0: print("STRUCT", seed, kind, n, builder); bad += 1
                if lib.emul_depth(h) > 38: print("DEPTH", seed, kind, n, builder, lib.emul_depth(h)); bad += 1
                b, _ = T.trace(lib, h, rays, 1)
                for ww in (0, 2):
                    a, _ = T.trace(lib, h, rays, ww)
                    if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
                        idx = np.nonzero((a.view(np.uint32) != b.view(np.uint32)).any(1))[0]
                        print("MISMATCH seed", seed, "kind", kind, "n", n, "layout", layout, "builder", builder, "ww", ww, "rays", idx[:5], a[idx[0]], b[idx[0]], rays[idx[0]])
                        bad += 1
            finally:
                lib.emul_free(h)
print("done", sys.argv[1], sys.argv[2], "bad", bad)
