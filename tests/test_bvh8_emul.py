"""The wide-node encoder and the traversal of csrc/bvh8.cuh + csrc/traverse.cuh, compiled with g++ and run on the CPU.

The GPU build kernels call the same PG_HD functions once per wide node, so this covers the slot assignment, the 8-bit
quantisation (conservative decode), the meta / imask encoding and the stack traversal before a GPU is involved.
Expected: bit-identical hit records to a brute-force loop over the same triangle test, for any binary tree.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pgi_raytracing_b200 import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "bvh8_emul.cpp")
LIB = os.path.join(HERE, "emul", "libbvh8_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "pgi_raytracing_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("bvh8.cuh", "traverse.cuh", "common.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-mavx2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-I/usr/local/cuda/include", "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.emul_build.restype = C.c_void_p; lib.emul_build.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int]
    lib.emul_free.argtypes = [C.c_void_p]
    lib.emul_nodes.restype = C.c_uint32; lib.emul_nodes.argtypes = [C.c_void_p]
    lib.emul_depth.restype = C.c_uint32; lib.emul_depth.argtypes = [C.c_void_p]
    lib.emul_sah.restype = C.c_double; lib.emul_sah.argtypes = [C.c_void_p]
    lib.emul_check.restype = C.c_uint64; lib.emul_check.argtypes = [C.c_void_p]
    lib.emul_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
    return lib


def scene_pos(sc):
    return np.ascontiguousarray(np.concatenate([m.pos.reshape(-1, 9) for m in sc.meshes]).astype(np.float32))


def random_rays(pos, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = pos.reshape(-1, 3).min(0), pos.reshape(-1, 3).max(0)
    ext = np.maximum(hi - lo, 1.0)
    org = rng.uniform(lo - 0.5 * ext, hi + 0.5 * ext, (n, 3))
    tgt = rng.uniform(lo, hi, (n, 3))
    pick = rng.integers(0, pos.shape[0], n // 2)                                          # half the rays aim at a triangle
    w = rng.dirichlet((1.0, 1.0, 1.0), n // 2)[:, :, None]
    tgt[n // 2:n // 2 + pick.shape[0]] = (pos[pick].reshape(-1, 3, 3) * w).sum(1)
    d = (tgt - org) * rng.uniform(0.05, 3.0, (n, 1))
    d[:50, 0] = 0.0; d[50:100, 1] = 0.0; d[100:150, 2] = 0.0; d[150:160, :2] = 0.0      # axis-parallel rays
    org[160:400] = tgt[160:400]                                                           # origins inside the scene (secondary-like)
    d[160:400] = rng.normal(size=(240, 3))
    far = slice(500, min(700, n)); nf = max(0, min(700, n) - 500)                         # far origins: plane*idir - O*idir cancels
    org[far] = tgt[far] + rng.normal(size=(nf, 3)) * ext * 40.0
    d[far] = (tgt[far] - org[far]) * rng.uniform(0.5, 2.0, (nf, 1))
    rays = np.zeros((n, 8), np.float32)
    rays[:, :3] = org; rays[:, 3] = 0.01; rays[:, 4:7] = d; rays[:, 7] = np.finfo(np.float32).max
    rays[400:500, 7] = rng.uniform(0.2, 1.5, 100)                                         # bounded tfar (shadow-like)
    return rays


def trace(lib, h, rays, brute):
    out = np.zeros((rays.shape[0], 4), np.float32); st = np.zeros((rays.shape[0], 2), np.uint32)
    lib.emul_trace(h, rays.ctypes.data, rays.shape[0], out.ctypes.data, st.ctypes.data, int(brute))
    return out, st


CASES = {
    "single": lambda: scenes.single_triangle(),
    "cornell": lambda: scenes.cornell_like(),
    "soup3000": lambda: scenes.triangle_soup(3000, seed=5, resolution=(32, 32)),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("builder", [0, 1])
@pytest.mark.parametrize("layout", [0, 1])          # 0 = 80-B quantised nodes, 1 = 240-B float planes
def test_wide_tree_structure_and_hits(emul, name, builder, layout):
    pos = scene_pos(CASES[name]())
    h = emul.emul_build(pos.ctypes.data, pos.shape[0], builder, layout)
    try:
        assert emul.emul_check(h) == 0                       # every triangle once, decoded boxes conservative, meta consistent
        assert emul.emul_depth(h) <= 32                      # PGRT_STACK8 = 40 entries
        rays = random_rays(pos, 1500 if pos.shape[0] > 100 else 600, seed=3)
        a, st = trace(emul, h, rays, 0); b, _ = trace(emul, h, rays, 1)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        c, st_ww = trace(emul, h, rays, 2)                   # while-while loop shape: same hits, same work
        assert np.array_equal(c.view(np.uint32), b.view(np.uint32)) and st_ww[:, 0].sum() <= st[:, 0].sum() * 1.05 + 8
        hit = b.view(np.uint32)[:, 3] != 0xFFFFFFFF
        assert hit.mean() > 0.1
        if pos.shape[0] > 1000:                              # the tree prunes: far fewer triangle tests than brute force
            assert st[:, 1].mean() < 0.05 * pos.shape[0]
    finally:
        emul.emul_free(h)


def test_duplicates_and_degenerates(emul):
    """Coincident duplicates (ties -> lowest flat id), zero-area triangles and a flat (axis-aligned) sheet."""
    rng = np.random.default_rng(1)
    base = scene_pos(scenes.triangle_soup(400, seed=2, resolution=(8, 8)))
    sheet = np.zeros((200, 9), np.float32)
    sheet[:, [0, 3, 6]] = rng.uniform(-50, 50, (200, 3)); sheet[:, [1, 4, 7]] = rng.uniform(-50, 50, (200, 3)); sheet[:, [2, 5, 8]] = 7.0
    degenerate = np.repeat(rng.uniform(-50, 50, (20, 3)).astype(np.float32), 3, axis=0).reshape(20, 9)
    pos = np.ascontiguousarray(np.concatenate([base, base[:100], sheet, degenerate]).astype(np.float32))
    for layout in (0, 1):
        h = emul.emul_build(pos.ctypes.data, pos.shape[0], 0, layout)
        _check_degenerates(emul, h, pos)


def test_runs_of_ties_do_not_build_a_chain(emul):
    """Thousands of coincident triangles, and a regular row of identical ones: every merge candidate has the same area.
    The neighbour preference (nearer index, then the buddy i^1) must pair them up; picking the lowest index instead
    merged one pair per pass and produced a tree thousands of levels deep (beyond the traversal stack)."""
    one = np.array([[0, 0, 0, 4, 0, 0, 0, 4, 1]], np.float32)
    row = np.repeat(one, 2500, axis=0); row[:, [0, 3, 6]] += np.arange(2500, dtype=np.float32)[:, None] * 8.0
    for pos in (np.repeat(one, 3000, axis=0), row):
        pos = np.ascontiguousarray(pos)
        h = emul.emul_build(pos.ctypes.data, pos.shape[0], 0, 1)
        try:
            assert emul.emul_check(h) == 0
            assert emul.emul_depth(h) <= 12, emul.emul_depth(h)
            rays = random_rays(pos, 800, seed=6)
            a, _ = trace(emul, h, rays, False); b, _ = trace(emul, h, rays, True)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        finally:
            emul.emul_free(h)


def _check_degenerates(emul, h, pos):
    try:
        assert emul.emul_check(h) == 0
        rays = random_rays(pos, 1200, seed=4)
        a, _ = trace(emul, h, rays, False); b, _ = trace(emul, h, rays, True)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        ids = a.view(np.uint32)[:, 3]
        assert not np.any((ids >= 400) & (ids < 500))        # the duplicate copy never wins a tie
    finally:
        emul.emul_free(h)


def test_ploc_tree_is_better_than_median_split(emul):
    pos = scene_pos(scenes.cornell_like())
    h0 = emul.emul_build(pos.ctypes.data, pos.shape[0], 0, 0); h1 = emul.emul_build(pos.ctypes.data, pos.shape[0], 1, 0)
    try:
        assert emul.emul_sah(h0) < emul.emul_sah(h1)
    finally:
        emul.emul_free(h0); emul.emul_free(h1)


# ------------------------------------------------------------------------------------------------ fuzzing
FMAX = np.finfo(np.float32).max
def _fuzz_scene(rng, kind, n):
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
def _fuzz_rays(rng, pos, m):
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


@pytest.mark.parametrize("block", range(4))
def test_fuzzed_scenes_and_rays_match_brute_force(emul, block):
    """Pathological geometry (axis-aligned sheets on an integer grid, 1e-2..1e4 size mixes far from the origin, duplicates,
    slivers) x pathological rays (aimed at edges and vertices, zero components, axis-parallel, origins on the surface, empty
    intervals, 1e-12-scaled directions): both layouts, both builders, both loop shapes against brute force, bit for bit.
    This is how the missing absolute pad of the quantised layout was found (hits on the border of axis-aligned sheets by
    rays nearly parallel to an axis).  Directions whose components are ALL below 2^-80 are outside the supported domain
    (the reciprocal is clamped there, as Embree clamps at 1e-18)."""
    for seed in range(10 * block, 10 * block + 10):
        rng = np.random.default_rng(seed)
        kind = seed % 5; n = int(rng.choice([1, 2, 3, 7, 33, 200, 1500]))
        pos = _fuzz_scene(rng, kind, n); rays = _fuzz_rays(rng, pos, 400)
        for layout in (0, 1):
            for builder in (0, 1):
                h = emul.emul_build(pos.ctypes.data, pos.shape[0], builder, layout)
                try:
                    assert emul.emul_check(h) == 0 and emul.emul_depth(h) <= 38
                    b, _ = trace(emul, h, rays, 1)
                    for ww in (0, 2):
                        a, _ = trace(emul, h, rays, ww)
                        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (seed, kind, n, layout, builder, ww)
                finally:
                    emul.emul_free(h)


def test_product_traversal_on_cpu_equals_the_oracle(emul, oracle_mod):
    """The CUDA path's own tri_test + wide-BVH traversal (traverse.cuh compiled by g++) against the ORACLE's brute force on
    the fuzzed scenes and rays: hit / miss, t, u, v bit for bit.  (Flat triangle ids differ by the Morton permutation, so
    they are compared through the hit record only.)  The GPU repeats this through the ABI; here it needs no GPU."""
    for seed in range(40, 64):
        rng = np.random.default_rng(seed)
        kind = seed % 5; n = int(rng.choice([1, 2, 3, 7, 33, 200, 1500]))
        pos = _fuzz_scene(rng, kind, n); rays = _fuzz_rays(rng, pos, 400)
        nrm = np.tile(np.array([0, 0, 1], np.float32), (n, 3, 1)); uv = np.zeros((n, 3, 2), np.float32)
        sc = scenes.Scene("fuzz", [scenes.Mesh("m", pos.reshape(n, 3, 3), nrm, uv, 0)], [scenes.Material("m")], camera=scenes.Camera(8, 8))
        o = oracle_mod.Oracle(sc)
        rh = oracle_mod.make_rayhits(rays[:, :3], rays[:, 4:7])
        rh["tnear"] = rays[:, 3]; rh["tfar"] = rays[:, 7]
        b = o.intersect(rh, brute=True)
        hit_o = b["geomID"] != 0xFFFFFFFF
        for layout in (0, 1):
            h = emul.emul_build(pos.ctypes.data, n, 0, layout)
            try:
                e, _ = trace(emul, h, rays, 0)
            finally:
                emul.emul_free(h)
            hit_e = e.view(np.uint32)[:, 3] != 0xFFFFFFFF
            assert np.array_equal(hit_o, hit_e), (seed, kind, n, layout)
            for f, col in (("tfar", 0), ("u", 1), ("v", 2)):
                assert np.array_equal(b[f][hit_o].view(np.uint32), e[hit_o, col].view(np.uint32)), (seed, kind, n, layout, f)
