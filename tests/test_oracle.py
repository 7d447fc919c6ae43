"""The CPU oracle against every known answer the reference holds for the path and against
independent restatements of the Appendix-A quirks.  No GPU needed."""
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden
from pgi_raytracing_b200 import scenes

FLT_MAX = np.finfo(np.float32).max


@pytest.fixture(scope="module")
def tri_oracle(oracle_mod):
    return oracle_mod.Oracle(scenes.single_triangle())


# ------------------------------------------------------------------------------------------------ T1 (tutorial_1)
@pytest.mark.parametrize("brute", [True, False])
def test_T1_tutorial_1_known_answer(oracle_mod, tri_oracle, brute):
    """pg1/tutorials.cpp:39-41,67-69,87-97: t=2, u=0.05, v=0.0667, primID=geomID=0, normal (0,0,1), uv (0.050, 0.933)."""
    rh = oracle_mod.make_rayhits([[0.1, 0.2, 2.0]], [[0.0, 0.0, -1.0]], tnear=np.finfo(np.float32).tiny)
    out = tri_oracle.intersect(rh, brute=brute)
    assert out["geomID"][0] == 0 and out["primID"][0] == 0
    assert out["tfar"][0] == np.float32(2.0)
    assert abs(out["u"][0] - 0.05) < 1e-7 and abs(out["v"][0] - 0.2 / 3.0) < 1e-7
    assert (out["Ng_x"][0], out["Ng_y"][0], out["Ng_z"][0]) == (0.0, 0.0, 6.0)      # (p1-p0) x (p2-p0), un-normalised
    n = tri_oracle.interpolate([0], [0], out["u"], out["v"], 0)[0]
    uv = tri_oracle.interpolate([0], [0], out["u"], out["v"], 1)[0]
    assert "normal = (%0.3f, %0.3f, %0.3f)" % tuple(n) == "normal = (0.000, 0.000, 1.000)"       # printf at tutorials.cpp:134
    assert "tex_coord = (%0.3f, %0.3f)" % tuple(uv) == "tex_coord = (0.050, 0.933)"              # printf at tutorials.cpp:135


def test_T1_miss_leaves_record_untouched(oracle_mod, tri_oracle):
    rh = oracle_mod.make_rayhits([[5.0, 5.0, 2.0]], [[0.0, 0.0, -1.0]])
    out = tri_oracle.intersect(rh, brute=True)
    assert out["geomID"][0] == 0xFFFFFFFF and out["tfar"][0] == FLT_MAX


def test_range_is_open_at_tnear_closed_at_tfar(oracle_mod, tri_oracle):
    mk = oracle_mod.make_rayhits
    assert tri_oracle.intersect(mk([[0.1, 0.2, 2.0]], [[0, 0, -1]], tnear=2.0), brute=True)["geomID"][0] == 0xFFFFFFFF
    assert tri_oracle.intersect(mk([[0.1, 0.2, 2.0]], [[0, 0, -1]], tnear=0.0, tfar=2.0), brute=True)["geomID"][0] == 0
    assert tri_oracle.intersect(mk([[0.1, 0.2, 2.0]], [[0, 0, -1]], tnear=0.0, tfar=1.999), brute=True)["geomID"][0] == 0xFFFFFFFF
    # no back-face culling, direction need not be unit length
    out = tri_oracle.intersect(mk([[0.1, 0.2, -2.0]], [[0, 0, 4.0]]), brute=True)
    assert out["geomID"][0] == 0 and out["tfar"][0] == np.float32(0.5)


# ------------------------------------------------------------------------------------------------ T2 (tutorial_2)
def _image(buf, w, bpp):
    return scenes.Image(np.ascontiguousarray(buf), w, buf.shape[0], buf.shape[1], bpp)


def test_T2_tutorial_2_known_answer(oracle_mod):
    """pg1/tutorials.cpp:173-175 on data/test4.png: prints (r = 1.000, g = 0.000, b = 0.500)."""
    sc = scenes.single_triangle()
    sc.textures = [_image(golden("test4_bgra.npy"), 64, 4)]
    o = oracle_mod.Oracle(sc)
    texel = o.texture_get_texel(0, [[(1.0 / 64) * 2.5, 0.0]])[0]
    assert "(r = %0.3f, g = %0.3f, b = %0.3f)" % tuple(texel) == "(r = 1.000, g = 0.000, b = 0.500)"


def test_get_texel_quirks(oracle_mod):
    """black on integer coordinates, for x in (0,1), on the last row; wrap column weights."""
    rng = np.random.default_rng(0)
    w, h = 7, 5
    rgb = rng.integers(1, 255, (h, w, 3), dtype=np.uint8)
    img = scenes.Image.from_rgb(rgb)
    assert img.pitch == 24 and img.pitch != w * 3      # FreeImage 4-byte row padding is honoured
    sc = scenes.single_triangle(); sc.textures = [img]
    o = oracle_mod.Oracle(sc)

    def texel(x, y):
        return o.texture_get_texel(0, [[np.float32(x) / np.float32(w), np.float32(y) / np.float32(h)]])[0]

    def px(x, y):   # Color3f{b, g, r}: .r carries the blue byte
        return rgb[y, x, ::-1].astype(np.float32) / np.float32(255)

    assert np.all(texel(3.0, 2.5) == 0)            # integer x -> x1 == x2
    assert np.all(texel(2.5, 3.0) == 0)            # integer y
    assert np.all(texel(0.5, 2.5) == 0)            # x in (0,1): x1 = 0+1 == x2 = 1
    assert np.all(texel(2.5, 4.5) == 0)            # last row: y2 = tmp_y1
    got = texel(2.25, 1.75)
    exp = (px(2, 1) * 0.75 + px(3, 1) * 0.25) * 0.25 + (px(2, 2) * 0.75 + px(3, 2) * 0.25) * 0.75
    assert np.allclose(got, exp, atol=1e-6)
    got = texel(6.5, 1.5)                          # wrap column: x1 = 6, x2 = 0 -> weights (0-6.5)/(0-6), (6.5-6)/(0-6)
    q1, q2 = np.float32((0 - 6.5) / (0 - 6)), np.float32((6.5 - 6) / (0 - 6))
    exp = (px(6, 1) * q1 + px(0, 1) * q2) * 0.5 + (px(6, 2) * q1 + px(0, 2) * q2) * 0.5
    assert np.allclose(got, exp, atol=1e-6)


# ------------------------------------------------------------------------------------------------ colour maths
def _expand(u):
    return 0.0 if u <= 0 else 1.0 if u >= 1 else (u / 12.92 if u <= 0.04045 else ((u + 0.055) / 1.055) ** 2.4)


def _compress(u):
    return 0.0 if u <= 0 else 1.0 if u >= 1 else (12.92 * u if u <= 0.0031308 else 1.00 * u ** (1 / 2.4) - 0.055)


def test_mix_srgb_matches_independent_restatement(oracle_mod):
    """A-6: out.r = C(a E(c0.b) + (1-a) E(c1.b)), out.g likewise on g, out.b on r; alpha 0.1; mis-scaled encode."""
    rng = np.random.default_rng(1)
    n = 2000
    c0 = rng.uniform(-0.2, 1.3, (n, 4)).astype(np.float32); c1 = rng.uniform(-0.2, 1.3, (n, 4)).astype(np.float32)
    a = rng.uniform(0, 1, n).astype(np.float32)
    out = oracle_mod.Oracle().mix_srgb(c0, c1, a)
    for i in range(0, n, 7):
        for dst, src in ((0, 2), (1, 1), (2, 0)):
            lin = np.float32(a[i] * np.float32(_expand(float(c0[i, src]))) + np.float32(1 - a[i]) * np.float32(_expand(float(c1[i, src]))))
            assert abs(out[i, dst] - _compress(float(lin))) < 2e-6, (i, dst)
        assert out[i, 3] == np.float32(0.1)
    # the encode tops out at 1*u^(1/2.4) - 0.055 (< 0.945) before jumping to 1
    top = oracle_mod.Oracle().mix_srgb([[0.999, 0.999, 0.999, 1]], [[0.999, 0.999, 0.999, 1]], [0.5])[0]
    assert 0.94 < top[0] < 0.946


def test_gamma_is_c_to_the_2g_with_channel_swap(oracle_mod):
    o = oracle_mod.Oracle()
    c = np.array([[0.2, 0.5, 0.8, 1.0], [-0.1, 0.3, 0.0, 1.0]], np.float32)
    out = o.gamma(c, 0.5)
    assert np.allclose(out[0, :3], [0.8, 0.5, 0.2], atol=1e-6) and out[0, 3] == 1.0      # {b', g', r', 1}
    assert np.isnan(out[1, 2]) and out[1, 0] == 0.0                                      # negative mean -> NaN (A-11)
    out = o.gamma(c[:1], 0.25)
    assert np.allclose(out[0, :3], np.sqrt([0.8, 0.5, 0.2]), atol=1e-6)


def test_secondary_rays(oracle_mod):
    o = oracle_mod.Oracle()
    d = np.array([1.0, 0.0, -1.0]) / math.sqrt(2); n = np.array([0.0, 0.0, 1.0]); hp = np.array([1.0, 2.0, 3.0])
    item = np.concatenate([d, n, hp, [1.000293, 1.5]]).astype(np.float32)[None]
    refl = o.secondary_rays(item, False)[0]
    assert np.allclose(refl[:3], hp) and refl[3] == np.float32(0.01) and refl[8] == FLT_MAX
    assert np.allclose(refl[4:7], [d[0], 0.0, -d[2]], atol=1e-6) and refl[7] == np.float32(1.000293)     # time = n1
    refr = o.secondary_rays(item, True)[0]
    eta = 1.000293 / 1.5
    sin_t = eta * d[0]
    assert np.allclose(refr[4:7], [sin_t, 0.0, -math.sqrt(1 - sin_t ** 2)], atol=1e-6) and refr[7] == np.float32(1.5)
    # total internal reflection: glass -> air at a grazing angle gives NaN (pg1/raytracer.cpp:188,309)
    g = np.array([math.sin(1.2), 0.0, -math.cos(1.2)])
    tir = o.secondary_rays(np.concatenate([g, n, hp, [1.5, 1.000293]]).astype(np.float32)[None], True)[0]
    assert np.isnan(tir[4]) and np.isnan(tir[6])


def test_rng_spec(oracle_mod):
    """The counter-based RNG both sides implement (replaces the clock-seeded mt19937): restated in Python."""
    def mix(h):
        h &= 0xFFFFFFFF; h ^= h >> 16; h = (h * 0x7feb352d) & 0xFFFFFFFF; h ^= h >> 15; h = (h * 0x846ca68b) & 0xFFFFFFFF; h ^= h >> 16
        return h
    o = oracle_mod.Oracle()
    for seed, px, s, dim in ((1, 0, 0, 0), (1, 12345, 8, 3), (77, 2073599, 63, 2), (0xFFFFFFFF, 7, 1, 1)):
        h = mix((seed + 0x9E3779B9 * (px + 1)) & 0xFFFFFFFF)
        h = mix(h ^ ((s * 4 + dim + 0x85EBCA6B) & 0xFFFFFFFF))
        assert o.rng_u01(seed, px, s, dim) == np.float32((h >> 8) / 16777216.0)


# ------------------------------------------------------------------------------------------------ camera
def test_camera_matches_reference_construction(oracle_mod):
    """PinHoleCamera.cpp:5-29 with the default view (tutorials.cpp:186-191)."""
    sc = scenes.single_triangle()
    o = oracle_mod.Oracle(sc)
    k = o.camera_constants()
    fov = sc.camera.fov_y
    assert abs(k[0] - 480 / (2 * math.tan(fov / 2))) < 1e-3
    M = k[1:].reshape(3, 3)
    z = np.array([-140.0, -175.0, 40.0]); z /= np.linalg.norm(z)
    assert np.allclose(M[:, 2], z, atol=1e-6) and np.allclose(M.T @ M, np.eye(3), atol=1e-6)
    # centre pixel, no lens shift: the ray goes from `from` towards `at`
    rays = o.primary_rays(oracle_mod.make_params(sampling_width=1, jitter=0, aperture=0.0))
    c = rays[240 * 640 + 320]
    assert np.allclose(c[:3], [-140, -175, 80]) and c[3] == np.float32(0.01) and c[7] == np.float32(1.000293)
    assert np.allclose(c[4:7], -z, atol=1e-6)
    # camera obscura overload: tnear 0.001 (PinHoleCamera.cpp:54)
    rays = o.primary_rays(oracle_mod.make_params(sampling_width=1, jitter=0, camera_mode=1))
    assert rays[0][3] == np.float32(0.001)
    # y flips: the top row looks above the centre
    assert rays[320][6] > rays[479 * 640 + 320][6]


def test_jitter_and_lens_ranges(oracle_mod):
    sc = scenes.single_triangle(); sc.camera = scenes.Camera(16, 12, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
    o = oracle_mod.Oracle(sc)
    r = o.primary_rays(oracle_mod.make_params(sampling_width=3, jitter=1, aperture=5.0, seed=4))
    org = r[:, :3] - np.array([-140, -175, 80], np.float32)
    assert 1.0 < np.abs(org).max() < 3.6 and len(np.unique(r[:, 4])) > 1000          # square aperture of side 5
    r2 = o.primary_rays(oracle_mod.make_params(sampling_width=3, jitter=1, aperture=5.0, seed=5))
    assert not np.array_equal(r, r2)
    assert np.array_equal(r, o.primary_rays(oracle_mod.make_params(sampling_width=3, jitter=1, aperture=5.0, seed=4)))


# ------------------------------------------------------------------------------------------------ acceleration structure
def test_bvh_equals_brute_force(oracle_mod, cornell):
    o = oracle_mod.Oracle(cornell)
    rng = np.random.default_rng(2)
    n = 4000
    org = rng.uniform(-150, 150, (n, 3)).astype(np.float32); org[:, 2] = np.abs(org[:, 2])
    target = rng.uniform(-70, 70, (n, 3)).astype(np.float32); target[:, 2] = rng.uniform(-5, 60, n)
    d = (target - org) * rng.uniform(0.01, 3.0, (n, 1)).astype(np.float32)      # directions are not unit length
    rh = oracle_mod.make_rayhits(org, d, tnear=0.01)
    a, b = o.intersect(rh, brute=True), o.intersect(rh, brute=False)
    assert (a["geomID"] != 0xFFFFFFFF).mean() > 0.3
    for f in ("tfar", "u", "v", "geomID", "primID"):
        assert np.array_equal(a[f], b[f]), f


def test_soup_bvh_equals_brute_force(oracle_mod):
    sc = scenes.triangle_soup(3000, seed=3, resolution=(64, 36))
    o = oracle_mod.Oracle(sc)
    rays = o.primary_rays(oracle_mod.make_params(sampling_width=1, jitter=0, aperture=0.0))
    rh = oracle_mod.make_rayhits(rays[:, :3], rays[:, 4:7], tnear=0.01)
    a, b = o.intersect(rh, brute=True), o.intersect(rh, brute=False)
    for f in ("tfar", "geomID", "primID"):
        assert np.array_equal(a[f], b[f]), f


# ------------------------------------------------------------------------------------------------ golden renders
@pytest.mark.parametrize("tag", ["c1", "dof3x3", "lambert", "pinhole_depth3"])
def test_golden_renders_reproduce(oracle_mod, cornell, tag):
    """The committed fixtures were rendered with brute-force intersection; the BVH path must reproduce them bit for bit."""
    z = np.load(os.path.join(GOLDEN, f"cornell_{tag}.npz"))
    p = json.loads(str(z["params"]))
    rgba, g, pr, st = oracle_mod.Oracle(cornell).render(oracle_mod.make_params(**p), threads=2)
    assert np.array_equal(rgba, z["rgba"], equal_nan=True)
    assert np.array_equal(g, z["geom"]) and np.array_equal(pr, z["prim"])
    assert [st["primary"], st["shadow"], st["reflection"], st["refraction"]] == list(z["rays"])


def test_render_is_thread_count_independent(oracle_mod, cornell):
    o = oracle_mod.Oracle(cornell)
    p = oracle_mod.make_params(seed=11)
    a = o.render(p, threads=1)[0]; b = o.render(p, threads=4)[0]
    assert np.array_equal(a, b, equal_nan=True)


def test_depth_cutoff_and_ray_accounting(oracle_mod, cornell):
    """A-3: level >= max_depth returns black on a hit only; every get_ray_hit call is counted once."""
    o = oracle_mod.Oracle(cornell)
    base = dict(sampling_width=1, jitter=0, aperture=0.0)
    s0 = o.render(oracle_mod.make_params(max_depth=0, **base))[3]
    assert s0["shadow"] == 0 and s0["reflection"] == 0 and s0["refraction"] == 0 and s0["primary"] == 64 * 48
    s1 = o.render(oracle_mod.make_params(max_depth=1, **base))[3]
    s7 = o.render(oracle_mod.make_params(max_depth=7, **base))[3]
    assert 0 < s1["reflection"] < s7["reflection"] and s7["refraction"] <= s7["reflection"]
    img0 = o.render(oracle_mod.make_params(max_depth=0, **base))[0]
    hit = o.render(oracle_mod.make_params(max_depth=0, **base))[1] != 0xFFFFFFFF
    assert np.all(img0[hit][:, :3] == 0) and np.any(img0[~hit][:, :3] > 0)


def test_mtl_parse_pin():
    """data/6887_allied_avenger.mtl as LoadMTL reads it (fixture made with glibc sscanf on the reference's file):
    5 materials, shader 3 x4 + shader 4 x1, Ks = (1.0, 0.8, 0.8) from the malformed 'Ks 1.0. 1.0 1.0'."""
    with open(os.path.join(GOLDEN, "avenger_mtl_parse.json")) as f:
        ref = json.load(f)
    ours = scenes.avenger_materials()
    assert len(ref) == len(ours) == 5
    assert [m["type"] for m in ref] == [3, 4, 3, 3, 3]
    for r, m in zip(ref, ours):
        assert r["name"] == m.name and r["type"] == m.type
        assert np.allclose(r["specular"], [1.0, 0.8, 0.8]) and np.allclose(m.specular, r["specular"])
        assert np.allclose(m.diffuse, r["diffuse"]) and np.float32(m.ior) == np.float32(r["ior"]) and m.shininess == r["shininess"]
        assert (r["map_Kd"] or "") == m.map_kd


def test_path_tracing_mode_of_the_oracle_is_deterministic_and_adds_light(oracle_mod):
    """shader_mode 3 (README to-do, spec in include/pgrt.h): same seed -> same frame; another seed -> other bounces; dielectrics
    still branch; every diffuse hit above the depth cut-off casts one bounce; the environment adds light."""
    oracle = oracle_mod
    sc = scenes.cornell_like()
    o = oracle.Oracle(sc)
    kw = dict(sampling_width=1, seed=3, max_depth=4)
    a, _, _, sa = o.render(oracle.make_params(shader_mode=3, **kw), want_ids=False)
    a2, _, _, _ = o.render(oracle.make_params(shader_mode=3, **kw), want_ids=False)
    b, _, _, sb = o.render(oracle.make_params(shader_mode=0, **kw), want_ids=False)
    c, _, _, _ = o.render(oracle.make_params(shader_mode=3, **dict(kw, seed=4)), want_ids=False)
    assert np.array_equal(a, a2, equal_nan=True) and not np.array_equal(a, c, equal_nan=True)
    assert sa["primary"] == sb["primary"] and sa["refraction"] > 0 and sa["reflection"] > sb["reflection"]
    assert np.nanmean(a[..., :3]) > np.nanmean(b[..., :3])


def test_path_bounce_is_cosine_weighted(tri_oracle):
    """The bounce of shader_mode 3: unit vectors in the hemisphere of n, density proportional to cos(theta) -- E[cos] = 2/3,
    E[cos^2] = 1/2, azimuth uniform -- and a function of (hit point bits, level, seed) only."""
    rng = np.random.default_rng(5)
    hits = rng.uniform(-100, 100, (200_000, 3)).astype(np.float32)
    n = np.array([0.0, 0.6, 0.8], np.float32)
    d = tri_oracle.path_bounce(n, hits, seed=7, level=2)
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=2e-6)
    c = d @ n
    assert c.min() >= -1e-6
    assert abs(c.mean() - 2.0 / 3.0) < 3e-3 and abs((c * c).mean() - 0.5) < 3e-3
    t = np.cross(n, [1.0, 0.0, 0.0]); t /= np.linalg.norm(t); b = np.cross(n, t)
    phi = np.arctan2(d @ b, d @ t)
    hist, _ = np.histogram(phi, bins=16, range=(-np.pi, np.pi))
    assert np.abs(hist / hist.mean() - 1.0).max() < 0.04
    assert np.array_equal(d, tri_oracle.path_bounce(n, hits, seed=7, level=2))
    assert not np.array_equal(d, tri_oracle.path_bounce(n, hits, seed=8, level=2))
    assert not np.array_equal(d, tri_oracle.path_bounce(n, hits, seed=7, level=3))


def test_bench_clock_sampler_degrades_without_a_gpu():
    """bench.py's sampler must never take the run down: without NVML / nvidia-smi it reports that and nothing else."""
    import importlib, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    bench = importlib.import_module("bench")
    s = bench.ClockSampler(0); s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}


# ------------------------------------------------------------------------------------------------ the oracle's own BVH
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



def test_oracle_bvh_equals_oracle_brute_force_on_fuzzed_input(oracle_mod):
    """The checker checked: the oracle's BVH traversal (what its renders use) against its brute-force loop on pathological
    scenes and rays (axis-aligned sheets hit on their border, slivers, duplicates, axis-parallel rays ...).  Its box test
    once culled a handful of border hits in 150 000 rays; it now carries the same per-axis pads as the CUDA path."""
    for seed in range(0, 30):
        rng = np.random.default_rng(seed)
        kind = seed % 5; n = int(rng.choice([1, 2, 3, 7, 33, 200, 1500]))
        pos = _fuzz_scene(rng, kind, n); rays = _fuzz_rays(rng, pos, 400)
        nrm = np.tile(np.array([0, 0, 1], np.float32), (n, 3, 1)); uv = np.zeros((n, 3, 2), np.float32)
        sc = scenes.Scene("fuzz", [scenes.Mesh("m", pos.reshape(n, 3, 3), nrm, uv, 0)], [scenes.Material("m")], camera=scenes.Camera(8, 8))
        o = oracle_mod.Oracle(sc)
        rh = oracle_mod.make_rayhits(rays[:, :3], rays[:, 4:7])
        rh["tnear"] = rays[:, 3]; rh["tfar"] = rays[:, 7]
        a = o.intersect(rh, brute=False); b = o.intersect(rh, brute=True)
        for f in ("tfar", "u", "v", "geomID", "primID"):
            assert np.array_equal(a[f], b[f]), (seed, kind, n, f)
