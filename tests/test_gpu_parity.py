"""Parity of the CUDA path with the CPU oracle, through the C ABI (``libpgrt_b200.so``), on the B200.

Bars (BASELINE.json north_star / SURVEY.md section 8d): primary (geomID, primID) equal on >= 99.9 % of pixels; 8-bit
image within 2 LSB on >= 99.5 % of pixels and PSNR >= 45 dB; ray counts within 0.1 %; intersection records
(t, u, v, ids) bit-exact; integer / index work bit-exact.  Tolerances for floating point are written at each test.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden
from pgi_raytracing_b200 import scenes

pytestmark = pytest.mark.gpu
INVALID = 0xFFFFFFFF


@pytest.fixture(scope="module")
def P():
    import pgi_raytracing_b200 as p
    return p


@pytest.fixture(scope="module")
def cornell_pair(P, oracle_mod, cornell):
    return P.raytracer_for(cornell), oracle_mod.Oracle(cornell)


@pytest.fixture(scope="module")
def avenger(P, oracle_mod):
    sc = scenes.avenger_proxy()
    return sc, P.raytracer_for(sc), oracle_mod.Oracle(sc)


def image_bars(P, ref, img):
    a, b = P.to_srgb8(ref).astype(int), P.to_srgb8(img).astype(int)
    d = np.abs(a - b).max(axis=-1)
    mse = np.mean((a - b) ** 2.0)
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    return float(np.mean(d <= 2)), psnr


# ------------------------------------------------------------------------------------------------ known answers
def test_T1_through_the_abi(P, oracle_mod):
    """pg1/tutorials.cpp:142-167 through pgrt_intersect / pgrt_interpolate."""
    rt = P.raytracer_for(scenes.single_triangle())
    rh = oracle_mod.make_rayhits([[0.1, 0.2, 2.0]], [[0.0, 0.0, -1.0]], tnear=np.finfo(np.float32).tiny)
    out = rt.intersect(rh)
    assert out["geomID"][0] == 0 and out["primID"][0] == 0 and out["tfar"][0] == np.float32(2.0)
    assert (out["Ng_x"][0], out["Ng_y"][0], out["Ng_z"][0]) == (0.0, 0.0, 6.0)
    n = rt.interpolate([0], [0], out["u"], out["v"], 0)[0]; uv = rt.interpolate([0], [0], out["u"], out["v"], 1)[0]
    assert "normal = (%0.3f, %0.3f, %0.3f)" % tuple(n) == "normal = (0.000, 0.000, 1.000)"
    assert "tex_coord = (%0.3f, %0.3f)" % tuple(uv) == "tex_coord = (0.050, 0.933)"
    miss = rt.intersect(oracle_mod.make_rayhits([[5.0, 5.0, 2.0]], [[0.0, 0.0, -1.0]]))
    assert miss["geomID"][0] == INVALID and miss["tfar"][0] == np.finfo(np.float32).max      # untouched on a miss


def test_T2_through_the_abi(P):
    """pg1/tutorials.cpp:170-178 on the reference's data/test4.png fixture."""
    sc = scenes.single_triangle()
    buf = golden("test4_bgra.npy")
    sc.textures = [scenes.Image(np.ascontiguousarray(buf), 64, buf.shape[0], buf.shape[1], 4)]
    rt = P.raytracer_for(sc)
    texel = rt.texture_get_texel(0, [[(1.0 / 64) * 2.5, 0.0]])[0]
    assert "(r = %0.3f, g = %0.3f, b = %0.3f)" % tuple(texel) == "(r = 1.000, g = 0.000, b = 0.500)"


def test_empty_scene_renders_the_env_map(P, oracle_mod):
    sc = scenes.Scene("empty", [], [scenes.Material("m")], env=scenes.make_envmap(256, 128, 2), camera=scenes.Camera(64, 40))
    rt = P.raytracer_for(sc); o = oracle_mod.Oracle(sc)
    p = dict(sampling_width=1, jitter=0, aperture=0.0)
    img, st = rt.render(p); ref = o.render(oracle_mod.make_params(**p))[0]
    assert st["primary"] == 64 * 40 and st["shadow"] == 0
    ok, psnr = image_bars(P, ref, img)
    assert ok >= 0.995


# ------------------------------------------------------------------------------------------------ leaf functions
def test_primary_rays_bit_exact(cornell_pair, oracle_mod):
    """Camera + jitter + lens sampling are plain IEEE float ops and integer hashing on both sides: bit-exact."""
    rt, o = cornell_pair
    for p in (dict(sampling_width=1, jitter=0, aperture=0.0), dict(sampling_width=3, seed=7), dict(sampling_width=2, camera_mode=1, seed=3)):
        a = rt.primary_rays(p); b = o.primary_rays(oracle_mod.make_params(**p))
        assert np.array_equal(a, b), p


def test_secondary_rays_bit_exact(cornell_pair):
    rt, o = cornell_pair
    rng = np.random.default_rng(5)
    n = 5000
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    nn = rng.normal(size=(n, 3)); nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    flip = np.sum(d * nn, axis=1) > 0; nn[flip] *= -1
    hp = rng.uniform(-100, 100, (n, 3))
    ior = np.where(rng.random(n) < 0.5, 1.000293, 1.5)
    items = np.concatenate([d, nn, hp, ior[:, None], (2.500293 - ior)[:, None]], axis=1).astype(np.float32)
    for refr in (False, True):
        a = rt.secondary_rays(items, refr); b = o.secondary_rays(items, refr)
        assert np.array_equal(a, b, equal_nan=True)
    assert np.isnan(rt.secondary_rays(items, True)[:, 4]).sum() > 100          # TIR cases are exercised and agree


def test_mix_srgb(cornell_pair):
    """double pow on both sides; CUDA's pow is within 2 ulp of glibc's in double => float results differ by <= 1 ulp."""
    rt, o = cornell_pair
    rng = np.random.default_rng(6)
    n = 20000
    c0 = rng.uniform(-0.2, 1.3, (n, 4)).astype(np.float32); c1 = rng.uniform(-0.2, 1.3, (n, 4)).astype(np.float32)
    al = rng.uniform(0, 1, n).astype(np.float32)
    a, b = rt.mix_srgb(c0, c1, al), o.mix_srgb(c0, c1, al)
    assert np.max(np.abs(a - b)) <= 1.2e-7 and np.mean(a == b) > 0.99


def test_gamma(cornell_pair):
    rt, o = cornell_pair
    rng = np.random.default_rng(7)
    c = rng.uniform(-0.1, 1.5, (20000, 4)).astype(np.float32)
    for g in (0.5, 0.25, 1.0):
        a, b = rt.gamma(c, g), o.gamma(c, g)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-6)) <= 3e-7      # <= ~2 ulp relative


def test_texture_get_texel_including_out_of_range(cornell_pair):
    """Bilinear fetch is plain float arithmetic: bit-exact, including the black-texel rules, the wrap column and
    clamped out-of-range coordinates (tiled model UVs)."""
    rt, o = cornell_pair
    rng = np.random.default_rng(8)
    uv = rng.uniform(-0.5, 1.5, (20000, 2)).astype(np.float32)
    grid = (np.arange(0, 65)[:, None] / np.float32(64)).astype(np.float32)
    uv = np.concatenate([uv, np.concatenate([grid, grid[::-1]], axis=1), rng.uniform(0, 1, (20000, 2)).astype(np.float32)])
    for tex in (0, 1, -1):
        a, b = rt.texture_get_texel(tex, uv), o.texture_get_texel(tex, uv)
        assert np.array_equal(a, b), tex


def test_env_get_texel(cornell_pair):
    """(u,v) come from atan2f / asinf: 1-ulp differences move the bilinear weights by ~1e-4 texel; tolerance 2e-3 in
    colour, with a 0.1 % budget for samples that land exactly on a texel boundary (black-texel rule)."""
    rt, o = cornell_pair
    rng = np.random.default_rng(9)
    d = rng.normal(size=(50000, 3)).astype(np.float32)
    a, b = rt.env_get_texel(d), o.env_get_texel(d)
    bad = np.abs(a - b).max(axis=1) > 2e-3
    assert bad.mean() <= 1e-3
    assert np.mean(np.all(a == b, axis=1)) > 0.25


# ------------------------------------------------------------------------------------------------ intersection
def test_intersect_bit_exact_vs_brute_force(cornell_pair, oracle_mod):
    rt, o = cornell_pair
    rng = np.random.default_rng(10)
    n = 20000
    org = rng.uniform(-150, 150, (n, 3)).astype(np.float32); org[:, 2] = np.abs(org[:, 2])
    target = rng.uniform(-70, 70, (n, 3)).astype(np.float32); target[:, 2] = rng.uniform(-5, 60, n)
    d = (target - org) * rng.uniform(0.01, 3.0, (n, 1)).astype(np.float32)
    d[:100, 0] = 0.0; d[100:200, 1] = 0.0; d[200:300, 2] = 0.0                      # axis-parallel components
    rh = oracle_mod.make_rayhits(org, d, tnear=0.01)
    a = rt.intersect(rh); b = o.intersect(rh, brute=True)
    assert (b["geomID"] != INVALID).mean() > 0.3
    for f in ("tfar", "u", "v", "geomID", "primID", "Ng_x", "Ng_y", "Ng_z"):
        assert np.array_equal(a[f], b[f]), f


def test_intersect_soup_and_duplicates(P, oracle_mod):
    """Random soup + exactly coincident duplicate triangles: ties resolve to the lowest primID on both sides."""
    sc = scenes.triangle_soup(5000, seed=4, resolution=(96, 54))
    m = sc.meshes[0]
    sc.meshes.append(scenes.Mesh("dup", m.pos[:500].copy(), m.nrm[:500].copy(), m.uv[:500].copy(), 0))
    rt = P.raytracer_for(sc); o = oracle_mod.Oracle(sc)
    rays = o.primary_rays(oracle_mod.make_params(sampling_width=2, jitter=1, aperture=0.0))
    rh = oracle_mod.make_rayhits(rays[:, :3], rays[:, 4:7], tnear=0.01)
    a = rt.intersect(rh); b = o.intersect(rh, brute=True)
    for f in ("tfar", "u", "v", "geomID", "primID"):
        assert np.array_equal(a[f], b[f]), f
    assert not np.any((a["geomID"] == 1) & (a["primID"] < 500))                    # the duplicate never wins the tie


def test_interpolate_bit_exact(avenger):
    sc, rt, o = avenger
    rng = np.random.default_rng(11)
    n = 5000
    geom = rng.integers(0, len(sc.meshes), n).astype(np.uint32)
    prim = np.array([rng.integers(0, sc.meshes[g].ntris) for g in geom], np.uint32)
    u = rng.random(n).astype(np.float32) * 0.5; v = rng.random(n).astype(np.float32) * 0.5
    for slot in (0, 1):
        assert np.array_equal(rt.interpolate(geom, prim, u, v, slot), o.interpolate(geom, prim, u, v, slot))


# ------------------------------------------------------------------------------------------------ frames
@pytest.mark.parametrize("tag", ["c1", "dof3x3", "lambert", "pinhole_depth3"])
def test_frames_vs_committed_golden(P, cornell, tag):
    """CUDA frames against the committed oracle renders (no oracle needed at run time)."""
    z = np.load(os.path.join(GOLDEN, f"cornell_{tag}.npz"))
    p = json.loads(str(z["params"]))
    rt = P.raytracer_for(cornell)
    img, st = rt.render(p)
    g, pr = rt.primary_ids(p)
    assert np.mean((g == z["geom"]) & (pr == z["prim"])) >= 0.999
    ok, psnr = image_bars(P, z["rgba"], img)
    assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
    assert [st["primary"], st["shadow"], st["reflection"], st["refraction"]] == list(z["rays"])


def test_avenger_C1_parity_gates(P, oracle_mod, avenger):
    """Config C1: 640x480, Whitted, depth 7, 1 spp un-jittered, aperture 0."""
    sc, rt, o = avenger
    p = dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=7)
    ref, g0, p0, st0 = o.render(oracle_mod.make_params(**p))
    img, st = rt.render(p)
    g1, p1 = rt.primary_ids(p)
    assert np.mean((g0 == g1) & (p0 == p1)) >= 0.999
    ok, psnr = image_bars(P, ref, img)
    assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
    for k in ("primary", "shadow", "reflection", "refraction"):
        assert abs(st[k] - st0[k]) <= 1e-3 * max(st0[k], 1), (k, st[k], st0[k])
    assert st["reflection"] > 1000 and st["refraction"] > 1000 and st["shadow"] > 50000


def test_avenger_default_render_3x3_dof(P, oracle_mod, avenger):
    """The reference's shipped settings (3x3 stratified, thin lens f=200 a=5) with the shared counter RNG, on a crop-sized frame."""
    sc, rt, o = avenger
    rt.set_camera(320, 240, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
    o.set_camera(320, 240, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
    try:
        p = dict(seed=21)
        ref, g0, p0, st0 = o.render(oracle_mod.make_params(**p))
        img, st = rt.render(p)
        ok, psnr = image_bars(P, ref, img)
        assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
        assert st["primary"] == 320 * 240 * 9
        for k in ("shadow", "reflection", "refraction"):
            assert abs(st[k] - st0[k]) <= 1e-3 * max(st0[k], 1), (k, st[k], st0[k])
    finally:
        rt.set_camera(640, 480, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
        o.set_camera(640, 480, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)


def test_multiple_lights_sum_in_order(P, oracle_mod, cornell):
    """The commented coloured lights of pg1/raytracer.cpp:54-63 plus the white one."""
    sc = scenes.cornell_like()
    sc.camera = cornell.camera
    sc.lights = [scenes.Light((-10, 0, 200), (0.8, 0.2, 0.2), (0.8, 0.2, 0.2), (0.8, 0.2, 0.2)),
                 scenes.Light((0, 80, 150), (0.2, 0.8, 0.2), (0.2, 0.8, 0.2), (0.2, 0.8, 0.2)), scenes.Light()]
    rt = P.raytracer_for(sc); o = oracle_mod.Oracle(sc)
    p = dict(sampling_width=1, jitter=0, aperture=0.0)
    ref, _, _, st0 = o.render(oracle_mod.make_params(**p)); img, st = rt.render(p)
    ok, psnr = image_bars(P, ref, img)
    assert ok >= 0.995 and psnr >= 45.0
    assert st["shadow"] == st0["shadow"] and st["shadow"] > 2 * 64 * 48 * 0.3


def test_get_pixel_hook_serves_the_rendered_frame(cornell_pair):
    """The reference's plug-in hook Color4f get_pixel(x, y, t) (pg1/simpleguidx11.h:27)."""
    rt, o = cornell_pair
    p = dict(sampling_width=1, jitter=0, aperture=0.0)
    img, _ = rt.render(p)
    for (x, y) in ((0, 0), (63, 47), (31, 20)):
        px = rt.get_pixel(x, y, 0.0, p)
        assert np.array_equal(np.array(px, np.float32), img[y, x], equal_nan=True)


# ------------------------------------------------------------------------------------------------ size-independent properties
def test_render_is_deterministic_and_batching_invariant(P, cornell, monkeypatch):
    rt = P.raytracer_for(cornell)
    p = dict(seed=2)
    a, sa = rt.render(p); b, sb = rt.render(p)
    assert np.array_equal(a, b, equal_nan=True)
    monkeypatch.setenv("PGRT_MAX_BATCH_SAMPLES", "4096")            # many small batches
    rt2 = P.raytracer_for(cornell)
    c, sc_ = rt2.render(p)
    assert sc_["batches"] > 1 and np.array_equal(a, c, equal_nan=True)
    assert (sa["shadow"], sa["reflection"], sa["refraction"]) == (sc_["shadow"], sc_["reflection"], sc_["refraction"])


def test_build_and_queue_variants_render_the_same_frame(P, cornell, monkeypatch):
    """Knobs that change HOW the frame is computed must not change it: stored vs regenerated level-0 rays,
    quantised vs float node planes, Karras LBVH vs PLOC tree (the hit record does not depend on the tree)."""
    p = dict(seed=6)
    ref, st0 = P.raytracer_for(cornell).render(p)
    for env in (dict(PGRT_FUSE_RAYGEN="0"), dict(PGRT_NODE_LAYOUT="q8"), dict(PGRT_NODE_LAYOUT="f32"), dict(PGRT_BUILDER="lbvh"),
                dict(PGRT_BUILDER="lbvh", PGRT_NODE_LAYOUT="q8", PGRT_FUSE_RAYGEN="0")):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        rt = P.raytracer_for(cornell)
        img, st = rt.render(p)
        for k in env:
            monkeypatch.delenv(k)
        assert np.array_equal(img, ref, equal_nan=True), env
        assert [st[x] for x in ("primary", "shadow", "reflection", "refraction")] == [st0[x] for x in ("primary", "shadow", "reflection", "refraction")]
        if "PGRT_NODE_LAYOUT" in env:
            assert rt.build_stats["node_bytes"] == (80 if env["PGRT_NODE_LAYOUT"] == "q8" else 240)


def test_pipelined_frames_equal_synchronous_frames(P, avenger):
    """pgrt_render_begin / pgrt_render_end: four frames in flight in four slots (own streams, queues, counters) give the
    frames and ray counts of four synchronous renders; slots are reusable; a busy slot refuses a second frame."""
    import torch
    sc, rt, o = avenger
    plist = [dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=10), dict(sampling_width=2, seed=4), dict(seed=9, max_depth=3), dict(seed=11, scheduler=1)]
    ref = [rt.render(p) for p in plist]
    bufs = [torch.empty((rt.height, rt.width, 4), dtype=torch.float32).pin_memory() for _ in plist]
    for rnd in range(2):
        for k, p in enumerate(plist):
            rt.render_begin(k, p, host_ptr=bufs[k].data_ptr())
        with pytest.raises(P.PgrtError):
            rt.render_begin(1, plist[1], host_ptr=bufs[1].data_ptr())
        for k in reversed(range(len(plist))):
            st = rt.render_end(k)
            assert np.array_equal(bufs[k].numpy(), ref[k][0], equal_nan=True), (rnd, k)
            assert [st[x] for x in ("primary", "shadow", "reflection", "refraction")] == [ref[k][1][x] for x in ("primary", "shadow", "reflection", "refraction")]
            bufs[k].zero_()
    with pytest.raises(P.PgrtError):
        rt.render_end(0)                       # nothing in flight
    dev = torch.zeros((rt.height, rt.width, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rt.render_begin(2, plist[0], device_ptr=dev.data_ptr()); rt.render_end(2)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), ref[0][0], equal_nan=True)


def test_schedulers_agree_bit_for_bit(P, cornell, avenger):
    """The fused persistent kernel (shared ray pool, continuation-passing combine, no level barrier) and the level-synchronous
    wavefront evaluate the same tree with the same arithmetic: frames and ray counts must be identical."""
    sc, rt, o = avenger
    cases = [(P.raytracer_for(cornell), dict(seed=5)), (P.raytracer_for(cornell), dict(seed=5, max_depth=1)),
             (rt, dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=10)), (rt, dict(sampling_width=2, seed=9, max_depth=4))]
    for r, p in cases:
        a, sa = r.render(dict(p, scheduler=0)); b, sb = r.render(dict(p, scheduler=1))
        h, sh = r.render(dict(p, scheduler=2))       # hybrid: level 0 as wavefront kernels, levels >= 1 through the pool
        assert np.array_equal(a, b, equal_nan=True), p
        assert np.array_equal(a, h, equal_nan=True), p
        for k in ("primary", "shadow", "reflection", "refraction"):
            assert sa[k] == sb[k] == sh[k], (k, p)
        assert sa["launches"] < sb["launches"] or p.get("max_depth") == 1
        assert sa["launches"] < sh["launches"] <= 7
    la = rt.render(dict(cases[2][1], scheduler=0), profile=3)[1]; lv0 = rt.level_stats()
    lb = rt.render(dict(cases[2][1], scheduler=1), profile=3)[1]; lv1 = rt.level_stats()
    lh = rt.render(dict(cases[2][1], scheduler=2), profile=3)[1]; lv2 = rt.level_stats()
    assert [x["rays"] for x in lv0] == [x["rays"] for x in lv1] and [x["shadow_rays"] for x in lv0] == [x["shadow_rays"] for x in lv1]
    assert [x["rays"] for x in lv0] == [x["rays"] for x in lv2] and [x["shadow_rays"] for x in lv0] == [x["shadow_rays"] for x in lv2]
    assert (la["nodes_visited"], la["tris_tested"]) == (lb["nodes_visited"], lb["tris_tested"]) and la["nodes_visited"] > 0
    assert (la["nodes_visited"], la["tris_tested"]) == (lh["nodes_visited"], lh["tris_tested"])


@pytest.mark.parametrize("scheduler", [0, 1, 2])
def test_queue_overflow_is_retried_with_larger_queues(P, cornell, monkeypatch, scheduler):
    # the fused scheduler keeps ONE pool (4 x cap) for all levels; this frame needs far more than 4 x 2500 records
    monkeypatch.setenv("PGRT_MIN_LEVEL_CAP", "2500" if scheduler != 1 else "2048"); monkeypatch.setenv("PGRT_LEVEL_CAP_FACTOR", "0.05")
    rt = P.raytracer_for(cornell)
    p = dict(seed=2, scheduler=scheduler)
    img, st = rt.render(p)
    assert st["overflow_retries"] >= 1
    img2, st2 = rt.render(p)                                   # the capacity that survived is remembered: no second overflow
    assert st2["overflow_retries"] == 0 and np.array_equal(img, img2, equal_nan=True)
    monkeypatch.delenv("PGRT_MIN_LEVEL_CAP"); monkeypatch.delenv("PGRT_LEVEL_CAP_FACTOR")
    ref, _ = P.raytracer_for(cornell).render(p)
    assert np.array_equal(img, ref, equal_nan=True)


def test_queue_overflow_with_direct_pixel_stores(P, cornell, monkeypatch):
    """One sample per pixel: the hybrid's shading kernels and the fused kernel store finished pixels straight into the frame;
    an attempt whose pool overflowed has stored black for the nodes that did not fit, and the retry must overwrite all of it."""
    p = dict(seed=2, sampling_width=1, jitter=0, aperture=0.0)
    ref, st_ref = P.raytracer_for(cornell).render(dict(p, scheduler=1))
    monkeypatch.setenv("PGRT_MIN_LEVEL_CAP", "100"); monkeypatch.setenv("PGRT_LEVEL_CAP_FACTOR", "0.01")
    for sched in (0, 2):
        rt = P.raytracer_for(cornell)
        img, st = rt.render(dict(p, scheduler=sched))
        assert st["overflow_retries"] >= 1, sched
        assert np.array_equal(img, ref, equal_nan=True) and st["total"] == st_ref["total"], sched


def test_overflowed_attempt_does_not_release_stream_consumers(P, cornell, monkeypatch):
    """a consumer ordered behind a slot with pgrt_stream_wait_slot must not see the frame of an attempt whose
    queues overflowed (black dielectric nodes); it is released by the retry that pgrt_render_end runs."""
    import torch
    monkeypatch.setenv("PGRT_MIN_LEVEL_CAP", "2500"); monkeypatch.setenv("PGRT_LEVEL_CAP_FACTOR", "0.05")
    rt = P.raytracer_for(cornell)
    p = dict(seed=2)
    dev = torch.zeros((rt.height, rt.width, 4), dtype=torch.float32, device="cuda")
    copy = torch.zeros_like(dev)
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    rt.render_begin(1, p, device_ptr=dev.data_ptr())
    rt.stream_wait_slot(1, side.cuda_stream)
    with torch.cuda.stream(side):
        copy.copy_(dev, non_blocking=True)
    st = rt.render_end(1)
    assert st["overflow_retries"] >= 1
    torch.cuda.synchronize()
    monkeypatch.delenv("PGRT_MIN_LEVEL_CAP"); monkeypatch.delenv("PGRT_LEVEL_CAP_FACTOR")
    ref, _ = P.raytracer_for(cornell).render(p)
    assert np.array_equal(copy.cpu().numpy(), ref, equal_nan=True)


def test_rgba8_frame_is_the_quantised_float_frame(P, cornell, avenger):
    """pgrt_render_rgba8 = round(clamp(c,0,1)*255), NaN -> 0 of the float frame (what D3D11 presents, simpleguidx11.cpp:229,290)."""
    sc, rt, o = avenger
    for r, p in ((P.raytracer_for(cornell), dict(seed=4)), (rt, dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=10)),
                 (rt, dict(sampling_width=2, seed=3, scheduler=1))):
        f, sf = r.render(p)
        q, sq = r.render_rgba8(p)
        assert np.array_equal(q[..., :3], P.to_srgb8(f)) and np.all(q[..., 3] == 255)
        assert sf["total"] == sq["total"]


def test_sharded_render_is_bit_identical_to_single_gpu(P, avenger):
    """Tile sharding must not change any pixel: emulate 1/2/3/8 ranks on one GPU, un-tile with the CUDA kernel."""
    import torch
    sc, rt, o = avenger
    p = dict(sampling_width=2, seed=3)
    full, st_full = rt.render(p)
    try:
        for n in (2, 3, 8):
            spr = None; bufs = []; rays = 0
            for r in range(n):
                rt.set_shard(r, n)
                spr = rt.shard_pixels()
                t = torch.zeros((spr, 4), dtype=torch.float32, device="cuda")
                torch.cuda.synchronize()                     # the library renders on its own stream
                st = rt.render_shard_device(t.data_ptr(), p)
                torch.cuda.synchronize()
                bufs.append(t); rays += st["total"]
            g = torch.cat(bufs)
            out = torch.zeros((rt.height, rt.width, 4), dtype=torch.float32, device="cuda")
            torch.cuda.synchronize()
            rt.untile(g.data_ptr(), n, out.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(out.cpu().numpy(), full, equal_nan=True), n
            assert rays == st_full["total"]
            from pgi_raytracing_b200 import dist as D
            assert np.array_equal(D.untile_numpy(g.cpu().numpy(), rt.width, rt.height, n), full, equal_nan=True)
    finally:
        rt.set_shard(0, 1)


def test_full_size_1080p_properties(P, avenger):
    """BASELINE config C2 at full size, through properties that need no CPU render: ray accounting identities,
    determinism, depth monotonicity, and agreement of pgrt_intersect with the frame's own primary ids."""
    sc, rt, o = avenger
    rt.set_camera(1920, 1080, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
    try:
        p = dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=10)
        a, sa = rt.render(p); b, sb = rt.render(p)
        assert np.array_equal(a, b, equal_nan=True) and sa["total"] == sb["total"]
        assert sa["primary"] == 1920 * 1080 and sa["refraction"] <= sa["reflection"]
        assert sa["launches"] == 6                    # the automatic choice at this size: the hybrid scheduler (pgrt.h); one sample per pixel: no k_resolve
        for sched, launches in ((0, 3), (2, 6)):      # ... and the same frame from the fused and the explicit hybrid scheduler
            c, sc_ = rt.render(dict(p, scheduler=sched))
            assert np.array_equal(a, c, equal_nan=True) and sc_["launches"] == launches
            assert all(sa[k] == sc_[k] for k in ("primary", "shadow", "reflection", "refraction"))
        s7 = rt.render(dict(p, max_depth=7))[1]
        assert s7["reflection"] <= sa["reflection"] and s7["primary"] == sa["primary"]
        assert np.all(a[..., 3] == 1.0)
        g, pr = rt.primary_ids(p)
        rays = rt.primary_rays(p)
        sel = np.random.default_rng(0).integers(0, 1920 * 1080, 50000)
        from oracle.oracle import make_rayhits
        rh = make_rayhits(rays[sel, :3], rays[sel, 4:7], tnear=0.01)
        hit = rt.intersect(rh)
        assert np.array_equal(hit["geomID"], g.reshape(-1)[sel]) and np.array_equal(hit["primID"], pr.reshape(-1)[sel])
    finally:
        rt.set_camera(640, 480, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)


def test_errors_are_reported_not_thrown(P):
    rt = P.Raytracer(64, 48, 0.7, (0, 0, 10), (0, 0, 0))
    with pytest.raises(P.PgrtError) as e:
        rt.render(dict())                       # nothing committed
    assert e.value.code == 1 and "commit" in str(e.value)
    rt.LoadScene(scenes.single_triangle())
    with pytest.raises(P.PgrtError):
        rt.render(dict(sampling_width=0))
    with pytest.raises(P.PgrtError):
        rt.render(dict(max_depth=64))
    # maximum size: flat triangle ids are 29 bits; the call is refused before a single byte is read
    import ctypes as C
    dummy = np.zeros(18, np.float32)
    rc = rt.lib.pgrt_add_mesh(rt.h, dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, 1 << 29, 0, None)
    assert rc == 1 and b"2^29" in rt.lib.pgrt_last_error(rt.h)
    # a mesh that names a material nobody set is reported at render time, not dereferenced
    g = C.c_uint32()
    assert rt.lib.pgrt_add_mesh(rt.h, dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, 2, 7, C.byref(g)) == 0
    rt.commit()
    with pytest.raises(P.PgrtError) as e2:
        rt.render(dict())
    assert "material" in str(e2.value)


def test_two_ranks_p2p_and_nccl_gather_are_bit_identical():
    """One process per GPU (tools/p2p_check.py): the peer-mapped direct-write path (frame kernel stores into rank 0's frame over
    NVLink) with completion by flags in shared host memory and by the NCCL all-reduce fallback, the NCCL gather path, the
    shared-host-frame path in float and 8-bit, and a run whose ray pools overflow (retried frames must not be seen early) --
    all against a single-GPU render, bit for bit, including what a consumer ordered right behind begin() copies.
    Needs two GPUs; the single-GPU box of the round-end run skips it."""
    import subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", os.path.join(root, "tools", "p2p_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "p2p: 2 ranks" in out.stdout and "mode used = p2p" in out.stdout and "nccl: 2 ranks" in out.stdout and "host: 2 ranks" in out.stdout
    lines = out.stdout.splitlines()
    for proto in ("flags+counter", "flags", "nccl-allreduce"):                                      # all three completion protocols ran
        assert any(l.endswith("completion = " + proto) for l in lines), proto
    assert "host rgba8: 2 ranks" in out.stdout and "p2p squeezed-pool: 2 ranks" in out.stdout and "host squeezed-pool: 2 ranks" in out.stdout


def test_tiles_stored_straight_into_registered_host_memory(P, cornell):
    """pgrt_host_frame_register: the resolve kernel writes through the mapping of a page-aligned host buffer (what the
    ranks of a box share in mode "host" of dist.ShardedRenderer); the frame must equal the ordinary host render."""
    import ctypes, mmap
    rt = P.raytracer_for(cornell)
    params = dict(sampling_width=2, seed=4)
    ref, st = rt.render(params)
    nbytes = (rt.width * rt.height * 16 + 4095) // 4096 * 4096
    mm = mmap.mmap(-1, nbytes)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    dev_ptr = rt.host_frame_register(base, nbytes)
    try:
        rt.set_shard(0, 1)
        for slot in (0, 1):                      # two frames in flight into the same host frame (same pixels)
            rt.render_begin(slot, params, frame_ptr=dev_ptr)
        tot = [rt.render_end(slot)["total"] for slot in (0, 1)]
        got = np.frombuffer(mm, dtype=np.float32, count=rt.width * rt.height * 4).reshape(rt.height, rt.width, 4).copy()
    finally:
        rt.host_frame_unregister(base)
    assert tot == [st["total"], st["total"]]
    assert np.array_equal(got, ref, equal_nan=True)
    with pytest.raises(P.PgrtError):
        rt.host_frame_register(base + 8, 4096)   # not page-aligned


def test_dof_converges_to_the_same_mean_for_different_seeds(P, oracle_mod, cornell):
    """north_star: 'DOF is compared as a converged mean'.  The reference seeds its samplers from the clock; here two
    different seeds stand for two runs: GPU (seed 5) and oracle (seed 77), 12x12 samples per pixel, thin lens f=200 a=5.
    Same seeds are bit-level parity (golden tests above); different seeds must agree as means."""
    rt = P.raytracer_for(cornell); o = oracle_mod.Oracle(cornell)
    img, _ = rt.render(dict(sampling_width=12, seed=5))
    ref = o.render(oracle_mod.make_params(sampling_width=12, seed=77))[0]
    a, b = P.to_srgb8(ref).astype(float), P.to_srgb8(img).astype(float)
    psnr = 10 * np.log10(255.0 ** 2 / np.mean((a - b) ** 2))
    assert psnr >= 30.0, psnr                      # 144 spp Monte-Carlo noise of two independent runs
    assert abs(a.mean() - b.mean()) <= 0.5         # no bias: means of the whole frame within half an LSB
    img2, _ = rt.render(dict(sampling_width=12, seed=77))
    ok, psnr_same = image_bars(P, ref, img2)
    assert ok >= 0.995 and psnr_same >= 45.0       # and with the SAME seed the usual bars hold


def test_large_soup_quantised_nodes_match_the_oracle_bvh(P, oracle_mod):
    """300 k random triangles: above the tests' usual sizes, through the 80-B quantised layout (forced) and the float
    one; closest hits bit-exact against the oracle's own BVH traversal (brute force would take minutes)."""
    import os as _os
    sc = scenes.triangle_soup(300_000, seed=9, resolution=(160, 90))
    o = oracle_mod.Oracle(sc)
    rays = o.primary_rays(oracle_mod.make_params(sampling_width=2, jitter=1, aperture=0.0, seed=4))
    rh = oracle_mod.make_rayhits(rays[:, :3], rays[:, 4:7], tnear=0.01)
    b = o.intersect(rh, brute=False)
    assert (b["geomID"] != INVALID).mean() > 0.2
    for layout in ("q8", "f32"):
        _os.environ["PGRT_NODE_LAYOUT"] = layout
        try:
            rt = P.raytracer_for(sc)
        finally:
            del _os.environ["PGRT_NODE_LAYOUT"]
        assert rt.build_stats["node_bytes"] == (80 if layout == "q8" else 240)
        a = rt.intersect(rh)
        for f in ("tfar", "u", "v", "geomID", "primID"):
            assert np.array_equal(a[f], b[f]), (layout, f)


def test_graph_replay_of_frames(P, cornell, monkeypatch):
    """A frame whose inputs did not change is re-submitted as one CUDA graph launch (first frame direct, second captured,
    later ones replayed): every variant must give the same frame, and any change of params / camera / scene must be seen."""
    rt = P.raytracer_for(cornell)
    p = dict(seed=8)
    frames = [rt.render(p) for _ in range(5)]
    for img, st in frames[1:]:
        assert np.array_equal(img, frames[0][0], equal_nan=True) and st["total"] == frames[0][1]["total"]
        assert st["launches"] == frames[0][1]["launches"]
    q = dict(seed=9)
    other = [rt.render(q) for _ in range(3)]
    assert not np.array_equal(other[0][0], frames[0][0], equal_nan=True)
    assert all(np.array_equal(o[0], other[0][0], equal_nan=True) for o in other)
    c = cornell.camera
    rt.set_camera(c.width, c.height, c.fov_y, (c.view_from[0] + 5.0, c.view_from[1], c.view_from[2]), c.view_at)
    moved = [rt.render(p)[0] for _ in range(3)]
    assert not np.array_equal(moved[0], frames[0][0], equal_nan=True) and np.array_equal(moved[0], moved[2], equal_nan=True)
    rt.set_camera(c.width, c.height, c.fov_y, c.view_from, c.view_at)
    assert np.array_equal(rt.render(p)[0], frames[0][0], equal_nan=True)
    monkeypatch.setenv("PGRT_GRAPHS", "0")
    plain = P.raytracer_for(cornell)
    assert np.array_equal(plain.render(p)[0], frames[0][0], equal_nan=True)


def test_hard_shadows_flag_is_opt_in(P, cornell):
    """shadow_mode = 1 (README to-do 'hard shadows', no reference counterpart): shadows appear -- fewer lit Phong hits,
    some pixels darker, none brighter than 1 LSB -- and the default stays the shipped behaviour, bit for bit."""
    rt = P.raytracer_for(cornell)
    p = dict(sampling_width=1, jitter=0, aperture=0.0)
    a, sa = rt.render(p); b, sb = rt.render(dict(p, shadow_mode=1)); c, _ = rt.render(dict(p, shadow_mode=0))
    assert np.array_equal(a, c, equal_nan=True)
    qa, qb = P.to_srgb8(a).astype(int), P.to_srgb8(b).astype(int)
    # hard shadows only remove light from Phong hits seen directly; through glass the non-linear mix may move either way
    assert (qb.sum(-1) < qa.sum(-1) - 6).mean() > 0.01
    assert sb["primary"] == sa["primary"] and sb["shadow"] > 0
    for bad in (2, -1):
        with pytest.raises(P.PgrtError):
            rt.render(dict(p, shadow_mode=bad))


def test_awkward_geometry(P, oracle_mod):
    """Inputs that stress the builder rather than the shader: thousands of coincident triangles, a 'teapot in a stadium'
    (two huge triangles + a dense cluster), triangles on a line, zero-area triangles.  The tree must build within the
    traversal stack depth and closest hits must still equal brute force bit for bit."""
    rng = np.random.default_rng(12)
    def tri_cloud(n, centre, spread, size):
        c = centre + rng.normal(size=(n, 1, 3)) * spread
        return (c + rng.normal(size=(n, 3, 3)) * size).astype(np.float32)
    one = tri_cloud(1, np.zeros(3), 0.0, 5.0)
    cases = {
        "coincident": np.repeat(one, 3000, axis=0),
        "stadium": np.concatenate([np.array([[[-500, -500, 0], [500, -500, 0], [500, 500, 0]], [[-500, -500, 0], [500, 500, 0], [-500, 500, 0]]], np.float32),
                                   tri_cloud(6000, np.array([400.0, 400.0, 5.0]), 2.0, 0.05)]),
        "line": (np.arange(4000, dtype=np.float32)[:, None, None] * np.array([0.01, 0.0, 0.0], np.float32) + tri_cloud(4000, np.zeros(3), 0.0, 0.3)).astype(np.float32),
        "degenerate": np.concatenate([tri_cloud(500, np.zeros(3), 20.0, 2.0), np.repeat(tri_cloud(200, np.zeros(3), 20.0, 0.0)[:, :1], 3, axis=1)]),
    }
    for name, pos in cases.items():
        n = pos.shape[0]
        nrm = np.tile(np.array([0, 0, 1], np.float32), (n, 3, 1)); uv = np.zeros((n, 3, 2), np.float32)
        sc = scenes.Scene(name, [scenes.Mesh("m", pos, nrm, uv, 0)], [scenes.Material("m")], camera=scenes.Camera(32, 24))
        rt = P.raytracer_for(sc); o = oracle_mod.Oracle(sc)
        assert rt.build_stats["depth"] <= 38, (name, rt.build_stats)
        ctr = pos.reshape(-1, 3).mean(0)
        org = (ctr + rng.normal(size=(3000, 3)) * 300).astype(np.float32)
        tgt = pos[rng.integers(0, n, 3000)].mean(1) + rng.normal(size=(3000, 3)).astype(np.float32) * 0.02
        rh = oracle_mod.make_rayhits(org, (tgt - org).astype(np.float32), tnear=1e-3)
        a = rt.intersect(rh); b = o.intersect(rh, brute=True)
        for f in ("tfar", "u", "v", "geomID", "primID"):
            assert np.array_equal(a[f], b[f]), (name, f)
        assert (b["geomID"] != INVALID).mean() > 0.3, name


def test_non_finite_vertices_are_an_error_not_a_hang(P):
    pos = np.random.default_rng(3).normal(size=(600, 3, 3)).astype(np.float32)
    pos[17, 1, 2] = np.nan; pos[300] = np.inf
    nrm = np.tile(np.array([0, 0, 1], np.float32), (600, 3, 1)); uv = np.zeros((600, 3, 2), np.float32)
    sc = scenes.Scene("nan", [scenes.Mesh("m", pos, nrm, uv, 0)], [scenes.Material("m")], camera=scenes.Camera(32, 24))
    try:
        rt = P.raytracer_for(sc)
    except P.PgrtError as e:
        assert "cluster" in str(e) or "deeper" in str(e)
        return
    img, st = rt.render(dict(sampling_width=1, jitter=0, aperture=0.0))   # if the build went through, frames must still come back
    assert st["primary"] == 32 * 24


def test_cross_frame_accumulation(P, oracle_mod, cornell):
    """pgrt_render_accumulate: n finished frames, seeds s, s+1, ..., summed on the device in frame order and
    divided by n.  (1) bit-exact against the same float sum of the library's own single frames; (2) against the oracle's
    frames of the same seeds within the image bars; (3) the ray count is the sum; (4) more frames = closer to a reference
    of many samples."""
    rt = P.raytracer_for(cornell); o = oracle_mod.Oracle(cornell)
    base = dict(sampling_width=2, seed=11)
    n = 6                                            # more frames than the 4 slots it keeps in flight
    singles, rays = [], 0
    for i in range(n):
        img, st = rt.render(dict(base, seed=11 + i)); singles.append(img.copy()); rays += st["total"]
    acc = singles[0].copy()
    for f in singles[1:]:
        acc = acc + f
    expect = acc / np.float32(n)
    got, st = rt.render_accumulate(n, base)
    assert np.array_equal(got, expect, equal_nan=True)
    assert st["total"] == rays
    ref = None
    for i in range(n):
        f = o.render(oracle_mod.make_params(sampling_width=2, seed=11 + i))[0]
        ref = f.copy() if ref is None else ref + f
    ok, psnr = image_bars(P, ref / np.float32(n), got)
    assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
    # convergence: against 12x12 samples of another seed the 6-frame mean beats a single frame
    many, _ = rt.render(dict(sampling_width=12, seed=1234))
    err = lambda a: float(np.mean((P.to_srgb8(a).astype(float) - P.to_srgb8(many).astype(float)) ** 2))
    assert err(got) < 0.6 * err(singles[0])
    one, st1 = rt.render_accumulate(1, base)
    assert np.array_equal(one, singles[0], equal_nan=True)
    with pytest.raises(P.PgrtError):
        rt.render_accumulate(0, base)


@pytest.mark.parametrize("scheduler", [0, 1, 2])
def test_path_tracing_mode_matches_the_oracle(P, oracle_mod, cornell, scheduler):
    """shader_mode = 3 (README.md:21 'To do: path tracing'; no reference counterpart, so the spec is include/pgrt.h and the
    oracle restates it): every non-dielectric hit = its Phong value + albedo x one cosine-weighted bounce.  The bounce
    direction uses +,*,/,sqrt only and is keyed by the bits of the hit point, so GPU and oracle agree ray for ray."""
    rt = P.raytracer_for(cornell); o = oracle_mod.Oracle(cornell)
    kw = dict(sampling_width=2, seed=9, shader_mode=3, max_depth=5)
    img, st = rt.render(dict(kw, scheduler=scheduler))
    ref, _, _, rst = o.render(oracle_mod.make_params(**kw), want_ids=False)
    ok, psnr = image_bars(P, ref, img)
    assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
    assert st["total"] == rst["total"] and st["reflection"] == rst["reflection"] and st["shadow"] == rst["shadow"]
    assert st["overflow_retries"] == 0            # queues are sized for one bounce per hit and level
    whitted, wst = rt.render(dict(kw, shader_mode=0, scheduler=scheduler))
    assert st["reflection"] > 2 * wst["reflection"]                       # every diffuse hit bounces
    assert P.to_srgb8(img).astype(int).sum() > P.to_srgb8(whitted).astype(int).sum()   # the environment lights the scene


def test_path_tracing_converges_and_covers_textures_and_glass(P, oracle_mod, avenger):
    """The avenger stand-in (94 surfaces, textures, a dielectric) in path mode: parity with the oracle, and the
    accumulated mean of 8 frames is closer to a 16-frame mean of other seeds than one frame is."""
    sc, rt, o = avenger
    kw = dict(sampling_width=1, jitter=1, aperture=0.0, seed=21, shader_mode=3, max_depth=4)
    img, st = rt.render(kw)
    ref, _, _, rst = o.render(oracle_mod.make_params(**kw), want_ids=False)
    ok, psnr = image_bars(P, ref, img)
    assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
    assert st["total"] == rst["total"]
    one = img.copy()
    acc8, _ = rt.render_accumulate(8, kw)
    many, _ = rt.render_accumulate(16, dict(kw, seed=1000))
    err = lambda a: float(np.mean((np.clip(np.nan_to_num(a[..., :3]), 0, 1) - np.clip(np.nan_to_num(many[..., :3]), 0, 1)) ** 2))
    assert err(acc8) < 0.5 * err(one), (err(acc8), err(one))


# ------------------------------------------------------------------------------------------------ the class's query methods
@pytest.mark.parametrize("which", ["cornell", "avenger"])
def test_trace_on_caller_rays_matches_the_oracle(P, oracle_mod, cornell_pair, avenger, which):
    """pgrt_trace = Raytracer::trace(RTCRay, level) (pg1/raytracer.h:31, raytracer.cpp:237-394) on caller-supplied rays at
    levels 0..7: camera rays, rays from inside the scene, rays that start inside glass.  Tolerance: the colours agree to
    5e-4 absolute = 1/8 of an 8-bit step and are bit-equal on >= 95 % of the rays: atan2 / asin / pow / exp of CUDA and glibc
    differ in the last place, one ulp of the environment map's u is 2.4e-4 texels of a 4000-pixel map, and mix_srgb chains
    double pows through the steep start of the sRGB curve.  The device recursion and the frame kernel give the same colour bit for bit."""
    rt, o = cornell_pair if which == "cornell" else avenger[1:]
    pd = dict(sampling_width=1, jitter=0, aperture=0.0)
    p = oracle_mod.make_params(**pd)
    rng = np.random.default_rng(31)
    prim = o.primary_rays(p)
    n = 20000
    rays = np.zeros((n, 9), np.float32)
    rays[:, 0:3] = rng.uniform(-60, 60, (n, 3)); rays[:, 2] = rng.uniform(2, 70, n)
    rays[:, 4:7] = rng.normal(size=(n, 3)); rays[:, 3] = 0.01; rays[:, 8] = np.finfo(np.float32).max
    rays[:, 7] = np.where(rng.random(n) < 0.7, np.float32(1.000293), 1.5)
    allrays = np.concatenate([prim[rng.integers(0, prim.shape[0], n)], rays])
    for level in (0, 1, 3, 6, 7):
        a, b = rt.trace(allrays, level, pd), o.trace(p, allrays, level)
        eq = ((a == b) | (np.isnan(a) & np.isnan(b))).all(axis=-1)
        assert eq.mean() >= 0.95 and np.nanmax(np.abs(a - b)) <= 5e-4, (which, level, eq.mean(), np.nanmax(np.abs(a - b)))
    # level 0 on the frame's own primary rays = the frame before the resolve (one sample per pixel, gamma 0.5 = identity up to rounding)
    img, _ = rt.render(pd)
    t = rt.trace(prim, 0, pd).reshape(rt.height, rt.width, 4)
    lin = np.sqrt(t[..., [2, 1, 0]]) ** 2        # the resolve swaps into gamma and back (raytracer.cpp:431-446)
    ok = np.isfinite(lin).all(axis=-1)
    assert np.array_equal(lin[ok][:, [2, 1, 0]], img[..., :3][ok])


def test_is_illuminated_matches_the_oracle(P, oracle_mod, cornell_pair, cornell):
    """pgrt_is_illuminated = Raytracer::is_illuminated (pg1/raytracer.h:34, raytracer.cpp:150-176) as shipped, and the
    README's hard shadows (shadow_mode 1) against the oracle's statement of the same spec: exact agreement."""
    rt, o = cornell_pair
    rng = np.random.default_rng(32)
    n = 50000
    hit = rng.uniform(-80, 80, (n, 3)).astype(np.float32); hit[:, 2] = rng.uniform(0, 60, n)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    for light in (cornell.lights[0].position, (5.0, -60.0, 35.0)):
        for mode in (0, 1):
            a = rt.is_illuminated(light, hit, nrm, dict(shadow_mode=mode))
            b = o.is_illuminated(oracle_mod.make_params(shadow_mode=mode), light, hit, nrm)
            assert np.array_equal(a, b), (light, mode, (a != b).sum())
            assert 0.05 < a.mean() < 0.95


def test_hard_shadows_match_the_oracle(P, oracle_mod, cornell_pair, avenger):
    """shadow_mode = 1 (README.md:20 to-do, non-default): whole frames against the oracle's statement of the same rule."""
    for (rt, o), pd in ((cornell_pair, dict(seed=8, shadow_mode=1)), (avenger[1:], dict(sampling_width=1, jitter=0, aperture=0.0, shadow_mode=1))):
        ref, g0, p0, st0 = o.render(oracle_mod.make_params(**pd))
        img, st = rt.render(pd)
        ok, psnr = image_bars(P, ref, img)
        assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
        for k in ("primary", "shadow", "reflection", "refraction"):
            assert abs(st[k] - st0[k]) <= 1e-3 * max(st0[k], 1), (k, st[k], st0[k])
        plain, _ = rt.render(dict(pd, shadow_mode=0))
        assert (P.to_srgb8(img).astype(int).sum(-1) < P.to_srgb8(plain).astype(int).sum(-1) - 6).mean() > 0.005    # shadows did appear


# ------------------------------------------------------------------------------------------------ BASELINE.json configs at full size
def test_C2_full_frame_against_the_oracle(P, oracle_mod, avenger):
    """Config C2 (BASELINE.json configs[1]): avenger stand-in 1920x1080, Whitted, depth 10, 1 spp -- the whole frame against the
    oracle (all host threads), at the north-star bars: ids >= 99.9 %, <= 2 LSB on >= 99.5 %, PSNR >= 45 dB, ray counts 0.1 %."""
    sc, rt, o = avenger
    cam = sc.camera
    rt.set_camera(1920, 1080, cam.fov_y, cam.view_from, cam.view_at); o.set_camera(1920, 1080, cam.fov_y, cam.view_from, cam.view_at)
    try:
        pd = dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=10)
        ref, g0, p0, st0 = o.render(oracle_mod.make_params(**pd))
        img, st = rt.render(pd)
        g1, p1 = rt.primary_ids(pd)
        assert np.mean((g0 == g1) & (p0 == p1)) >= 0.999
        ok, psnr = image_bars(P, ref, img)
        assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
        for k in ("primary", "shadow", "reflection", "refraction"):
            assert abs(st[k] - st0[k]) <= 1e-3 * max(st0[k], 1), (k, st[k], st0[k])
        q, _ = rt.render_rgba8(pd)
        assert np.array_equal(q[..., :3], P.to_srgb8(img))
    finally:
        rt.set_camera(640, 480, cam.fov_y, cam.view_from, cam.view_at); o.set_camera(640, 480, cam.fov_y, cam.view_from, cam.view_at)


def test_C4_palm_grove_against_the_oracle(P, oracle_mod):
    """Config C4 (stand-in for PalmTrees): procedural palm grove, 1920x1080, Lambert + spherical env map, 1 spp -- thin fronds
    stress the builder; the frame against the oracle at the north-star bars."""
    sc = scenes.palm_grove(n_palms=150)
    rt = P.raytracer_for(sc); o = oracle_mod.Oracle(sc)
    assert sc.ntris > 500_000 and (rt.width, rt.height) == (1920, 1080)
    pd = dict(sampling_width=1, jitter=0, aperture=0.0, max_depth=7, shader_mode=1)
    ref, g0, p0, st0 = o.render(oracle_mod.make_params(**pd))
    img, st = rt.render(pd)
    g1, p1 = rt.primary_ids(pd)
    assert np.mean((g0 == g1) & (p0 == p1)) >= 0.999
    ok, psnr = image_bars(P, ref, img)
    assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
    for k in ("primary", "shadow", "reflection", "refraction"):
        assert abs(st[k] - st0[k]) <= 1e-3 * max(st0[k], 1), (k, st[k], st0[k])
    assert (g0 != INVALID).mean() > 0.2


def test_C5_soup_one_million_triangles_through_quantised_nodes(P, oracle_mod, monkeypatch):
    """Config C5 scaled to what the oracle builds in seconds: 1 M-triangle soup through the 80-B quantised nodes (forced, as
    the 10 M soup selects them), Lambert, 2x2 jittered samples on a 480x270 frame: ids, image bars and ray counts."""
    sc = scenes.triangle_soup(1_000_000, seed=1, resolution=(480, 270))
    monkeypatch.setenv("PGRT_NODE_LAYOUT", "q8")
    rt = P.raytracer_for(sc)
    monkeypatch.delenv("PGRT_NODE_LAYOUT")
    assert rt.build_stats["node_bytes"] == 80
    o = oracle_mod.Oracle(sc)
    pd = dict(sampling_width=2, jitter=1, aperture=0.0, max_depth=7, seed=1, shader_mode=1)
    ref, g0, p0, st0 = o.render(oracle_mod.make_params(**pd))
    img, st = rt.render(pd)
    g1, p1 = rt.primary_ids(pd)
    assert np.mean((g0 == g1) & (p0 == p1)) >= 0.999
    ok, psnr = image_bars(P, ref, img)
    assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
    for k in ("primary", "shadow"):
        assert abs(st[k] - st0[k]) <= 1e-3 * max(st0[k], 1), (k, st[k], st0[k])


def test_C3_dof_64spp_converges_to_the_high_spp_oracle(P, oracle_mod, avenger):
    """Config C3 in the wording of SURVEY 8(d): thin lens f = 200, a = 5, 8x8 stratified samples, compared as a converged mean:
    the 64-spp GPU frame of a 640x360 view against a 32x32 = 1024-spp oracle frame of a crop of it (different sample sets,
    so the bar is statistical: PSNR of the 8-bit crop >= 30 dB and mean difference below half an LSB), plus the exact ray count."""
    sc, rt, o = avenger
    cam = sc.camera
    rt.set_camera(640, 360, cam.fov_y, cam.view_from, cam.view_at); o.set_camera(640, 360, cam.fov_y, cam.view_from, cam.view_at)
    try:
        pd = dict(sampling_width=8, jitter=1, aperture=5.0, focal_distance=200.0, max_depth=7, seed=1)
        img, st = rt.render(pd)
        assert st["primary"] == 640 * 360 * 64
        x0, y0, x1, y1 = 260, 130, 420, 230
        ref = o.render(oracle_mod.make_params(**dict(pd, sampling_width=32, seed=5)), want_ids=False, region=(x0, y0, x1, y1))[0][y0:y1, x0:x1]
        a, b = P.to_srgb8(np.nan_to_num(img[y0:y1, x0:x1])).astype(float), P.to_srgb8(np.nan_to_num(ref)).astype(float)
        mse = np.mean((a - b) ** 2)
        assert 10 * np.log10(255.0 ** 2 / max(mse, 1e-9)) >= 30.0 and abs(a.mean() - b.mean()) < 0.5, (mse, a.mean(), b.mean())
        # the same 64 samples on both sides: the ordinary bars
        ref64, _, _, st0 = o.render(oracle_mod.make_params(**pd), want_ids=False, region=(x0, y0, x1, y1))
        ok, psnr = image_bars(P, ref64[y0:y1, x0:x1], img[y0:y1, x0:x1])
        assert ok >= 0.995 and psnr >= 45.0, (ok, psnr)
    finally:
        rt.set_camera(640, 480, cam.fov_y, cam.view_from, cam.view_at); o.set_camera(640, 480, cam.fov_y, cam.view_from, cam.view_at)
