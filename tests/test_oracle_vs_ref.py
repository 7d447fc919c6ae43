"""The CPU oracle (oracle/pg_oracle.cpp, a restatement) against the reference's OWN object code: oracle/_ref/libpg_ref.so is
built by oracle/ref.mk from the unmodified pg1/*.cpp files of /root/reference (stubs only for the Embree and FreeImage
binaries and the Win32 window, which the reference does not vendor).  Everything above the rtcIntersect1 / rtcInterpolate
boundary -- Raytracer::trace, is_illuminated, the ray makers, mix_srgb, Texture, SphericalMap, PinHoleCamera, gamma,
LoadOBJ / LoadMTL, tutorial_1 / tutorial_2 -- runs as the reference wrote it, so a misreading shared by the oracle and the
CUDA path cannot pass here.  The bar is BIT equality (the two sides call the same libm on the same host).  No GPU needed.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import golden
from pgi_raytracing_b200 import scenes

ref = pytest.importorskip("oracle.ref")
pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpg_ref.so is not built and /root/reference is absent")

IOR_AIR = np.float32(1.000293)
FLT_MAX = np.finfo(np.float32).max


def same(a, b):
    return (a == b) | (np.isnan(a) & np.isnan(b))


def scene_from_reference_loader(sc, directory):
    """The scene as the reference's LoadOBJ hands it to Embree (objloader.cpp:338 re-normalises every normal)."""
    obj, _ = ref.write_scene(sc, directory)
    surfaces, materials = ref.load_obj(obj)
    import copy
    out = copy.copy(sc)
    out.meshes = [scenes.Mesh(name, pos, nrm, uv, mi) for name, mi, pos, nrm, uv in surfaces]
    return out, surfaces, materials


@pytest.fixture(scope="module")
def cornell_pair(oracle_mod, tmp_path_factory):
    sc = scenes.cornell_like()
    sc.camera = scenes.Camera(64, 48, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
    d = str(tmp_path_factory.mktemp("ref_cornell"))
    loaded, _, _ = scene_from_reference_loader(sc, d)
    return sc, ref.Ref(sc, d), oracle_mod.Oracle(loaded), d


@pytest.fixture(scope="module")
def avenger_pair(oracle_mod, tmp_path_factory):
    sc = scenes.avenger_proxy(detail=0.35, env_size=(1000, 500))
    sc.camera = scenes.Camera(160, 120, sc.camera.fov_y, sc.camera.view_from, sc.camera.view_at)
    # the stand-in tiles some textures (u, v beyond 1): there the reference's Texture::get_pixel reads out of bounds
    # (texture.cpp:56-62 has no check) and nothing can be compared; keep the texture coordinates inside the image
    for m in sc.meshes:
        if sc.materials[m.material].diffuse_tex >= 0:
            m.uv = np.ascontiguousarray(np.clip(m.uv, 0.0, 0.999).astype(np.float32))
    d = str(tmp_path_factory.mktemp("ref_avenger"))
    loaded, _, _ = scene_from_reference_loader(sc, d)
    return sc, ref.Ref(sc, d), oracle_mod.Oracle(loaded), d


# ------------------------------------------------------------------------------------------------ the reference's own demos
def test_tutorial_1_prints_the_known_answer():
    """pg1/tutorials.cpp:147-165 run as compiled from the reference: the printf lines of :134-135."""
    out = ref.tutorial_1()
    assert "normal = (0.000, 0.000, 1.000)" in out and "tex_coord = (0.050, 0.933)" in out


def test_tutorial_2_prints_the_known_answer(tmp_path):
    """pg1/tutorials.cpp:170-178 on data/test4.png (bytes from the committed golden): (r = 1.000, g = 0.000, b = 0.500)."""
    buf = golden("test4_bgra.npy")
    img = scenes.Image(np.ascontiguousarray(buf), 64, buf.shape[0], buf.shape[1], 4)
    assert "(r = 1.000, g = 0.000, b = 0.500)" in ref.tutorial_2(img, str(tmp_path))


def test_deg2rad_is_the_camera_default():
    assert np.float32(ref._lib().ref_deg2rad(42.185)) == np.float32(scenes.Camera().fov_y)     # mymath.h:27-30, tutorials.cpp:188


# ------------------------------------------------------------------------------------------------ leaf functions
def test_mix_srgb_bit_exact(oracle_mod, cornell_pair):
    """utils.cpp:204-241 incl. the mis-scaled encode, the clamps and the net R<->B swap."""
    _, _, o, _ = cornell_pair
    rng = np.random.default_rng(0)
    c0 = rng.uniform(-0.3, 1.3, (200000, 4)).astype(np.float32); c1 = rng.uniform(-0.3, 1.3, (200000, 4)).astype(np.float32)
    c0[:2000] = rng.uniform(0, 0.05, (2000, 4)); c1[:2000] = rng.uniform(0, 0.005, (2000, 4))   # the linear segments of both curves
    al = rng.uniform(0, 1, 200000).astype(np.float32); al[:10] = [0, 1, 0.5, 1e-30, 1 - 1e-7, 0.25, 0.75, 0, 1, 0.5]
    assert np.array_equal(ref.mix_srgb(c0, c1, al), o.mix_srgb(c0, c1, al))


def test_gamma_bit_exact(cornell_pair):
    """raytracer.cpp:439-446: pow(c,g)*pow(c,g) per channel with the swap; NaN for negative input."""
    _, r, o, _ = cornell_pair
    c = np.random.default_rng(1).uniform(-0.2, 2.0, (100000, 4)).astype(np.float32)
    for g in (0.5, 0.45, 1.0, 0.1):
        r.set_gamma(g)
        a, b = r.gamma(c), o.gamma(c, g)
        assert same(a, b).all() and (np.isnan(a).any() or g == 1.0)
    r.set_gamma(0.5)


def test_texture_get_texel_bit_exact(cornell_pair, tmp_path):
    """texture.cpp:77-130 for u in [0,1), v in [0,1].  Outside -- and at u = 1 exactly, where x1 = W -- the reference reads
    past the end of the row (the first bytes of the next row or the pitch padding); the rewrite defines that by clamping the
    texel index, so those inputs are excluded here.  Integer coordinates and the first texel row / column
    return black; the wrap column gives weights outside [0,1]."""
    sc, _, o, _ = cornell_pair
    rng = np.random.default_rng(2)
    for t, img in enumerate(sc.textures):
        tex = ref.RefTexture(img, str(tmp_path / f"t{t}.png"))
        assert tex.size() == (img.width, img.height)
        uv = rng.uniform(0, 1, (100000, 2)).astype(np.float32)
        k = np.arange(2000)
        uv[:2000, 0] = (k % img.width) / np.float32(img.width)                # exact texel columns
        uv[2000:4000, 1] = (k % (img.height + 1)) / np.float32(img.height)    # exact texel rows incl. v = 1
        uv[4000:4100] = rng.uniform(0, 1.0 / max(img.width, img.height), (100, 2))   # x, y in (0,1): black
        uv[4100:4200, 0] = 1 - rng.uniform(0, 1e-3, 100)                              # wrap column
        uv[4200:4300, 1] = 1 - rng.uniform(0, 1e-3, 100)                              # last row
        a, b = tex.get_texel(uv), o.texture_get_texel(t, uv)
        assert np.array_equal(a, b)
        assert (a[:4100] == 0).all(axis=-1).sum() > 1000


def test_env_get_texel_bit_exact(cornell_pair, tmp_path):
    """SphericalMap.cpp:17-29: y/z swap, atan2 / asin through MSVC's float overloads then double arithmetic (resolves the
    'in double' wording of SURVEY A-10: with the global float overloads MSVC declares, the trig is float), final swap."""
    sc, _, o, _ = cornell_pair
    env = ref.RefEnv(sc.env, str(tmp_path / "env.jpg"))
    rng = np.random.default_rng(3)
    d = rng.normal(size=(200000, 3)).astype(np.float32)
    d[:6] = np.eye(3, dtype=np.float32).repeat(2, 0) * np.array([1, -1] * 3, np.float32)[:, None]   # the six axes (poles, seam)
    d[6:1000, 1] = 0
    assert np.array_equal(env.get_texel(d), o.env_get_texel(d))


def test_secondary_ray_makers_bit_exact(cornell_pair):
    """raytracer.cpp:178-235 incl. the NaN direction that marks total internal reflection."""
    _, r, o, _ = cornell_pair
    rng = np.random.default_rng(4)
    n = 100000
    items = np.zeros((n, 11), np.float32)
    items[:, 0:3] = rng.normal(size=(n, 3)); items[:, 3:6] = rng.normal(size=(n, 3)); items[:, 6:9] = rng.uniform(-100, 100, (n, 3))
    items[:, 9] = np.where(rng.random(n) < 0.5, IOR_AIR, 1.5); items[:, 10] = np.where(items[:, 9] == IOR_AIR, 1.5, IOR_AIR)
    for refr in (False, True):
        a, b = r.secondary_rays(items, refr), o.secondary_rays(items, refr)
        assert same(a, b).all()
    assert np.isnan(r.secondary_rays(items, True)[:, 4]).sum() > 1000


def test_camera_rays_bit_exact(oracle_mod, cornell_pair):
    """PinHoleCamera.cpp:5-105: both generate_ray overloads (integer W/2, H/2; no half-pixel; tnear 0.001 / 0.01)."""
    sc, r, o, _ = cornell_pair
    rng = np.random.default_rng(5)
    xy = rng.uniform(-2, 70, (5000, 2)).astype(np.float32)
    grid = np.stack(np.meshgrid(np.arange(64), np.arange(48)), -1).reshape(-1, 2).astype(np.float32)
    lens = r.generate_rays(grid, 200.0, 0.0)
    ours = o.primary_rays(oracle_mod.make_params(sampling_width=1, jitter=0, aperture=0.0))
    lens[:, 7] = IOR_AIR                                                     # get_pixel sets time afterwards (raytracer.cpp:416)
    assert np.array_equal(lens, ours)
    pin = r.generate_rays(grid, pinhole=True)
    ours = o.primary_rays(oracle_mod.make_params(sampling_width=1, jitter=0, camera_mode=1))
    pin[:, 7] = IOR_AIR
    assert np.array_equal(pin, ours)
    assert np.isfinite(r.generate_rays(xy, 150.0, 0.0)).all()


def test_lens_shift_is_a_square_of_side_aperture(cornell_pair):
    """PinHoleCamera.cpp:77-81: the clock-seeded shift is U[-a/2, a/2) on both lens axes (a SQUARE aperture); only its
    distribution can be compared.  |origin - view_from| <= a/2 * sqrt(2), and both lens coordinates use the full range."""
    sc, r, _, _ = cornell_pair
    xy = np.tile(np.array([[32.0, 24.0]], np.float32), (4000, 1))
    rays = r.generate_rays(xy, 200.0, 5.0)
    off = rays[:, :3] - np.asarray(sc.camera.view_from, np.float32)
    dist = np.linalg.norm(off, axis=1)
    assert dist.max() <= 2.5 * np.sqrt(2) + 1e-3 and dist.max() > 2.5 and len(np.unique(rays[:, 0])) > 100
    focus = rays[:, :3] + rays[:, 4:7] * np.linalg.norm(rays[0, :3] + 200 * rays[0, 4:7] - rays[:, :3], axis=1)[:, None]
    assert np.abs(rays[:, 3] - 0.01).max() == 0 and np.ptp(focus, axis=0).max() < 1.0     # every ray passes (close to) the one focal point


# ------------------------------------------------------------------------------------------------ loader
def test_load_obj_matches_the_written_scene(cornell_pair, tmp_path):
    """objloader.cpp:210-507: surfaces in group order with their usemtl material, un-indexed corners, normals re-normalised
    at load (:338); LoadMTL fields (Kd, Ks, Ns, Ni, shader, map_Kd)."""
    sc, _, _, _ = cornell_pair
    _, surfaces, materials = scene_from_reference_loader(sc, str(tmp_path))
    assert [s[0] for s in surfaces] == [m.name for m in sc.meshes] and [s[1] for s in surfaces] == [m.material for m in sc.meshes]
    for (name, mi, pos, nrm, uv), m in zip(surfaces, sc.meshes):
        assert np.array_equal(pos, m.pos) and np.array_equal(uv, m.uv)
        n2 = (m.nrm.astype(np.float32) ** 2).sum(-1, dtype=np.float32)
        assert np.abs(nrm - m.nrm).max() <= 1.2e-7 and np.abs((nrm ** 2).sum(-1) - 1).max() < 3e-7, name
    for got, want in zip(materials, sc.materials):
        assert got["name"] == want.name and got["type"] == want.type and got["has_diffuse_texture"] == (want.diffuse_tex >= 0)
        assert np.allclose(got["diffuse"], want.diffuse, rtol=0, atol=1e-7) and np.float32(got["ior"]) == np.float32(want.ior)


def test_reference_mtl_file_parses_as_the_survey_derived():
    """data/6887_allied_avenger.mtl through the reference's own LoadMTL: 5 materials, shader 3 x4 + 4 x1, and the malformed
    'Ks 1.0. 1.0 1.0' line leaves specular = (1.0, 0.8, 0.8)."""
    mtl = "/root/reference/data/6887_allied_avenger.mtl"
    if not os.path.exists(mtl):
        pytest.skip("reference data is not on this box")
    import tempfile
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "one.obj"), "w") as f:
        f.write("mtllib 6887_allied_avenger.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvn 0 0 1\nvn 0 0 1\nvt 0 0 0\nvt 1 0 0\nvt 0 1 0\ng a\nusemtl white_plastic\nf 1/1/1 2/2/2 3/3/3\n")
    import shutil
    shutil.copy(mtl, d)
    _, materials = ref.load_obj(os.path.join(d, "one.obj"))
    assert len(materials) == 5 and sorted(m["type"] for m in materials) == [3, 3, 3, 3, 4]
    for m in materials:
        assert tuple(np.float32(x) for x in m["specular"]) == (np.float32(1.0), np.float32(0.8), np.float32(0.8)), m


# ------------------------------------------------------------------------------------------------ is_illuminated / trace
def test_is_illuminated_bit_exact(oracle_mod, cornell_pair):
    """raytracer.cpp:150-176 + LightSource.cpp:11-32: the n.L_pos early-out, the ray that leaves the light with the hit
    POSITION as its direction, closest occluder decides, a dielectric occluder does not shadow."""
    sc, r, o, _ = cornell_pair
    rng = np.random.default_rng(6)
    n = 20000
    hit = rng.uniform(-80, 80, (n, 3)).astype(np.float32); hit[:, 2] = rng.uniform(0, 60, n)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    light = np.array(sc.lights[0].position, np.float32)
    p = oracle_mod.make_params(sampling_width=1, jitter=0, aperture=0.0)
    a, b = r.is_illuminated(light, hit, nrm), o.is_illuminated(p, light, hit, nrm)
    assert np.array_equal(a, b) and 0.2 < a.mean() < 0.8
    # a light inside the scene, so that occluders (Phong and dielectric) are actually met
    light2 = np.array([5, -60, 35], np.float32)
    a, b = r.is_illuminated(light2, hit, nrm), o.is_illuminated(p, light2, hit, nrm)
    assert np.array_equal(a, b)
    assert (a != (nrm @ light2 >= 0)).any()          # some queries are decided by an occluder, not by the early-out


@pytest.mark.parametrize("which", ["cornell", "avenger"])
def test_trace_bit_exact_at_every_level(oracle_mod, cornell_pair, avenger_pair, which):
    """Raytracer::trace(ray, level) (raytracer.cpp:237-394) on caller-supplied rays at levels 0..7: camera rays, rays from
    inside the scene, rays that start in glass (time = the material's IOR)."""
    sc, r, o, _ = cornell_pair if which == "cornell" else avenger_pair
    p = oracle_mod.make_params(sampling_width=1, jitter=0, aperture=0.0)
    rng = np.random.default_rng(7)
    prim = o.primary_rays(p)
    n = 6000
    rays = np.zeros((n, 9), np.float32)
    rays[:, 0:3] = rng.uniform(-60, 60, (n, 3)); rays[:, 2] = rng.uniform(2, 70, n)
    rays[:, 4:7] = rng.normal(size=(n, 3)); rays[:, 3] = 0.01; rays[:, 8] = FLT_MAX
    rays[:, 7] = np.where(rng.random(n) < 0.7, IOR_AIR, 1.5)
    allrays = np.concatenate([prim[rng.integers(0, prim.shape[0], 6000)], rays])
    diel = 0
    for level in range(0, 8):
        a, b = r.trace(allrays, level), o.trace(p, allrays, level)
        assert same(a, b).all(), (which, level, np.nonzero(~same(a, b).all(axis=-1))[0][:5])
        assert (a[:, 3] == 1).all()
        diel += int((a != r.trace(allrays[:1], 7)[0]).any())
    assert (r.trace(allrays, 7)[:, :3] == 0).all(axis=-1).mean() > 0.3     # level >= 7 on a hit returns black (:282-283)


@pytest.mark.parametrize("which", ["cornell", "avenger"])
def test_unjittered_frame_bit_exact(oracle_mod, cornell_pair, avenger_pair, which):
    """One Producer iteration (simpleguidx11.cpp:105-118) of get_pixel with one un-jittered sample (raytracer.cpp:405-436):
    camera ray, trace, sum, swap, gamma, pixel packing -- every pixel of the frame, bit for bit."""
    sc, r, o, _ = cornell_pair if which == "cornell" else avenger_pair
    p = oracle_mod.make_params(sampling_width=1, jitter=0, aperture=0.0)
    a = r.render_unjittered(200.0)
    b, geom, _, st = o.render(p)
    assert same(a, b).all()
    types = {sc.materials[sc.meshes[g].material].type for g in np.unique(geom) if g != 0xFFFFFFFF}
    assert types == {3, 4} and (geom == 0xFFFFFFFF).any() and st["refraction"] > 0 and st["shadow"] > 0


def test_shipped_get_pixel_converges_to_the_oracle_mean(oracle_mod, cornell_pair):
    """Raytracer::get_pixel exactly as shipped (3x3 strata, clock-seeded mt19937 jitter U[-1/6,1/6) and square lens shift
    U[-2.5,2.5), raytracer.cpp:398-416) cannot be compared sample for sample; the mean of many of its frames must agree with
    the mean of as many oracle frames (counter-based RNG, different seeds) within the Monte-Carlo error."""
    sc, r, o, _ = cornell_pair
    x0, y0, x1, y1 = 20, 14, 44, 34
    n = 24
    a = np.mean([np.nan_to_num(r.get_pixels(x0, y0, x1, y1)) for _ in range(n)], axis=0)
    frames = [np.nan_to_num(o.render(oracle_mod.make_params(seed=100 + k), want_ids=False, region=(x0, y0, x1, y1))[0][y0:y1, x0:x1]) for k in range(n)]
    b = np.mean(frames, axis=0)
    sd = np.std(frames, axis=0) / np.sqrt(n)
    z = np.abs(a - b)[..., :3] / (np.sqrt(2) * sd[..., :3] + 2e-3)
    assert np.mean(z < 4) > 0.99 and np.abs(a.mean() - b.mean()) < 2e-3, (np.mean(z < 4), a.mean(), b.mean())
